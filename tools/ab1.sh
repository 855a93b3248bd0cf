mkdir -p gpurun_out
nproc; free -g | head -2
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t9_gpu_tests.log 2>&1; tail -5 gpurun_out/t9_gpu_tests.log
timeout 600 python bench.py > gpurun_out/t9_bench.json 2> gpurun_out/t9_bench.err; tail -3 gpurun_out/t9_bench.err; cat gpurun_out/t9_bench.json
