mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/r2_gpu_tests.log 2>&1; tail -2 gpurun_out/r2_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; python -c "
import json; d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1]); r=json.loads(open('gpurun_out/r2_bench_reference.json').read().strip().splitlines()[-1]); print('value',d['value'],'e2e',d['e2e']['value'],'ref',r['value'],'ratio e2e',d['e2e']['value']/r['value'], d['roofline']['per_kernel_ms'], 'parity', d['parity']['mismatches'], 'launches', d['gpu_launches'])"
