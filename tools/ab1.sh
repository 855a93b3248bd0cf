mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_incremental.py -m gpu -x -q --timeout 300 2>&1 | tail -3
timeout 300 python tools/flows_bench.py > gpurun_out/r2_flows_bulk.json 2> gpurun_out/r2_flows_bulk.err; tail -3 gpurun_out/r2_flows_bulk.err; cat gpurun_out/r2_flows_bulk.json
