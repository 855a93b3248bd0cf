#!/bin/sh
# Everything profiles/ is made from, in one GPU-box call (run from the repo root under gpurun):
#   tests, both bench arms, the ncu launch list of the bench, one full ncu capture per kernel,
#   the packets / chunk-size sweep.  Outputs land in gpurun_out/.
R=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${R}_gpu_tests.log 2>&1; tail -2 gpurun_out/${R}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --impl reference > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err
timeout 600 python bench.py > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err
cat gpurun_out/${R}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/${R}_bench_under_ncu.log 2>&1
for k in k1_match k23_parse_pack k4_decode; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/${R}_$k \
      python tools/prof.py --mib 1024 --kind mixed --iters 1 > /dev/null 2>&1
done
timeout 900 python tools/sweep.py > gpurun_out/${R}_sweep.jsonl 2> gpurun_out/${R}_sweep.err; tail -3 gpurun_out/${R}_sweep.jsonl
ls -la gpurun_out | tail -12
timeout 600 python tools/inc_device_bench.py --flows 1048576 --sample 256 > gpurun_out/${R}_incremental_device.json 2> gpurun_out/${R}_incremental_device.err; cat gpurun_out/${R}_incremental_device.json
timeout 300 python tools/flows_bench.py > gpurun_out/${R}_flows_bulk.json 2> gpurun_out/${R}_flows_bulk.err; cat gpurun_out/${R}_flows_bulk.json
timeout 400 python tools/pieces_bench.py > gpurun_out/${R}_pieces_bench.jsonl 2> gpurun_out/${R}_pieces_bench.err; cut -c1-200 gpurun_out/${R}_pieces_bench.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${R}_pieces_launches.csv \
    python tools/prof.py --mib 1024 --chunk 1048576 --iters 2 --compress-only > /dev/null 2>&1
timeout 200 python tools/zeros_probe.py 2>&1 | tail -2
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${R}_decode_pieces_launches.csv \
    python tools/decode_prof.py --mib 1024 --chunk 1048576 --iters 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k4p_copy -c 1 -f -o gpurun_out/${R}_k4p_copy \
    python tools/decode_prof.py --mib 256 --chunk 1048576 --iters 1 > /dev/null 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize.py 2>&1 | tail -7 > gpurun_out/${R}_memcheck.txt; tail -3 gpurun_out/${R}_memcheck.txt
timeout 600 compute-sanitizer --tool synccheck python tools/sanitize.py 2>&1 | tail -3 > gpurun_out/${R}_synccheck.txt; tail -2 gpurun_out/${R}_synccheck.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${R}_decode_jump_launches.csv \
    python tools/decode_prof.py --mib 1024 --chunk 268435456 --iters 1 --jump > /dev/null 2>&1
