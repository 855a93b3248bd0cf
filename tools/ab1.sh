mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 -k "pinned or slots or packed" 2>&1 | tail -3
for z in 1 0; do echo "== zerocopy $z"; LZS_B200_ZEROCOPY=$z timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-pageable 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['decompress_gbs'], d['e2e'])"; done
LZS_B200_TRACE=1 timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-pageable 2>&1 | grep "trace" | tail -40
