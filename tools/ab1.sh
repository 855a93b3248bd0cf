mkdir -p gpurun_out
for o in 1 0; do
  echo "== order $o"
  LZS_B200_ORDER=$o timeout 200 python tools/prof.py --mib 1024 --kind text,binary,random,mixed --iters 3 --time 2>&1 | grep "iter 2" | cut -c70-140
  LZS_B200_ORDER=$o timeout 200 python tools/prof.py --mib 1024 --chunk 1500 --kind packet --iters 3 --time 2>&1 | grep "iter 2" | cut -c70-140
  LZS_B200_ORDER=$o timeout 200 python tools/prof.py --mib 1024 --chunk 4096 --kind mixed --iters 3 --time 2>&1 | grep "iter 2" | cut -c70-140
  LZS_B200_ORDER=$o timeout 200 python tools/prof.py --mib 1024 --chunk 262144 --kind mixed --iters 3 --time 2>&1 | grep "iter 2" | cut -c70-140
done > gpurun_out/t14_ab.log 2>&1
cat gpurun_out/t14_ab.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -2
