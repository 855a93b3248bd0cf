mkdir -p gpurun_out
for lanes in 8 4 16; do
  echo "== lanes $lanes"
  timeout 200 python tools/prof.py --mib 1024 --lanes $lanes --kind text,binary,random,mixed --iters 2 --time 2>&1 | grep -v "iter 0" | tail -6 | cut -c60-140
  timeout 200 python tools/prof.py --mib 1024 --lanes $lanes --chunk 1500 --kind packet --iters 2 --time 2>&1 | grep -v "iter 0" | tail -2 | cut -c60-140
  timeout 200 python tools/prof.py --mib 1024 --lanes $lanes --chunk 4096 --kind mixed --iters 2 --time 2>&1 | grep -v "iter 0" | tail -2 | cut -c60-140
done > gpurun_out/t8_ab.log 2>&1
cat gpurun_out/t8_ab.log
