mkdir -p gpurun_out
for v in default eager eq20 q18 q20 t384 t320 u1; do
  if [ $v = default ]; then lib=lzs-compression_b200/liblzs.so; else lib=variants/$v.so; fi
  echo "== $v"
  LZS_B200_LIB=$PWD/$lib timeout 200 python tools/prof.py --mib 1024 --kind text,binary,random,mixed --iters 3 --time 2>&1 | grep "iter 2" | cut -c1-48
done > gpurun_out/t13_ab.log 2>&1
cat gpurun_out/t13_ab.log
