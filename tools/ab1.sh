mkdir -p gpurun_out
for v in default g1 l2q16 l2q20 l3q16 l3q20 l3q24 l4q24; do
  if [ $v = default ]; then lib=lzs-compression_b200/liblzs.so; else lib=variants/$v.so; fi
  echo "== $v"
  LZS_B200_LIB=$PWD/$lib timeout 120 python tools/prof.py --mib 1024 --kind text,binary,random,mixed --iters 2 --time 2>&1 | grep -v "iter 0" | tail -6 | cut -c1-60
done > gpurun_out/t6_ab.log 2>&1
cat gpurun_out/t6_ab.log
