mkdir -p gpurun_out
for v in default s1 s2 s4 s6 s8 s3h8 s2h6 s4q12 s4l2; do
  if [ $v = default ]; then lib=lzs-compression_b200/liblzs.so; else lib=variants/$v.so; fi
  echo "== $v"
  LZS_B200_LIB=$PWD/$lib timeout 120 python tools/prof.py --mib 1024 --kind text,binary,random,mixed --iters 2 --time 2>&1 | grep -v "iter 0" | tail -6 | cut -c1-60
done > gpurun_out/t5_ab.log 2>&1
cat gpurun_out/t5_ab.log
