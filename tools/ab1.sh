mkdir -p gpurun_out
LZS_B200_LIB=$PWD/variants/bulk.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -3
for v in default bulk; do
  if [ $v = default ]; then lib=lzs-compression_b200/liblzs.so; else lib=variants/$v.so; fi
  echo "== $v"
  LZS_B200_LIB=$PWD/$lib timeout 200 python tools/prof.py --mib 1024 --kind text,binary,random,mixed --iters 3 --time 2>&1 | grep "iter 2" | cut -c1-64
  LZS_B200_LIB=$PWD/$lib timeout 200 python tools/prof.py --mib 1024 --chunk 1500 --kind packet --iters 3 --time 2>&1 | grep "iter 2" | cut -c1-64
done > gpurun_out/t12_ab.log 2>&1
cat gpurun_out/t12_ab.log
timeout 600 python tools/inc_device_bench.py --flows 1048576 --sample 256 > gpurun_out/t12_inc.json 2> gpurun_out/t12_inc.err; tail -5 gpurun_out/t12_inc.err; cat gpurun_out/t12_inc.json
