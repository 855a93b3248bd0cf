mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k1_match -c 1 -f -o gpurun_out/t3_k1 \
   python tools/prof.py --mib 256 --kind text --iters 1 > gpurun_out/t3_ncu.log 2>&1
tail -3 gpurun_out/t3_ncu.log
ls -la gpurun_out/t3_k1.ncu-rep
