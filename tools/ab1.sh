mkdir -p gpurun_out
ASAN=$(gcc -print-file-name=libasan.so)
LD_PRELOAD=$ASAN ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 LZS_B200_LIB=$PWD/variants/asan.so timeout 900 python tools/sanitize.py > gpurun_out/r2_host_asan_ubsan.log 2>&1
echo "rc=$?"; grep -E "workload|ERROR: AddressSanitizer|runtime error|SUMMARY" gpurun_out/r2_host_asan_ubsan.log | head -12; tail -3 gpurun_out/r2_host_asan_ubsan.log
