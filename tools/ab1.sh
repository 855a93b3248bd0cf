for pp in 1 0; do echo "== persistent $pp"; LZS_B200_K23_PERSISTENT=$pp timeout 200 python tools/compress_time.py text,binary,random,mixed 2>&1 | tail -4; done
