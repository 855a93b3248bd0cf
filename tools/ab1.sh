for v in q8 q12; do
  echo "== $v"; LZS_B200_LIB=$PWD/variants/$v.so timeout 200 python tools/k1_time_only.py random,text 2>&1 | tail -2
done
