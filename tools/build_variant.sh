#!/bin/sh
# tools/build_variant.sh NAME -DFLAG...   ->  variants/NAME.so  (select with LZS_B200_LIB=variants/NAME.so)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
name=$1; shift
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC "$@" \
    -shared -o variants/$name.so lzs-compression_b200/csrc/*.cu -lcudart
