#!/bin/sh
# Build a tuning variant of the CUDA library next to the product: scripts/mkvariant.sh NAME -DFOO=1 ...
# -> lisa_b200/variants/liblisa_rt_NAME.so (git-ignored; select with LISA_RT_LIB=<path>).
set -e
cd "$(dirname "$0")/../lisa_b200"
name=$1; shift
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -fmad=false -lineinfo -std=c++17 -Xcompiler -fPIC \
  -ccbin "$(command -v /usr/bin/g++ || command -v g++)" "$@" -shared -o variants/liblisa_rt_$name.so \
  csrc/lisa_rt.cu csrc/bvh_build.cu csrc/estimator.cu csrc/devmem.cu csrc/sort_scan.cu csrc/multi.cu csrc/split.cu -ldl
echo variants/liblisa_rt_$name.so
