#!/bin/bash
# Builds libslzw_<name>.so with extra -D flags for A/B runs on the GPU box (SLZW_LIB=<path> selects one).
#   tools/build_variants.sh name1 "-DSLZW_U0=2" name2 "-DSLZW_HIT_REDUX=0" ...
set -e
cd "$(dirname "$0")/../lzw_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  $NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC,-fvisibility=hidden $flags -c -o variants/enc_$name.o encode_kernels.cu
  $NVCC $ARCH -shared -o variants/libslzw_$name.so slzw_api.o slzw_multi.o variants/enc_$name.o decode_kernels.o sched_kernels.o predictor_kernels.o -cudart static -lpthread
  echo built variants/libslzw_$name.so
done
