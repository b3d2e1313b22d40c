#!/bin/bash
# Build libmaskbit_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -shared ${MB_NVCC_EXTRA} \
    -o libmaskbit_b200.so api.cu
