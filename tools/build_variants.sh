#!/bin/bash
# Builds A/B variants of the library into tools/lib_<name>.so (git-ignored, shipped to the GPU box by gpurun).
# usage: tools/build_variants.sh name1="-DFOO=1 -DBAR=2" name2="" ...
set -e
cd "$(dirname "$0")/../maskbit_b200/csrc"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
pids=()
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  [ "$name" = "$spec" ] && flags=""
  $NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared $flags \
      -o ../../tools/lib_$name.so api.cu &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
ls -la ../../tools/lib_*.so
