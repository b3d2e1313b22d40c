#!/bin/bash
# attention variants on one box: correctness (test_attention, incl. the extreme-logit cases), isolated time, sustained J / launch
mkdir -p gpurun_out
for v in base p4 p6 p8 f0 f4 f6 old; do
  lib=tools/lib_$v.so; [ $v = base ] && lib=maskbit_b200/csrc/libmaskbit_b200.so
  [ -f $lib ] || continue
  echo "== $v"
  if [ $v != old ]; then MASKBIT_B200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "test_attention" 2>&1 | tail -2; fi
  MASKBIT_B200_LIB=$lib timeout 120 python tools/kbench.py --only attention --iters 20
  MASKBIT_B200_LIB=$lib timeout 120 python tools/kpower.py --only attention --seconds 2 2>&1 | grep attention
done 2>&1 | tee gpurun_out/attn_ab.txt
