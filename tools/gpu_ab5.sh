#!/bin/bash
mkdir -p gpurun_out
for s in 36 73 512; do
  echo "=== seqs $s" >> gpurun_out/ab5_kpower.log
  MASKBIT_B200_LIB=tools/lib_new.so timeout 300 python tools/kpower.py --seconds 2.5 --seqs $s >> gpurun_out/ab5_kpower.log 2>&1
done
