#!/bin/bash
mkdir -p gpurun_out
for v in new base; do
  echo "=== $v" >> gpurun_out/ab3_kpower.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kpower.py --seconds 3 >> gpurun_out/ab3_kpower.log 2>&1
done
