#!/bin/bash
mkdir -p gpurun_out
for v in new3 noresld; do
  echo "=== $v" >> gpurun_out/ab8_kpower.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kpower.py --seconds 2.5 --only gemm_x --shapes gemm_x_out_epi7:1024:1024:7,gemm_x_down_epi7:1024:4096:7,gemm_x_out_epi5:1024:1024:5 >> gpurun_out/ab8_kpower.log 2>&1
done
