#!/bin/bash
mkdir -p gpurun_out
echo "=== new + shapes" >> gpurun_out/ab4b_kpower.log
MASKBIT_B200_LIB=tools/lib_new.so timeout 300 python tools/kpower.py --seconds 2.5 --only gemm_x --shapes gemm_x_out_epi0:1024:1024:0,gemm_x_out_epi5:1024:1024:5,gemm_x_out_epi7:1024:1024:7,gemm_x_down_epi0:1024:4096:0,gemm_x_down_epi5:1024:4096:5,gemm_x_down_epi7:1024:4096:7,gemm_x_qkv_epi0:3072:1024:0,gemm_x_qkv_epi5:3072:1024:5 >> gpurun_out/ab4b_kpower.log 2>&1
