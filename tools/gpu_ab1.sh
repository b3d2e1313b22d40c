#!/bin/bash
# one gpurun call: GEMM A/B variants, traces, ncu source capture, kernel parity tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/ab1_gpu.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/ab1_pytest_kernels.log 2>&1
for v in base new nopre st5 nomma noepi nostore; do
  echo "=== $v" >> gpurun_out/ab1_kbench.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kbench.py --iters 8 --only gemm >> gpurun_out/ab1_kbench.log 2>&1
done
echo "=== new (attention too)" >> gpurun_out/ab1_kbench.log
timeout 300 python tools/kbench.py --iters 8 >> gpurun_out/ab1_kbench.log 2>&1
MASKBIT_B200_LIB=tools/lib_trace.so timeout 300 python tools/gemm_trace.py > gpurun_out/ab1_trace_prefetch.log 2>&1
MASKBIT_B200_LIB=tools/lib_trace0.so timeout 300 python tools/gemm_trace.py > gpurun_out/ab1_trace_noprefetch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm2_bf16 -c 12 -o gpurun_out/ab1_gemm2 \
    python tools/kbench.py --iters 1 --only gemm > gpurun_out/ab1_ncu.log 2>&1
ls -la gpurun_out
