#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/ab2_pytest_kernels.log 2>&1
for v in base new nopre high nomma noepi; do
  echo "=== $v" >> gpurun_out/ab2_kbench.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kbench.py --iters 8 --only gemm >> gpurun_out/ab2_kbench.log 2>&1
done
MASKBIT_B200_LIB=tools/lib_trace.so timeout 300 python tools/gemm_trace.py > gpurun_out/ab2_trace.log 2>&1
MASKBIT_B200_LIB=tools/lib_traceh.so timeout 300 python tools/gemm_trace.py > gpurun_out/ab2_trace_high.log 2>&1
MASKBIT_B200_LIB=tools/lib_atrace.so timeout 300 python tools/attn_trace.py > gpurun_out/ab2_attn_trace.log 2>&1
timeout 900 python bench.py > gpurun_out/ab2_bench.json 2> gpurun_out/ab2_bench.err
timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -x -q > gpurun_out/ab2_pytest_path.log 2>&1
ls -la gpurun_out
