#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/ab13_pytest_kernels.log 2>&1
for v in tma ldg; do
  echo "=== $v" >> gpurun_out/ab13_kpower.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kpower.py --seconds 2.5 --only gemm_ >> gpurun_out/ab13_kpower.log 2>&1
  echo "=== $v" >> gpurun_out/ab13_kbench.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kbench.py --iters 8 --only gemm >> gpurun_out/ab13_kbench.log 2>&1
done
timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "forward or teacher or stepwise" > gpurun_out/ab13_pytest_path.log 2>&1
