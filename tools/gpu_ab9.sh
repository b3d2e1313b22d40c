#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "eval_driver or prenorm" > gpurun_out/ab9_pytest.log 2>&1
for v in g0 g1 g2; do
  echo "=== $v" >> gpurun_out/ab9_kpower.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kpower.py --seconds 2.5 --only gemm_up >> gpurun_out/ab9_kpower.log 2>&1
done
for v in g0 g1 g2; do
  echo "=== $v" >> gpurun_out/ab9_kbench.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kbench.py --iters 8 --only gemm_up >> gpurun_out/ab9_kbench.log 2>&1
done
