#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/ab6_pytest_kernels.log 2>&1
for v in new2 new; do
  echo "=== $v" >> gpurun_out/ab6_kpower.log
  MASKBIT_B200_LIB=tools/lib_$v.so timeout 300 python tools/kpower.py --seconds 2.5 --only "" >> gpurun_out/ab6_kpower.log 2>&1
done
timeout 900 python bench.py > gpurun_out/ab6_bench.json 2> gpurun_out/ab6_bench.err
timeout 600 python -m pytest tests/test_gpu_path.py -m gpu -x -q > gpurun_out/ab6_pytest_path.log 2>&1
