#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "gemm or forward or sample" > gpurun_out/pf_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pf_pytest.log
echo "== L2_PREFETCH=0"; MASKBIT_B200_L2_PREFETCH=0 timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/pf_latency_0.txt
echo "== L2_PREFETCH=1"; timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/pf_latency_1.txt
