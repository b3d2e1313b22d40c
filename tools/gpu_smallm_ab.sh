#!/bin/bash
# small-M GEMM schedule (1-CTA BN = 128 tiles when they fit in one wave): kernel + forward tests, then latency with it off / on
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -s -k "gemm or forward or teacher or sample_matches or smoke" > gpurun_out/smallm_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/smallm_pytest.log; grep -E "sequences:|small-M" gpurun_out/smallm_pytest.log
echo "== SMALL_M=0"; MASKBIT_B200_SMALL_M=0 timeout 300 python tools/latency_probe.py 2>&1 | grep -v "^   " | tee gpurun_out/smallm_latency_0.txt
echo "== SMALL_M=1"; timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/smallm_latency_1.txt
