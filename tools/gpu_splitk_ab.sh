#!/bin/bash
# NOTE: the split-K schedule this script toggles (MASKBIT_B200_SPLITK) was measured slower and removed from the tree (see git history,
# commit "Small batches: select step on a thread-block cluster ..."); kept as the record of how profiles/r02_splitk_rejected.txt was made.
# split-K schedule of the CTA-pair GEMM for small M: kernel tests, then small-batch latency with it off / on (one box)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x > gpurun_out/splitk_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/splitk_pytest.log
echo "== SPLITK=0"; MASKBIT_B200_SPLITK=0 timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/splitk_latency_0.txt
echo "== SPLITK=1"; timeout 300 python tools/latency_probe.py 2>&1 | tee gpurun_out/splitk_latency_1.txt
