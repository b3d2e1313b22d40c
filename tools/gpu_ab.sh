#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -s -k "attention" > gpurun_out/r02h_pytest_attention.log 2>&1; echo "pytest attention rc=$?"; tail -3 gpurun_out/r02h_pytest_attention.log
echo "== old"; MASKBIT_B200_LIB=tools/lib_old.so python tools/kbench.py --iters 10 --only attention 2>&1 | tee gpurun_out/r02h_kbench_old.txt
echo "== prev (no half split)"; MASKBIT_B200_LIB=tools/lib_nocenter.so python tools/kbench.py --iters 10 --only attention 2>&1
echo "== new"; python tools/kbench.py --iters 10 --only attention 2>&1 | tee gpurun_out/r02h_kbench_new.txt
echo "== trace"; MASKBIT_B200_LIB=tools/lib_trace.so python tools/attn_trace.py check 2>&1 | tee gpurun_out/r02h_attn_trace.txt
timeout 900 python -m pytest tests/test_gpu_path.py -x -q -k "forward_matches or stepwise" 2>&1 | tail -2
