#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train_fwd.py -x -q -s -k "attention or train or mlm" > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02f_pytest.log; grep -E "MLMLoss" gpurun_out/r02f_pytest.log | head -4
echo "== old"; MASKBIT_B200_LIB=tools/lib_old.so python tools/kbench.py --iters 10 --only attention 2>&1 | tee gpurun_out/r02f_kbench_old.txt
echo "== new"; python tools/kbench.py --iters 10 --only attention 2>&1 | tee gpurun_out/r02f_kbench_new.txt
echo "== trace"; MASKBIT_B200_LIB=tools/lib_trace.so python tools/attn_trace.py check 2>&1 | tee gpurun_out/r02f_attn_trace.txt
timeout 900 python -m pytest tests/test_gpu_path.py -x -q -k "forward_matches or sample_device or stepwise" 2>&1 | tail -2
