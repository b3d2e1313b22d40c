#!/bin/bash
mkdir -p gpurun_out
echo "== old"; MASKBIT_B200_LIB=tools/lib_old.so python tools/kbench.py --iters 10 2>&1 | tee gpurun_out/r02g_kbench_old.txt
echo "== new"; python tools/kbench.py --iters 10 2>&1 | tee gpurun_out/r02g_kbench_new.txt
echo "== new, no ping-pong"; MASKBIT_B200_LIB=tools/lib_nopp.so python tools/kbench.py --iters 10 --only attention 2>&1 | tee gpurun_out/r02g_kbench_nopp.txt
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r02g_pytest_gpu.log; grep -E "return_attn" gpurun_out/r02g_pytest_gpu.log | head -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02g_smoke.log
