#!/bin/bash
# One gpurun call: all GPU tests with the working-tree library (-s: achieved errors / agreement printed), kernel timings old vs
# new on the same box, the attention timeline, and a bench line.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s -x > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -5 gpurun_out/r02c_pytest_gpu.log
echo "== old"; MASKBIT_B200_LIB=tools/lib_old.so python tools/kbench.py --iters 10 2>&1 | tee gpurun_out/r02c_kbench_old.txt
echo "== new"; python tools/kbench.py --iters 10 2>&1 | tee gpurun_out/r02c_kbench_new.txt
echo "== trace"; MASKBIT_B200_LIB=tools/lib_trace.so python tools/attn_trace.py check 2>&1 | tee gpurun_out/r02c_attn_trace.txt
echo "== bench"; python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; cat gpurun_out/r02c_bench.json
