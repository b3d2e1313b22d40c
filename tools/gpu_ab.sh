#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02j_pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/r02j_pytest_gpu.log
echo "== PDL on"; python tools/latency_probe.py 2>&1 | tee gpurun_out/r02j_latency_pdl1.txt
echo "== PDL off"; MASKBIT_B200_PDL=0 python tools/latency_probe.py 2>&1 | tee gpurun_out/r02j_latency_pdl0.txt
echo "== sweep 1 GPU"; python tools/sweep.py > gpurun_out/r02_sweep_12bit_1gpu.jsonl 2> gpurun_out/r02_sweep_12bit_1gpu.err; cat gpurun_out/r02_sweep_12bit_1gpu.jsonl
