#!/bin/bash
# compute-sanitizer passes over the small-shape kernel tests (memcheck + racecheck); slow, run on demand
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual or layernorm_folded or attention" > gpurun_out/sanitize_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual_layernorm_stats or layernorm_folded" > gpurun_out/sanitize_racecheck.log 2>&1
tail -n 15 gpurun_out/sanitize_memcheck.log; tail -n 15 gpurun_out/sanitize_racecheck.log
