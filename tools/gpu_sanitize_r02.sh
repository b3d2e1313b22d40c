#!/bin/bash
# compute-sanitizer over the code added in round 2: small-M GEMM schedule with the shared statistics slot, folded upsample convs +
# rewritten conv_out (decode golden), select on a thread-block cluster (DSMEM stores)
mkdir -p gpurun_out
timeout 300 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "residual_layernorm_stats or layernorm_folded" > gpurun_out/r02_sanitize_memcheck_gemm.log 2>&1; tail -n 3 gpurun_out/r02_sanitize_memcheck_gemm.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "decode_matches_reference_golden or select_vs_torch_oracle or encode_matches" > gpurun_out/r02_sanitize_memcheck_path.log 2>&1; tail -n 3 gpurun_out/r02_sanitize_memcheck_path.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "select_vs_torch_oracle" > gpurun_out/r02_sanitize_racecheck_select.log 2>&1; tail -n 4 gpurun_out/r02_sanitize_racecheck_select.log
