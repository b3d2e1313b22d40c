#!/bin/bash
# NOTE: the GEMM_EPI7_PIPE variant this script compares was measured neutral and removed from the tree; kept as the record of how
# profiles/r02_epi7_pipe_neutral.txt was made.
# residual-epilogue pipelining (GEMM_EPI7_PIPE) off / on and the rewritten conv_out kernel: kernel tests, decode parity, isolated and
# sustained GEMM timings, tokenizer bench -- one box
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "gemm or decode or encode or forward_12bit or teacher" > gpurun_out/e7_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/e7_pytest.log
for v in e7off base; do
  lib=tools/lib_$v.so; [ $v = base ] && lib=maskbit_b200/csrc/libmaskbit_b200.so
  echo "== $v"
  MASKBIT_B200_LIB=$lib timeout 200 python tools/kbench.py --only gemm_ --iters 10 | grep -E "out|down"
  MASKBIT_B200_LIB=$lib timeout 200 python tools/kpower.py --only gemm_out --seconds 2 2>&1 | grep gemm_
  MASKBIT_B200_LIB=$lib timeout 200 python tools/kpower.py --only gemm_down --seconds 2 2>&1 | grep gemm_
done 2>&1 | tee gpurun_out/e7_ab.txt
timeout 600 python bench.py --workload tokenizer --steps 3 --warmup 3 > gpurun_out/e7_tok.json 2> gpurun_out/e7_tok.err; cut -c1-130 gpurun_out/e7_tok.json
