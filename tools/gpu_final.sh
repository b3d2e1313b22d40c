#!/bin/bash
# end-of-session verification: full GPU test suite, smoke(), bench (generator, tokenizer, reference arm), step sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
timeout 600 python bench.py --workload tokenizer > gpurun_out/final_bench_tokenizer.json 2> gpurun_out/final_bench_tokenizer.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
timeout 900 python tools/sweep.py > gpurun_out/final_sweep.jsonl 2> gpurun_out/final_sweep.err
tail -n 2 gpurun_out/final_pytest_gpu.log; tail -n 2 gpurun_out/final_smoke.log
