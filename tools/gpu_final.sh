#!/bin/bash
# end-of-session verification: full GPU test suite (-s: measured values in the log), smoke(), bench (generator, tokenizer,
# reference arm), step sweep.  TAG names the output files under gpurun_out/.
mkdir -p gpurun_out
TAG=${TAG:-final}
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench_B256_T64.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --workload tokenizer > gpurun_out/${TAG}_bench_tokenizer_B512.json 2> gpurun_out/${TAG}_bench_tokenizer.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
timeout 900 python tools/sweep.py > gpurun_out/${TAG}_sweep_12bit_1gpu.jsonl 2> gpurun_out/${TAG}_sweep.err
tail -n 2 gpurun_out/${TAG}_pytest_gpu.log; tail -n 2 gpurun_out/${TAG}_smoke.log; cut -c1-200 gpurun_out/${TAG}_bench_B256_T64.json; cut -c1-120 gpurun_out/${TAG}_bench_tokenizer_B512.json; cat gpurun_out/${TAG}_sweep_12bit_1gpu.jsonl | cut -c1-200
