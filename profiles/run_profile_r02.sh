#!/bin/bash
# Round-2 evidence run (one gpurun call): full GPU test suite, smoke, default bench line, ncu launch list, ncu --set full capture of
# the trunk kernels (one layer) and of every conv launch of one 32-image tokenizer pass (-> profiles/ncu_traffic.json), tokenizer
# bench (BASELINE configs[3]), small-batch points of the sweep, sustained power / energy per launch.
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 1 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_B256_T64.json 2> gpurun_out/${TAG}_bench_B256_T64.err; cut -c1-200 gpurun_out/${TAG}_bench_B256_T64.json
BENCH="python bench.py --steps 1 --warmup 0 --batch 256 --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline --no-e2e --no-library-ref --profile-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16|attention_tc' -s 5 -c 5 -o gpurun_out/${TAG}_prof_trunk -f $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'conv_tcgen05' -c 80 -o gpurun_out/${TAG}_prof_conv -f python bench.py --workload tokenizer --batch 32 --steps 1 --warmup 0 --profile-steps 0 > gpurun_out/${TAG}_ncu_conv.log 2>&1
echo "== tokenizer bench"; timeout 600 python bench.py --workload tokenizer --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_tokenizer_B512.json 2> gpurun_out/${TAG}_bench_tokenizer_B512.err; cut -c1-200 gpurun_out/${TAG}_bench_tokenizer_B512.json
echo "== small-batch sweep points"; timeout 600 python tools/sweep.py --points 1x64,2x64,4x64,8x64,32x64 > gpurun_out/${TAG}_sweep_small_1gpu.jsonl 2> gpurun_out/${TAG}_sweep_small.err; cat gpurun_out/${TAG}_sweep_small_1gpu.jsonl
echo "== kpower"; timeout 300 python tools/kpower.py --seconds 2 2>&1 | tee gpurun_out/${TAG}_kpower.txt
ls -la gpurun_out | grep ${TAG}_ | tail -16
