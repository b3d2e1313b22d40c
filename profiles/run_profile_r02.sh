#!/bin/bash
# Round-2 evidence run (one gpurun call): default bench line, ncu launch list, ncu --set full capture of the trunk kernels
# (one layer) and of the decoder conv + act_split, tokenizer bench (BASELINE configs[3]), sustained power / energy per launch
# of the round-1 library (tools/lib_old.so, built from commit 7714de4) vs the current one.
mkdir -p gpurun_out
TAG=${TAG:-r02}
timeout 900 python bench.py > gpurun_out/${TAG}_bench_B256_T64.json 2> gpurun_out/${TAG}_bench_B256_T64.err; cut -c1-600 gpurun_out/${TAG}_bench_B256_T64.json
BENCH="python bench.py --steps 1 --warmup 0 --batch 256 --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline --no-e2e --no-library-ref --profile-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16|attention_tc' -s 5 -c 5 -o gpurun_out/${TAG}_prof_trunk -f $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'conv_tcgen05|act_split' -s 60 -c 4 -o gpurun_out/${TAG}_prof_conv -f $BENCH > gpurun_out/${TAG}_ncu_conv.log 2>&1
echo "== tokenizer bench"; timeout 600 python bench.py --workload tokenizer --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_tokenizer_B512.json 2> gpurun_out/${TAG}_bench_tokenizer_B512.err; cut -c1-400 gpurun_out/${TAG}_bench_tokenizer_B512.json
echo "== kpower new"; timeout 300 python tools/kpower.py --seconds 2 2>&1 | tee gpurun_out/${TAG}_kpower_new.txt
[ -f tools/lib_old.so ] && { echo "== kpower old (round-1 library)"; MASKBIT_B200_LIB=tools/lib_old.so timeout 300 python tools/kpower.py --seconds 2 2>&1 | tee gpurun_out/${TAG}_kpower_old.txt; }
ls -la gpurun_out | grep ${TAG}_ | tail -14
