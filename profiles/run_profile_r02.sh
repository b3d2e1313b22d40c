#!/bin/bash
# Round-2 evidence run (one gpurun call): sustained power / energy per launch old vs new, ncu launch list, ncu --set full capture
# of the trunk kernels (one layer) and of the decoder conv + act_split, tokenizer bench (BASELINE configs[3]).
mkdir -p gpurun_out
echo "== kpower old (round-1 library)"; MASKBIT_B200_LIB=tools/lib_old.so python tools/kpower.py --seconds 3 2>&1 | tee gpurun_out/r02_kpower_old.txt
echo "== kpower new"; python tools/kpower.py --seconds 3 2>&1 | tee gpurun_out/r02_kpower_new.txt
BENCH="python bench.py --steps 1 --warmup 0 --batch 256 --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline --no-e2e --no-library-ref --profile-steps 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/r02_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16|attention_tc' -s 5 -c 5 -o gpurun_out/r02_prof_trunk -f $BENCH > gpurun_out/r02_ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:'conv_tcgen05|act_split' -s 60 -c 4 -o gpurun_out/r02_prof_conv -f $BENCH > gpurun_out/r02_ncu_conv.log 2>&1
echo "== tokenizer bench"; python bench.py --workload tokenizer --steps 3 --warmup 3 > gpurun_out/r02_bench_tokenizer_B512.json 2> gpurun_out/r02_bench_tokenizer_B512.err; cat gpurun_out/r02_bench_tokenizer_B512.json | cut -c1-400
ls -la gpurun_out | grep r02_ | tail -12
