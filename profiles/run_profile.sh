#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list + one full capture of the top kernels of the sampling path.
# Outputs go to gpurun_out/ (scratch); summaries are copied into profiles/ with profiles/summarize.py.
# Numbers printed by a run under ncu are never bench values.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --batch ${PROF_BATCH:-256} --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline --no-e2e"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
# full capture: the 4 GEMMs + attention of one layer (after the first layer), then one decoder conv at 256^2 and the select kernel
ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16|attention_tc' -s 5 -c 5 \
    -o gpurun_out/prof_trunk -f $BENCH > gpurun_out/ncu_full.log 2>&1
# (gpurun brings back at most 64 MiB: keep the captures few and without source for the small kernels)
ncu --set full --clock-control none -k regex:'conv_tcgen05|select_step|act_split|embed_kernel|gn_partial' -s 60 -c 12 \
    -o gpurun_out/prof_rest -f $BENCH > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
