#!/bin/bash
# Run ON THE GPU BOX (via gpurun): ncu launch list + one full capture of the top kernels of the sampling path.
# Outputs go to gpurun_out/ (scratch); summaries are copied into profiles/ by hand after reading them.
# Numbers printed by a run under ncu are never bench values.
set -x
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --batch ${PROF_BATCH:-256} --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline"
# every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
# full capture: GEMMs of one layer (qkv, out, up, down) + attention + layernorm, after the first layer
ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16|attention_kernel|layernorm_kernel' -s 8 -c 7 \
    -o gpurun_out/prof_trunk -f $BENCH > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
