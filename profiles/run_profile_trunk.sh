#!/bin/bash
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 0 --batch 256 --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline --no-e2e --no-library-ref"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gemm2_bf16|attention_tc' -s 5 -c 5 -o gpurun_out/prof_trunk -f $BENCH > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -5
