#!/bin/bash
# Evidence for the final code of round 2, small outputs only (gpurun brings back <= 64 MiB: the .ncu-rep files stay on the box,
# their raw pages are exported as CSV): pytest -m gpu -s log, smoke, ncu launch list, raw pages of the trunk capture and of every conv
# launch of one 32-image tokenizer pass, tokenizer bench line.
mkdir -p gpurun_out /tmp/ncu
TAG=${TAG:-r02e}
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -n 1 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -n 1 gpurun_out/${TAG}_smoke.log
BENCH="python bench.py --steps 1 --warmup 0 --batch 256 --sampling-steps 1 --skip-dead-uncond 0 --no-cpu-baseline --no-e2e --no-library-ref --profile-steps 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'gemm2_bf16|attention_tc' -s 5 -c 5 -o /tmp/ncu/trunk -f $BENCH > gpurun_out/${TAG}_ncu_full.log 2>&1
ncu -i /tmp/ncu/trunk.ncu-rep --page raw --csv > gpurun_out/${TAG}_trunk_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'conv_tcgen05' -c 80 -o /tmp/ncu/conv -f python bench.py --workload tokenizer --batch 32 --steps 1 --warmup 0 --profile-steps 0 > gpurun_out/${TAG}_ncu_conv.log 2>&1
ncu -i /tmp/ncu/conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_raw.csv 2>/dev/null
timeout 600 python bench.py --workload tokenizer --steps 3 --warmup 3 > gpurun_out/${TAG}_bench_tokenizer_B512.json 2> gpurun_out/${TAG}_bench_tokenizer_B512.err; cut -c1-200 gpurun_out/${TAG}_bench_tokenizer_B512.json
du -sh gpurun_out; ls -la gpurun_out | grep ${TAG}_
