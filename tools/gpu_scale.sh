#!/bin/bash
# bench.py under torchrun at N GPUs of one box (N = $1), 12-bit config #2
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale${N}_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps ${STEPS:-2} --warmup 3 > gpurun_out/scale${N}_12bit.json 2> gpurun_out/scale${N}_12bit.err
tail -n 1 gpurun_out/scale${N}_12bit.json | cut -c1-200
