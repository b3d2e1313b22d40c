#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/scale2_gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/scale2_12bit.json 2> gpurun_out/scale2_12bit.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --bits 14 --batch 128 > gpurun_out/scale2_14bit.json 2> gpurun_out/scale2_14bit.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/scale2_ref.json 2> gpurun_out/scale2_ref.err
tail -3 gpurun_out/scale2_12bit.err
