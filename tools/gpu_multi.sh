#!/bin/bash
# Multi-GPU measurements on N GPUs of one box (N = $1):  BASELINE configs[2] (14-bit, 128 images per GPU -> global batch 128 N),
# configs[4] (12-bit step-count sweep + small-batch latency), and -- at N = 2 -- the NCCL shard-parity test.
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_${N}gpu_devices.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -q -s > gpurun_out/${TAG}_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/${TAG}_pytest_multi.log
fi
timeout 900 $RUN --master-port 29521 bench.py --gpus $N --bits 14 --batch 128 --steps ${STEPS:-3} --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu_14bit_B128.json 2> gpurun_out/${TAG}_bench_${N}gpu_14bit_B128.err
tail -n 1 gpurun_out/${TAG}_bench_${N}gpu_14bit_B128.json | cut -c1-300
timeout 1200 $RUN --master-port 29533 tools/sweep.py --points ${POINTS:-256x8,256x16,256x32,256x64,256x256,1x64,8x64} > gpurun_out/${TAG}_sweep_12bit_${N}gpu.jsonl 2> gpurun_out/${TAG}_sweep_12bit_${N}gpu.err
cat gpurun_out/${TAG}_sweep_12bit_${N}gpu.jsonl
