#!/usr/bin/env python
"""BASELINE configs[4]: step-count sweep (8/16/32/64/256 decoding steps at B=256 per GPU) and small-batch latency (B=1, 8 at 64
steps) for the 12-bit generator, at 1 GPU or -- under torchrun -- N GPUs of one box (weak scaling: `batch` images PER GPU, one
all-gather of the finished uint8 images per call, like bench.py).

    python tools/sweep.py                                             # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/sweep.py

CUDA-event timing of sample() + decode (+ gather), barrier on both sides, MAX over ranks, 2 warm-up calls, median of 3 (1 for the
256-step points).  Rank 0 prints one JSON line per point:
    {"bits", "n_gpus", "batch_per_gpu", "steps", "ms", "images_per_s" (all GPUs), "ms_per_step"}."""
import argparse
import json
import os
import statistics
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import build_models, load_config, sample, sampler_kwargs  # noqa: E402
from maskbit_b200.sharding import gather_images, rank_seed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=12)
    ap.add_argument("--points", default="256x8,256x16,256x32,256x64,256x256,1x64,8x64,32x64")
    a = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        real = os.fdopen(os.dup(1), "w")      # NCCL prints its banner on fd 1
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    else:
        real = sys.stdout
    cfg = load_config(f"maskbit_generator_{a.bits}bit")
    kw = sampler_kwargs(cfg)
    tokenizer, gen = build_models(cfg, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for pt in a.points.split(","):
        b, t = (int(x) for x in pt.split("x"))
        labels = ((torch.arange(b) + rank * b) * 37 % 1000).to(dev)
        reps = 3 if t * b <= 64 * 256 else 1
        ts = []
        for i in range(2 + reps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            img, _ = sample(gen, tokenizer, num_samples=b, labels=labels, noise="device", seed=rank_seed(i, rank), return_trace=False,
                            skip_zero_scale_uncond=True, **dict(kw, num_steps=t))
            if world > 1:
                gather_images(tokenizer.postprocess_uint8(img), world * b)
            e1.record()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            ts.append(ms.item())
        ms = statistics.median(ts[2:])
        if rank == 0:
            print(json.dumps({"bits": a.bits, "n_gpus": world, "batch_per_gpu": b, "steps": t, "ms": round(ms, 3),
                              "images_per_s": round(world * b / ms * 1000, 3), "ms_per_step": round(ms / t, 4)}), file=real, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
