#!/usr/bin/env python
"""BASELINE config #5: step-count sweep (8/16/32/64/256 decoding steps at B=256) and small-batch latency (B=1, 8 at 64 steps)
for the 12-bit generator on one GPU.  CUDA-event timing of sample() + decode, 2 warm-up calls, median of 3.
Prints one JSON line per point:  {"bits", "batch", "steps", "ms", "images_per_s", "ms_per_step"}."""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import build_models, load_config, sample, sampler_kwargs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bits", type=int, default=12)
    ap.add_argument("--points", default="256x8,256x16,256x32,256x64,256x256,1x64,8x64,32x64")
    a = ap.parse_args()
    cfg = load_config(f"maskbit_generator_{a.bits}bit")
    kw = sampler_kwargs(cfg)
    tokenizer, gen = build_models(cfg, device="cuda")
    for pt in a.points.split(","):
        b, t = (int(x) for x in pt.split("x"))
        labels = (torch.arange(b) * 37 % 1000).cuda()
        ts = []
        for i in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sample(gen, tokenizer, num_samples=b, labels=labels, noise="device", seed=i, return_trace=False,
                   skip_zero_scale_uncond=True, **dict(kw, num_steps=t))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts[2:])
        print(json.dumps({"bits": a.bits, "batch": b, "steps": t, "ms": round(ms, 3), "images_per_s": round(b / ms * 1000, 3),
                          "ms_per_step": round(ms / t, 4)}), flush=True)


if __name__ == "__main__":
    main()
