#!/usr/bin/env python
"""Small-batch latency probe: wall / CUDA-event time of sample() at B = 1, 8 next to the sum of its kernels' own durations
(per-class CUDA events, mb_profile_*): the difference is launch gaps."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import build_models, load_config, sample, sampler_kwargs  # noqa: E402

cfg = load_config("maskbit_generator_12bit")
kw = sampler_kwargs(cfg)
tok, gen = build_models(cfg, device="cuda")
for B in (1, 8):
    labels = (torch.arange(B) * 37 % 1000).cuda()
    for prof in (False, True):
        for i in range(3):
            if prof and i == 2:
                gen.profile_enable(True); tok.profile_enable(True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.perf_counter()
            e0.record()
            sample(gen, tok, num_samples=B, labels=labels, noise="device", seed=i, return_trace=False, skip_zero_scale_uncond=True, **kw)
            w_enq = time.perf_counter() - w0
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        if prof:
            p = dict(gen.profile_read()); p.update(tok.profile_read())
            gen.profile_enable(False); tok.profile_enable(False)
            ksum = sum(v[0] for v in p.values()); n = sum(v[1] for v in p.values())
            print(f"B={B} profiled: {ms:.1f} ms total, kernels sum {ksum:.1f} ms over {n} launches ({1000 * ksum / n:.1f} us avg), host enqueue {1000 * w_enq:.1f} ms")
            print("   " + "  ".join(f"{k} {1000 * v[0] / v[1]:.1f}us x{v[1]}" for k, v in sorted(p.items(), key=lambda kv: -kv[1][0]) if v[1]))
        else:
            print(f"B={B}: {ms:.1f} ms total (64 steps + decode), host enqueue {1000 * w_enq:.1f} ms")
