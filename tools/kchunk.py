#!/usr/bin/env python
"""Experiment: does running the MLP (up-projection + GELU -> down-projection) in row chunks whose hidden activation stays in the
126 MB L2 (one reused chunk-sized buffer: dirty lines are overwritten before they are evicted) lower the energy per layer?

Sustained loops under the power cap, like tools/kpower.py:  `full` = the two GEMMs over all 131 584 rows (hidden [M,4096] bf16 =
1.08 GB written to and re-read from HBM);  `chunk N` = for every chunk of N 256-row tiles: up-GEMM into the reused buffer, then
down-GEMM out of it.  Prints ms and J per MLP pass for each schedule."""
import ctypes
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import _lib  # noqa: E402
from tools.kpower import Sampler  # noqa: E402


def vp(t, byte_off=0):
    return ctypes.c_void_p(t.data_ptr() + byte_off) if t is not None else None


def main():
    L = _lib.lib()
    st = _lib.current_stream()
    M, D, H = 512 * 257, 1024, 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    y = torch.randn((M, D), device="cuda", generator=g).to(torch.bfloat16)
    w1 = (torch.randn((H, D), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    w2 = (torch.randn((D, H), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
    b1, u1 = torch.randn((H,), device="cuda", generator=g), torch.randn((H,), device="cuda", generator=g)
    b2, g2 = torch.randn((D,), device="cuda", generator=g), torch.randn((D,), device="cuda", generator=g)
    stats = torch.rand((M, 8, 2), device="cuda", generator=g) + 1.0
    stats[:, :, 1] += 20.0
    sto = torch.empty((M, 8, 2), device="cuda")
    out = torch.empty((M, D), dtype=torch.bfloat16, device="cuda")
    hidden = torch.empty((M, H), dtype=torch.bfloat16, device="cuda")

    def mlp(rows0, rows, hid, hid_off_rows):
        # up: A = y[rows0:rows0+rows], LN-in + GELU -> hid ; down: A = hid, residual y rows, -> out rows
        _lib.check(L.mb_test_gemm_ex(vp(y, rows0 * D * 2), vp(w1), vp(b1), vp(u1), None, vp(stats, rows0 * 64), None,
                                     vp(hid, hid_off_rows * H * 2), rows, H, D, 6, 0, 0, 1.0 / 1024, 1e-12, st))
        _lib.check(L.mb_test_gemm_ex(vp(hid, hid_off_rows * H * 2), vp(w2), vp(b2), vp(g2), vp(y, rows0 * D * 2), vp(stats, rows0 * 64),
                                     vp(sto, rows0 * 64), vp(out, rows0 * D * 2), rows, D, H, 7, 0, 0, 1.0 / 1024, 1e-12, st))

    def run_full():
        mlp(0, M, hidden, 0)

    def make_chunked(tiles):
        rows_c = tiles * 256

        def run():
            r = 0
            while r < M:
                n = min(rows_c, M - r)
                mlp(r, n, hidden, 0)          # every chunk reuses the first rows_c rows of `hidden`
                r += n
        return run

    sampler = Sampler()
    results = []
    for name, fn in [("full", run_full)] + [(f"chunk {t} tiles ({t * 256 * H * 2 / 1e6:.0f} MB hidden)", make_chunked(t)) for t in (74, 37, 19, 10)] + [("full again", run_full)]:
        torch.cuda.synchronize()
        t_end = time.perf_counter() + 3.0
        half = time.perf_counter() + 1.5
        ms, n, w0 = 0.0, 0, None
        while time.perf_counter() < t_end:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                fn()
            e1.record()
            torch.cuda.synchronize()
            if time.perf_counter() >= half:
                if w0 is None:
                    w0 = time.perf_counter()
                else:
                    ms += e0.elapsed_time(e1); n += 4
        watts, mhz = sampler.window(w0, time.perf_counter())
        results.append((name, ms / max(n, 1), watts, mhz))
    sampler.close()
    for name, ms, w, mhz in results:
        print(f"{name:40s} {ms:8.4f} ms per MLP pass  {w:7.1f} W  {mhz:6.0f} MHz  {w * ms / 1e3:7.4f} J")


if __name__ == "__main__":
    main()
