#!/usr/bin/env python
"""Times the individual kernels of the generator trunk at BASELINE config #2 size (512 sequences, M = 131 584 rows) through
the C-ABI test hooks, with CUDA events, L2 flushed between iterations.  Same-box A/B tool: run it before and after a kernel
change inside ONE gpurun call.  `--only attention` etc. restricts the set; `--iters` sets repetitions (median reported)."""
import argparse
import ctypes
import statistics
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import _lib  # noqa: E402


def p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def timeit(fn, iters, flush):
    ts = []
    for _ in range(iters + 2):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts[2:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    L = _lib.lib()
    st = _lib.current_stream()
    M = a.seqs * 257
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = []

    def gemm(name, N, K, epi):
        if a.only and a.only not in name:
            return
        A = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
        W = (torch.randn((N, K), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
        bias = torch.randn((N,), device="cuda", generator=g)
        vec2 = torch.randn((N,), device="cuda", generator=g)
        res = torch.randn((M, N), device="cuda", generator=g).to(torch.bfloat16) if epi == 7 else None
        stats = torch.rand((M, 8, 2), device="cuda", generator=g) + 1.0
        stats[:, :, 1] += 20.0
        sto = torch.empty((M, 8, 2), device="cuda") if epi in (7, 8) else None
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: _lib.check(L.mb_test_gemm_ex(p(A), p(W), p(bias), p(vec2), p(res), p(stats), p(sto), p(out), M, N, K, epi, 0, 0,
                                                         1.0 / 1024, 1e-12, st)), a.iters, flush)
        rows.append((name, ms, 2.0 * M * N * K / ms / 1e9, "TFLOP/s"))

    gemm("gemm_qkv  N3072 K1024 LN-in->bf16", 3072, 1024, 5)
    gemm("gemm_out  N1024 K1024 +LN(res)->bf16+stats", 1024, 1024, 7)
    gemm("gemm_up   N4096 K1024 LN-in+gelu->bf16", 4096, 1024, 6)
    gemm("gemm_down N1024 K4096 +LN(res)->bf16+stats", 1024, 4096, 7)
    if not a.only or a.only in "attention":
        qkv = torch.randn((M, 3072), device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty((M, 1024), dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: _lib.check(L.mb_test_attention(p(qkv), p(out), a.seqs, 257, 1024, 16, st)), a.iters, flush)
        rows.append(("attention S257 H16 d64", ms, 4.0 * 257 * 257 * 64 * 16 * a.seqs / ms / 1e9, "TFLOP/s"))
    for name, ms, rate, unit in rows:
        print(f"{name:42s} {ms:8.4f} ms  {rate:9.1f} {unit}")


if __name__ == "__main__":
    main()
