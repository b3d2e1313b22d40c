#!/usr/bin/env python
"""Per-tile timeline of the CTA-pair GEMM (leader CTA of pair 0) from a -DGEMM_TRACE=1 build of the library
(MASKBIT_B200_LIB=tools/lib_trace.so).  For each of the four trunk GEMM shapes prints, per tile: the MMA issuer's wait for a
free accumulator stage, its total wait on `full` barriers (operands not yet landed), the mainloop issue span, when the
accumulator became ready for the epilogue, and how long the epilogue took.  All in SM clocks."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import _lib  # noqa: E402


def p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def main():
    L = _lib.lib()
    M = 512 * 257
    g = torch.Generator(device="cuda").manual_seed(0)
    tr = torch.zeros((4, 4, 32), dtype=torch.int64, device="cuda")
    _lib.check(L.mb_test_gemm_trace(p(tr)))
    for name, N, K, epi in (("qkv", 3072, 1024, 5), ("out", 1024, 1024, 7), ("up", 4096, 1024, 6), ("down", 1024, 4096, 7)):
        A = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
        W = (torch.randn((N, K), device="cuda", generator=g) * 0.03).to(torch.bfloat16)
        bias = torch.randn((N,), device="cuda", generator=g)
        vec2 = torch.randn((N,), device="cuda", generator=g)
        res = torch.randn((M, N), device="cuda", generator=g).to(torch.bfloat16) if epi == 7 else None
        stats = torch.rand((M, 8, 2), device="cuda", generator=g) + 1.0
        stats[:, :, 1] += 20.0
        sto = torch.empty((M, 8, 2), device="cuda") if epi == 7 else None
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
        for _ in range(3):
            tr.zero_()
            _lib.check(L.mb_test_gemm_ex(p(A), p(W), p(bias), p(vec2), p(res), p(stats), p(sto), p(out), M, N, K, epi, 0, 0,
                                         1.0 / 1024, 1e-12, None))
        torch.cuda.synchronize()
        t = tr.cpu()
        t0 = int(t[0, 0, 0])
        print(f"== {name}: N={N} K={K} epi={epi}   (ideal mainloop {K // 64 * 512} clk per tile)")
        print(" tile  mma_start  acc_wait  full_wait  issue_span | acc_ready  d(acc_ready)  epi_wait  epi_run | prod_empty_wait")
        prev = None
        for i in range(4, 24):
            ms, ma, fw, me = (int(t[0, e, i]) for e in range(4))
            e0, e1, e2 = (int(t[1, e, i]) for e in range(3))
            pw = int(t[2, 0, i])
            if ms == 0:
                break
            d = e1 - prev if prev is not None else 0
            prev = e1
            print(f" {i:4d} {ms - t0:10d} {ma - ms:9d} {fw:10d} {me - ma:11d} | {e1 - t0:9d} {d:13d} {e1 - e0:9d} {e2 - e1:8d} | {pw:8d}")


if __name__ == "__main__":
    main()
