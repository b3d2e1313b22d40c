#!/usr/bin/env python
"""Where a batch-1 GEMM launch spends its time (M = 514 rows: 2 sequences).  Needs a -DGEMM_TRACE=1 build
(MASKBIT_B200_LIB=tools/lib_trace.so).  For the four trunk GEMM shapes: CUDA-event time per launch over a back-to-back run
(launch overhead included), and the in-kernel timeline of the leader CTA of pair 0 in SM clocks: entry -> set up (barriers, TMEM,
cluster sync) -> predecessor complete (griddepcontrol.wait) -> first operands landed -> accumulator ready -> epilogue done -> exit."""
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
    M = 2 * 257
    st = _lib.current_stream()
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

        def launch():
            _lib.check(L.mb_test_gemm_ex(p(A), p(W), p(bias), p(vec2), p(res), p(stats), p(sto), p(out), M, N, K, epi, 0, 0,
                                         1.0 / 1024, 1e-12, st))
        for _ in range(5):
            launch()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 50
        e0.record()
        for _ in range(n):
            launch()
        e1.record()
        torch.cuda.synchronize()
        tr.zero_()
        launch()
        torch.cuda.synchronize()
        t = tr.cpu()
        t0 = int(t[3, 0, 0])
        life = [int(t[3, e, 0]) - t0 for e in range(4)]
        ms, ma, fw, me = (int(t[0, e, 0]) for e in range(4))       # MMA issuer, tile 0: start | accumulator free | full-wait total | issued
        ep0, ep1, ep2 = (int(t[1, e, 0]) for e in range(3))        # epilogue warp 0, tile 0: waiting | accumulator ready | done
        print(f"== {name}: M={M} N={N} K={K} epi={epi}: {1e3 * e0.elapsed_time(e1) / n:7.2f} us per launch back to back "
              f"({K // 64} k-blocks, ideal mainloop {K // 64 * 512} clk)")
        print(f"   clk from entry: set up {life[1]}, predecessor complete {life[2]}, MMA issue start {ms - t0}, all MMAs issued {me - t0} "
              f"(waited {fw} on operands), accumulator ready {ep1 - t0}, epilogue done {ep2 - t0}, exit {life[3]}")


if __name__ == "__main__":
    main()
