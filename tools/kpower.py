#!/usr/bin/env python
"""Sustained (power-capped) rate and energy per launch of each trunk kernel at BASELINE config #2 size.

The 64-step job runs under the 1 kW cap (`sw_power_cap` active, SM clock ~1.4 GHz), where a kernel's in-situ duration is set
by the energy it burns rather than by its burst-clock duration.  Each kernel is therefore looped back to back for
`--seconds`, timed with CUDA events over the second half of the loop, while nvidia-smi samples power and SM clock:
    energy per launch [J] = mean power [W] x mean duration [s].
Output: one line per kernel (ms, TFLOP/s sustained, W, MHz, J per launch) and the per-forward energy budget."""
import argparse
import ctypes
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import _lib  # noqa: E402


def p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class Sampler:
    def __init__(self):
        self.rows, self.stop_flag = [], False
        self.proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=power.draw,clocks.sm", "--format=csv,noheader,nounits",
                                      "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        for line in self.proc.stdout:
            try:
                w, mhz = (float(x) for x in line.split(","))
                self.rows.append((time.perf_counter(), w, mhz))
            except ValueError:
                pass

    def window(self, t0, t1):
        r = [(w, m) for t, w, m in self.rows if t0 <= t <= t1]
        if not r:
            return float("nan"), float("nan")
        return sum(x[0] for x in r) / len(r), sum(x[1] for x in r) / len(r)

    def close(self):
        self.proc.terminate()


def sustain(fn, seconds, sampler):
    torch.cuda.synchronize()
    t_end = time.perf_counter() + seconds
    half = time.perf_counter() + seconds / 2
    ms_sum, n = 0.0, 0
    w0 = None
    while time.perf_counter() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if time.perf_counter() >= half:
            if w0 is None:
                w0 = time.perf_counter()
            else:
                ms_sum += e0.elapsed_time(e1)
                n += 20
    w1 = time.perf_counter()
    watts, mhz = sampler.window(w0, w1)
    return ms_sum / max(n, 1), watts, mhz


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=512)
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--shapes", default="", help="extra GEMM rows: name:N:K:epi,name:N:K:epi,...")
    a = ap.parse_args()
    L = _lib.lib()
    st = _lib.current_stream()
    M = a.seqs * 257
    g = torch.Generator(device="cuda").manual_seed(0)
    sampler = Sampler()
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
        ms, w, mhz = sustain(lambda: _lib.check(L.mb_test_gemm_ex(p(A), p(W), p(bias), p(vec2), p(res), p(stats), p(sto), p(out), M, N, K,
                                                                  epi, 0, 0, 1.0 / 1024, 1e-12, st)), a.seconds, sampler)
        rows.append((name, ms, 2.0 * M * N * K / ms / 1e9, w, mhz))

    gemm("gemm_qkv", 3072, 1024, 5)
    gemm("gemm_out", 1024, 1024, 7)
    gemm("gemm_up", 4096, 1024, 6)
    gemm("gemm_down", 1024, 4096, 7)
    for spec in filter(None, a.shapes.split(",")):
        nm, n_, k_, e_ = spec.split(":")
        gemm(nm, int(n_), int(k_), int(e_))
    if not a.only or a.only in "attention":
        qkv = torch.randn((M, 3072), device="cuda", generator=g).to(torch.bfloat16)
        out = torch.empty((M, 1024), dtype=torch.bfloat16, device="cuda")
        ms, w, mhz = sustain(lambda: _lib.check(L.mb_test_attention(p(qkv), p(out), a.seqs, 257, 1024, 16, st)), a.seconds, sampler)
        rows.append(("attention", ms, 4.0 * 257 * 257 * 64 * 16 * a.seqs / ms / 1e9, w, mhz))
    if not a.only or a.only in "cublas":
        # library comparison point: the same QKV-shaped product through torch.matmul (cuBLASLt), no epilogue
        A = torch.randn((M, 1024), device="cuda", generator=g).to(torch.bfloat16)
        W = torch.randn((3072, 1024), device="cuda", generator=g).to(torch.bfloat16)
        o = torch.empty((M, 3072), dtype=torch.bfloat16, device="cuda")
        ms, w, mhz = sustain(lambda: torch.matmul(A, W.t(), out=o), a.seconds, sampler)
        rows.append(("cublas_qkv_shape", ms, 2.0 * M * 3072 * 1024 / ms / 1e9, w, mhz))
        A4 = torch.randn((M, 4096), device="cuda", generator=g).to(torch.bfloat16)
        W4 = torch.randn((1024, 4096), device="cuda", generator=g).to(torch.bfloat16)
        o4 = torch.empty((M, 1024), dtype=torch.bfloat16, device="cuda")
        ms, w, mhz = sustain(lambda: torch.matmul(A4, W4.t(), out=o4), a.seconds, sampler)
        rows.append(("cublas_down_shape", ms, 2.0 * M * 1024 * 4096 / ms / 1e9, w, mhz))
    sampler.close()
    total = 0.0
    for name, ms, rate, w, mhz in rows:
        j = w * ms / 1e3
        if name in ("gemm_qkv", "gemm_out", "gemm_up", "gemm_down", "attention"):
            total += j
        print(f"{name:18s} {ms:8.4f} ms  {rate:8.1f} TFLOP/s  {w:7.1f} W  {mhz:6.0f} MHz  {j:7.4f} J/launch")
    print(f"sum over one layer (512 sequences): {total:.3f} J; x24 layers = {24 * total:.1f} J per 512-sequence forward")


if __name__ == "__main__":
    main()
