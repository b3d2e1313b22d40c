#!/usr/bin/env python
"""Where does the attention kernel differ from fp64 torch?  Prints, per case, the number of wrong rows and their pattern
(sequence / head / row) -- debugging aid for kernel variants (MASKBIT_B200_LIB selects the library)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from maskbit_b200 import _lib  # noqa: E402


def p(t):
    return ctypes.c_void_p(t.data_ptr())


def run(n_seq, scale, S=257, D=1024, H=16):
    g = torch.Generator(device="cuda").manual_seed(S + n_seq)
    qkv = (torch.randn((n_seq * S, 3 * D), device="cuda", generator=g) * scale).to(torch.bfloat16)
    out = torch.empty((n_seq * S, D), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().mb_test_attention(p(qkv), p(out), n_seq, S, D, H, _lib.current_stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.double().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(n_seq, S, H, 64)
    o = out.double().view(n_seq, S, H, 64)
    err = (o - ref).abs()
    bound = ref.abs() * 2 ** -8 + 5.4e-3 * scale
    bad = (~(err <= bound)).any(-1)            # [n_seq, S, H] (NaN counts as bad)
    print(f"n_seq={n_seq} scale={scale}: bad rows {int(bad.sum())} of {bad.numel()}, NaN outputs {int(torch.isnan(o).sum())}, "
          f"worst err/bound {float((err / bound).nan_to_num(1e9).max()):.3g}")
    if bad.any():
        idx = bad.nonzero()
        items = torch.unique(idx[:, 0] * H + idx[:, 2])
        print("   items (seq*H+head) affected:", items[:24].tolist(), "..." if len(items) > 24 else "", f"({len(items)} items)")
        print("   local index of affected items (item // 148):", torch.unique(items // 148).tolist()[:20])
        rows = torch.unique(idx[:, 1])
        print("   rows affected:", rows[:40].tolist(), f"({len(rows)} distinct rows)")
        b = idx[0]
        print("   first bad row: seq", int(b[0]), "row", int(b[1]), "head", int(b[2]), "out", o[b[0], b[1], b[2], :4].tolist(), "ref", ref[b[0], b[1], b[2], :4].tolist())


for n_seq, scale in [(3, 1.5), (40, 1.5), (40, 4.0), (20, 10.0)]:
    run(n_seq, scale)
