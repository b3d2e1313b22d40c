"""GPU unit tests of the individual sm_100a kernels through the C ABI test hooks (include/maskbit_b200.h).

Floating-point kernels are compared with a plain torch fp32 evaluation of the same op on the same (bf16-rounded)
inputs; tolerances are written next to each assert.
"""
import ctypes
import math

import pytest
import torch

from maskbit_b200 import _lib

pytestmark = pytest.mark.gpu


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return _lib.current_stream()


def _gemm(a, w, bias, residual, epi, seq_in=0, seq_out=0):
    M, K = a.shape
    N = w.shape[0]
    if epi in (0, 1):
        out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    elif epi == 3:
        out = torch.full(((M // seq_in) * seq_out, N), float("nan"), dtype=torch.float32, device="cuda")
    else:
        out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    _lib.check(_lib.lib().mb_test_gemm(_p(a), _p(w), _p(bias), _p(residual), _p(out), M, N, K, epi, seq_in, seq_out, _stream()))
    torch.cuda.synchronize()
    return out


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (257, 128, 1024), (771, 1024, 1024), (1028, 3072, 1024), (514, 1024, 4096),
                                   (2056, 4096, 1024), (131584 // 8, 256, 1024)])
@pytest.mark.parametrize("epi", [0, 1, 2, 4])
def test_gemm_epilogues(M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K + epi)
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((N, K), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn((N,), device="cuda", generator=g)
    res = torch.randn((M, N), device="cuda", generator=g).to(torch.bfloat16) if epi == 2 else None
    out = _gemm(a, w, bias, res, epi)
    ref = a.double() @ w.double().t() + bias.double()
    if epi in (1, 4):
        ref = _gelu(ref)
    if epi == 2:
        ref = ref + res.double()
    err = (out.double() - ref).abs()
    scale = ref.abs().max().item()
    if out.dtype == torch.bfloat16:
        # bf16 output rounding: 2^-9 relative on each element, plus fp32 accumulation noise
        assert (err <= ref.abs() * 2 ** -8 + 1e-3 * scale).all(), f"max err {err.max().item()} (scale {scale})"
    else:
        # fp32 accumulate over K <= 4096 bf16 products: 1e-5 relative to the output scale
        assert err.max().item() <= 2e-5 * scale * math.sqrt(K / 64), f"max err {err.max().item()} (scale {scale})"


def test_gemm_class_row_drop():
    """EPI 3: prediction layer writes fp32 logits with the class-token row of every sequence removed (bert.py:503)."""
    S, n_seq, N, K = 257, 5, 128, 1024
    g = torch.Generator(device="cuda").manual_seed(3)
    a = torch.randn((n_seq * S, K), device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn((N, K), device="cuda", generator=g) * 0.05).to(torch.bfloat16)
    bias = torch.randn((N,), device="cuda", generator=g)
    out = _gemm(a, w, bias, None, 3, seq_in=S, seq_out=S - 1)
    ref = (a.double() @ w.double().t() + bias.double()).view(n_seq, S, N)[:, : S - 1].reshape(-1, N)
    assert out.shape == ref.shape
    assert not torch.isnan(out).any()
    assert (out.double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() * 4


def test_gemm_rejects_bad_shapes():
    a = torch.zeros((128, 96), dtype=torch.bfloat16, device="cuda")
    w = torch.zeros((64, 96), dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros((64,), device="cuda")
    out = torch.zeros((128, 64), dtype=torch.bfloat16, device="cuda")
    rc = _lib.lib().mb_test_gemm(_p(a), _p(w), _p(bias), None, _p(out), 128, 64, 96, 0, 0, 0, _stream())
    assert rc == -1 and b"gemm shape" in _lib.lib().mb_last_error()


@pytest.mark.parametrize("n_seq,S", [(1, 257), (3, 257), (40, 257), (2, 256), (2, 65), (1, 272)])
def test_attention(n_seq, S):
    D, H = 1024, 16
    g = torch.Generator(device="cuda").manual_seed(S + n_seq)
    qkv = (torch.randn((n_seq * S, 3 * D), device="cuda", generator=g) * 1.5).to(torch.bfloat16)
    out = torch.empty((n_seq * S, D), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().mb_test_attention(_p(qkv), _p(out), n_seq, S, D, H, _stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.double().view(n_seq, S, 3, H, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) / 8.0, dim=-1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(n_seq * S, D)
    err = (out.double() - ref).abs()
    # output rounded to bf16 (2^-9 relative, |o| up to ~6) + P rounded to bf16 before P.V (2^-9 relative on |v| ~ 1.5)
    assert (err <= ref.abs() * 2 ** -8 + 8e-3).all(), f"attention max err {err.max().item()}"


@pytest.mark.parametrize("rows", [1, 7, 257, 1028])
def test_layernorm(rows):
    g = torch.Generator(device="cuda").manual_seed(rows)
    x = torch.randn((rows, 1024), device="cuda", generator=g) * 3 + 0.5
    gamma = torch.randn((1024,), device="cuda", generator=g)
    beta = torch.randn((1024,), device="cuda", generator=g)
    out = torch.empty((rows, 1024), dtype=torch.bfloat16, device="cuda")
    _lib.check(_lib.lib().mb_test_layernorm(_p(x), _p(gamma), _p(beta), 1e-12, _p(out), rows, 1024, _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x.double(), (1024,), gamma.double(), beta.double(), eps=1e-12)
    err = (out.double() - ref).abs()
    assert (err <= ref.abs() * 2 ** -8 + 1e-5).all(), f"layernorm max err {err.max().item()}"
