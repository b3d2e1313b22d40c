"""ORACLE -- test infrastructure.  ctypes loader for the plain-C select oracle (select_oracle.c)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_build", "libselect_oracle.so")
        if not os.path.isfile(path):
            build()
        L = ctypes.CDLL(path)
        L.mbo_expf.restype = ctypes.c_float
        L.mbo_expf.argtypes = [ctypes.c_float]
        L.mbo_logf.restype = ctypes.c_float
        L.mbo_logf.argtypes = [ctypes.c_float]
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int64)
        L.mbo_select_step.restype = ctypes.c_int
        L.mbo_select_step.argtypes = [fp, fp, ctypes.c_float, ctypes.c_float, fp, fp, ctypes.c_float, ctypes.c_float,
                                      ctypes.c_float, ip, ip, ip, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                      ctypes.c_int, ctypes.c_int64]
        _LIB = L
    return _LIB


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def select_step(logits_c, logits_u, scale, temperature, q, gumbel, randomize_temperature, one_minus_progress, mask_len,
                tokens_in, mask_token, seq_stride=None):
    """numpy in / numpy out wrapper.  logits_* [B, seq_stride, m, V]; q [B*n*m, V]; gumbel, tokens_in [B, n, m]."""
    tokens_in = np.ascontiguousarray(tokens_in, dtype=np.int64)
    B, n, m = tokens_in.shape
    V = q.shape[-1]
    if seq_stride is None:
        seq_stride = logits_c.shape[1]
    lc, lcp = _f(logits_c)
    if logits_u is not None:
        lu, lup = _f(logits_u)
    else:
        lup = None
    qq, qp = _f(q)
    gg, gp = _f(gumbel)
    pred = np.empty_like(tokens_in)
    out = np.empty_like(tokens_in)
    ip = ctypes.POINTER(ctypes.c_int64)
    k = lib().mbo_select_step(lcp, lup, float(np.float32(scale)), float(np.float32(temperature)), qp, gp,
                              float(np.float32(randomize_temperature)), float(np.float32(one_minus_progress)),
                              float(np.float32(mask_len)), tokens_in.ctypes.data_as(ip), pred.ctypes.data_as(ip),
                              out.ctypes.data_as(ip), B, n, m, V, seq_stride, mask_token)
    return pred, out, k
