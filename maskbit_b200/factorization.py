"""Token factorisation helpers (modeling/modules/factorization.py:7-46) on device tensors."""
import ctypes
import math

import torch

from . import _lib


def combine_factorized_tokens(tokens: torch.Tensor, codebook_size: int, splits: int) -> torch.Tensor:
    """[B, n, m] group tokens -> [B, n] full index.  Like the reference (factorization.py:19) the result is float32."""
    bit_shift = int(math.log2(codebook_size)) // splits
    combined = torch.zeros((tokens.shape[0], tokens.shape[1]), device=tokens.device)
    for i in range(splits):
        combined += (tokens[..., i] << (i * bit_shift))
    return combined


def split_factorized_tokens(tokens: torch.Tensor, codebook_size: int, splits: int) -> torch.Tensor:
    """[B, n] full index -> [B, n, splits] group tokens (factorization.py:27-46), mb_split_tokens on CUDA tensors."""
    bit_shift = int(math.log2(codebook_size)) // splits
    if tokens.device.type != "cuda":
        raise _lib.MaskbitError("split_factorized_tokens runs on a CUDA device; there is no CPU fallback")
    tok = tokens.to(torch.int64).contiguous()
    out = torch.empty(tuple(tok.shape) + (splits,), dtype=torch.int64, device=tok.device)
    if tok.numel():
        with torch.cuda.device(tok.device):
            _lib.check(_lib.lib().mb_split_tokens(ctypes.c_void_p(tok.data_ptr()), tok.numel(), splits, bit_shift,
                                                  ctypes.c_void_p(out.data_ptr()), _lib.current_stream()))
    return out
