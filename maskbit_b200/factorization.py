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
    bit_shift = int(math.log2(codebook_size)) // splits
    bit_mask = (1 << bit_shift) - 1
    return torch.stack([(tokens & (bit_mask << (i * bit_shift))) >> (i * bit_shift) for i in range(splits)], dim=2)
