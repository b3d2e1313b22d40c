"""MLMLoss -- drop-in mirror of modeling/modules/losses.py:289-339, forward only (no autograd graph: the backward half of the
training step is outside this round's scope).  The label-smoothed cross entropy and the accuracy counters are one
pass over the logits in libmaskbit_b200 (mb_mlm_loss), deterministic."""
import ctypes
from typing import Mapping, Text, Tuple

import torch

from . import _lib


class MLMLoss:
    def __init__(self, label_smoothing: float = 0.1, sum_splits: bool = False):
        self.label_smoothing = float(label_smoothing)
        self.sum_splits = bool(sum_splits)

    def __call__(self, inputs, targets, masks):
        return self.forward(inputs, targets, masks)

    @torch.no_grad()
    def forward(self, inputs: torch.Tensor, targets: torch.Tensor, masks: torch.Tensor) -> Tuple[torch.Tensor, Mapping[Text, torch.Tensor]]:
        """inputs fp32 [b, n, m, V] logits, targets int64 [b, n, m], masks bool [b, n, m] -> (loss, loss_dict) like the reference:
        mlm_loss, correct_tokens, masked_token_loss, masked_correct_tokens (0-d tensors on the logits' device)."""
        if inputs.device.type != "cuda":
            raise _lib.MaskbitError("MLMLoss runs on a CUDA device; there is no CPU fallback")
        b, n, m, v = inputs.shape
        if tuple(targets.shape) != (b, n, m) or tuple(masks.shape) != (b, n, m):
            raise ValueError(f"targets / masks must have shape {(b, n, m)}")
        dev = inputs.device
        x = inputs.detach().to(torch.float32).contiguous()
        t = targets.to(device=dev, dtype=torch.int64).contiguous()
        mk = masks.to(device=dev, dtype=torch.bool).contiguous()
        if int(t.min()) < 0 or int(t.max()) >= v:
            raise ValueError("target token outside the vocabulary")
        L = _lib.lib()
        scratch = torch.empty(L.mb_mlm_loss_scratch_bytes(), dtype=torch.uint8, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(L.mb_mlm_loss(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(t.data_ptr()), ctypes.c_void_p(mk.data_ptr()),
                                     b * n * m, v, m, self.label_smoothing, int(self.sum_splits), ctypes.c_void_p(scratch.data_ptr()),
                                     ctypes.c_void_p(out.data_ptr()), _lib.current_stream()))
        loss_dict = {"mlm_loss": out[0], "correct_tokens": out[1], "masked_token_loss": out[2], "masked_correct_tokens": out[3]}
        return out[0], loss_dict
