"""LFQBert -- drop-in mirror of modeling/bert.py:344-508 (constructor arguments, attributes, forward signature,
state_dict key layout) whose forward runs in libmaskbit_b200's sm_100a kernels."""
import ctypes
import math

import torch

from . import _lib
from .base_model import EngineModel
from .weights import bert_spec, lfq_bert_spec, synthetic_bert_state_dict, synthetic_lfq_bert_state_dict


class LFQBert(EngineModel):
    _model_id = _lib.MB_GENERATOR
    _generator_cls = 0   # mb_config.generator_cls: 0 = LFQBert (bit-token input projection), 1 = Bert (embedding tables)

    def __init__(self, img_size=256, hidden_dim=768, codebook_size=1024, codebook_splits=1, depth=24, heads=8,
                 mlp_dim=3072, dropout=0.1, nclass=1000, input_stride: int = 16, use_prenorm: bool = False):
        super().__init__()
        self.nclass = nclass
        self.drop_label = nclass
        self.seq_len = (img_size // input_stride) ** 2
        self.splits = codebook_splits
        self.bits = int(math.log2(codebook_size))
        effective_bits = self.bits // self.splits
        self.effective_codebook_size = int(2 ** effective_bits)
        self.mask_token = self.effective_codebook_size
        self.hidden_dim, self.depth, self.heads, self.mlp_dim = hidden_dim, depth, heads, mlp_dim
        self.codebook_size = codebook_size
        self.use_prenorm = use_prenorm
        self.dropout = dropout  # inference only: dropout is the identity in eval mode
        # tokenizer-side fields of mb_config are unused by a generator handle
        self._dec = dict(hidden_channels=128, channel_mult=(1, 1, 2, 2, 4), num_resolutions=5, num_res_blocks=2)

    def get_group_splits(self) -> int:
        return self.splits

    def _arch(self):
        return dict(hidden_dim=self.hidden_dim, codebook_size=self.codebook_size, codebook_splits=self.splits,
                    depth=self.depth, mlp_dim=self.mlp_dim, nclass=self.nclass, seq_len=self.seq_len,
                    use_prenorm=bool(self.use_prenorm))

    def _expected_spec(self):
        return [(n, s) for n, s, _ in lfq_bert_spec(**self._arch())]

    def _default_state_dict(self):
        return synthetic_lfq_bert_state_dict(seed=0, **self._arch())

    def _mb_config(self):
        c = _lib.MBConfig()
        c.hidden_dim, c.depth, c.heads, c.mlp_dim = self.hidden_dim, self.depth, self.heads, self.mlp_dim
        c.token_bits, c.codebook_splits, c.nclass, c.seq_len = self.bits, self.splits, self.nclass, self.seq_len
        c.use_prenorm = int(bool(self.use_prenorm))
        c.generator_cls = self._generator_cls
        c.dec_hidden_channels = self._dec["hidden_channels"]
        for i, v in enumerate(self._dec["channel_mult"]):
            c.dec_channel_mult[i] = v
        c.dec_num_resolutions, c.dec_num_res_blocks, c.num_channels = self._dec["num_resolutions"], self._dec["num_res_blocks"], 3
        return c

    @torch.no_grad()
    def forward(self, img_tokens, class_labels, drop_label_mask=None, return_attn=False):
        """bert.py:456-508.  img_tokens int64 [N, seq_len, splits], class_labels int64 [N], drop_label_mask bool [N]
        or None (= drop every label, the reference's behaviour for None).  Returns fp32 [N, seq_len, splits, V]; with
        return_attn=True (bert.py:505-506) the pair (logits, [depth tensors fp32 [N, seq_len+1, seq_len+1]]): per layer the
        head-averaged attention weights of nn.MultiheadAttention(need_weights=True), computed by a side kernel per layer.
        Unlike the reference (bert.py:484) the caller's class_labels tensor is not modified."""
        h = self._engine()
        dev = self._device
        tok = img_tokens.to(device=dev, dtype=torch.int64).contiguous()
        lab = class_labels.to(device=dev, dtype=torch.int64).contiguous().view(-1)
        n = tok.shape[0]
        if tok.dim() != 3 or tok.shape[1] != self.seq_len or tok.shape[2] != self.splits or lab.numel() != n:
            raise ValueError(f"expected img_tokens [N,{self.seq_len},{self.splits}] and class_labels [N], got {tuple(tok.shape)}, {tuple(lab.shape)}")
        drop_ptr = None
        if drop_label_mask is not None:
            drop = drop_label_mask.to(device=dev).to(torch.uint8).contiguous().view(-1)
            if drop.numel() != n:
                raise ValueError("drop_label_mask must have one entry per sequence")
            drop_ptr = ctypes.c_void_p(drop.data_ptr())
        with torch.cuda.device(dev):
            logits = torch.empty((n, self.seq_len, self.splits, self.effective_codebook_size), dtype=torch.float32, device=dev)
            if return_attn:
                s = self.seq_len + 1
                attn = torch.empty((self.depth, n, s, s), dtype=torch.float32, device=dev)
                _lib.check(_lib.lib().mb_generator_forward_attn(h, ctypes.c_void_p(tok.data_ptr()), n, ctypes.c_void_p(lab.data_ptr()), n,
                                                                drop_ptr, n, ctypes.c_void_p(logits.data_ptr()),
                                                                ctypes.c_void_p(attn.data_ptr()), _lib.current_stream()))
                return logits, list(attn.unbind(0))
            _lib.check(_lib.lib().mb_generator_forward(h, ctypes.c_void_p(tok.data_ptr()), n, ctypes.c_void_p(lab.data_ptr()), n,
                                                       drop_ptr, n, ctypes.c_void_p(logits.data_ptr()), _lib.current_stream()))
        return logits

    __call__ = forward


class Bert(LFQBert):
    """Mirror of modeling/bert.py:184-340, the embedding-table generator (``model_cls: "bert"``; no shipped config uses
    it): same constructor, attributes and forward as LFQBert.  Tokens are looked up in one embedding table per split
    (row ``effective_codebook_size`` = the mask token), and the logits of split i are the head output times the first
    ``effective_codebook_size`` rows of that same table plus a per-position bias (bert.py:313-333)."""
    _generator_cls = 1

    def _expected_spec(self):
        return [(n, s) for n, s, _ in bert_spec(**self._arch())]

    def _default_state_dict(self):
        return synthetic_bert_state_dict(seed=0, **self._arch())
