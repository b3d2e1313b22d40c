"""Deterministic synthetic checkpoints in the reference's state-dict layout.

There is no network for real checkpoints, so parity tests and the benchmark use random-init weights.
The reference initialises with torch's RNG (bert.py:421-435, conv default kaiming) whose CPU kernels are
not guaranteed bit-stable across hosts; the golden fixtures must be reproducible on the GPU box, so the
values here come from an integer counter hash (splitmix64 finaliser) evaluated in numpy: identical on
every machine.  Key names and shapes are exactly those of LFQBert.state_dict() (bert.py:345-419) and
ConvVQModel.state_dict() (conv_vqgan.py:39-62, autoencoder.py:230-286,358-397), so the reference model
loads them with ``load_state_dict(strict=True)`` (tests/golden/make_golden.py does that).

Unlike the reference init, biases and norm affine parameters are non-trivial on purpose: zero biases and
unit scales would hide bugs in the bias / affine paths of the kernels.
"""
import math
import zlib
from collections import OrderedDict

import numpy as np
import torch

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def hash_normal(name, shape, seed=0):
    """Approximately N(0,1) values (Irwin-Hall of four 16-bit uniforms), float64, reproducible bit-for-bit."""
    n = int(np.prod(shape)) if len(shape) else 1
    key = np.uint64((zlib.crc32(name.encode()) << 20) ^ (seed * 0x1000193 + 0x811C9DC5))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) * np.uint64(0xD1342543DE82EF95) + key
        x = _splitmix(idx)
    m = np.uint64(0xFFFF)
    s = ((x & m) + ((x >> np.uint64(16)) & m) + ((x >> np.uint64(32)) & m) + ((x >> np.uint64(48)) & m)).astype(np.int64)
    z = (s - 2 * 65535).astype(np.float64) * (1.0 / (65536.0 * math.sqrt(4.0 / 12.0)))
    return z.reshape(shape)


def _t(name, shape, scale, seed, mean=0.0):
    return torch.from_numpy((hash_normal(name, shape, seed) * scale + mean).astype(np.float32))


def lfq_bert_spec(hidden_dim=1024, codebook_size=4096, codebook_splits=2, depth=24, mlp_dim=4096, nclass=1000,
                  seq_len=256, use_prenorm=False):
    """(name, shape, kind) for every LFQBert tensor, in state_dict order."""
    bits = int(math.log2(codebook_size))
    eff = bits // codebook_splits
    D = hidden_dim
    spec = [("pos_emb", (1, seq_len + 1, D), "w"), ("bits_to_indices", (eff,), "b2i"),
            ("class_emb.weight", (nclass + 1, D), "w"),
            ("input_proj.weight", (D, bits), "w"), ("input_proj.bias", (D,), "b"),
            ("first_layer.0.weight", (D,), "g"), ("first_layer.0.bias", (D,), "b")]
    for l in range(depth):
        p = f"transformer.layers.{l}."
        spec += [(p + "0.mha.in_proj_weight", (3 * D, D), "w"), (p + "0.mha.in_proj_bias", (3 * D,), "b"),
                 (p + "0.mha.out_proj.weight", (D, D), "w"), (p + "0.mha.out_proj.bias", (D,), "b"),
                 (p + "0.norm.weight", (D,), "g"), (p + "0.norm.bias", (D,), "b"),
                 (p + "1.net.0.weight", (mlp_dim, D), "w"), (p + "1.net.0.bias", (mlp_dim,), "b"),
                 (p + "1.net.2.weight", (D, mlp_dim), "w"), (p + "1.net.2.bias", (D,), "b"),
                 (p + "1.norm.weight", (D,), "g"), (p + "1.norm.bias", (D,), "b")]
    if use_prenorm:   # bert.py:407-408
        spec += [("norm_after_transformer.weight", (D,), "g"), ("norm_after_transformer.bias", (D,), "b")]
    spec += [("last_layer.0.weight", (D, D), "w"), ("last_layer.0.bias", (D,), "b"),
             ("last_layer.2.weight", (D,), "g"), ("last_layer.2.bias", (D,), "b"),
             ("prediction_layer.weight", (codebook_splits * 2 ** eff, D), "w"),
             ("prediction_layer.bias", (codebook_splits * 2 ** eff,), "b")]
    return spec


def bert_spec(hidden_dim=1024, codebook_size=4096, codebook_splits=2, depth=24, mlp_dim=4096, nclass=1000, seq_len=256,
              use_prenorm=False):
    """(name, shape, kind) for every tensor of the embedding-table generator ``Bert`` (bert.py:184-258): the transformer
    and last_layer of LFQBert, per-split token embeddings (row V = the mask token) that are also the output projection,
    and a per-position logit bias per split."""
    bits = int(math.log2(codebook_size))
    v = 2 ** (bits // codebook_splits)
    D = hidden_dim
    trunk = [e for e in lfq_bert_spec(hidden_dim, codebook_size, codebook_splits, depth, mlp_dim, nclass, seq_len, use_prenorm)
             if e[0].startswith(("first_layer", "transformer", "norm_after_transformer", "last_layer"))]
    spec = [("pos_emb", (1, seq_len + 1, D), "w"), ("class_emb.weight", (nclass + 1, D), "w")]
    spec += [(f"tok_emb_list.{i}.weight", (v + 1, D), "w") for i in range(codebook_splits)]
    spec += trunk
    spec += [(f"bias.{i}", (seq_len, v), "b") for i in range(codebook_splits)]
    return spec


def synthetic_bert_state_dict(seed=0, weight_std=0.02, **arch):
    sd = OrderedDict()
    for name, shape, kind in bert_spec(**arch):
        sd[name] = _t(name, shape, {"w": weight_std, "b": 0.02, "g": 0.1}[kind], seed, mean=1.0 if kind == "g" else 0.0)
    return sd


def synthetic_lfq_bert_state_dict(seed=0, weight_std=0.02, **arch):
    """Synthetic LFQBert checkpoint.  weight_std 0.02 = the reference's trunc-normal sigma (bert.py:427-432)."""
    sd = OrderedDict()
    for name, shape, kind in lfq_bert_spec(**arch):
        if kind == "w":
            sd[name] = _t(name, shape, weight_std, seed)
        elif kind == "b":
            sd[name] = _t(name, shape, 0.02, seed)
        elif kind == "g":
            sd[name] = _t(name, shape, 0.1, seed, mean=1.0)
        elif kind == "b2i":
            sd[name] = (2 ** torch.arange(shape[0])).int()
    return sd


def _res_block(prefix, cin, cout):
    s = [(prefix + "norm1.weight", (cin,), "g"), (prefix + "norm1.bias", (cin,), "gb"),
         (prefix + "conv1.weight", (cout, cin, 3, 3), "c"),
         (prefix + "norm2.weight", (cout,), "g"), (prefix + "norm2.bias", (cout,), "gb"),
         (prefix + "conv2.weight", (cout, cout, 3, 3), "c")]
    if cin != cout:
        s.append((prefix + "nin_shortcut.weight", (cout, cout, 1, 1), "c"))
    return s


def conv_vq_spec(token_size=12, num_channels=3, hidden_channels=128, channel_mult=(1, 1, 2, 2, 4), num_resolutions=5,
                 num_res_blocks=2, with_encoder=True, num_res_blocks_encoder=None):
    """(name, shape, kind) for ConvVQModel (encoder, decoder, quantize buffers), in state_dict order.  num_res_blocks is the
    DECODER's count (config num_res_blocks_decoder, else num_res_blocks); the encoder always uses the config's num_res_blocks."""
    hc = hidden_channels
    enc_blocks = num_res_blocks if num_res_blocks_encoder is None else num_res_blocks_encoder
    cm = tuple(channel_mult)
    spec = []
    if with_encoder:
        spec.append(("encoder.conv_in.weight", (hc, num_channels, 3, 3), "c"))
        icm = (1,) + cm
        for lvl in range(num_resolutions):
            cin, cout = hc * icm[lvl], hc * icm[lvl + 1]
            for r in range(enc_blocks):
                spec += _res_block(f"encoder.down.{lvl}.res_blocks.{r}.", cin if r == 0 else cout, cout)
            if lvl < num_resolutions - 1:
                spec += [(f"encoder.down.{lvl}.down_conv.weight", (cout, cout, 3, 3), "c"),
                         (f"encoder.down.{lvl}.down_conv.bias", (cout,), "cb")]
        mid = hc * cm[num_resolutions - 1]
        for r in range(enc_blocks):
            spec += _res_block(f"encoder.mid.res_blocks.{r}.", mid, mid)
        spec += [("encoder.norm_out.weight", (mid,), "g"), ("encoder.norm_out.bias", (mid,), "gb"),
                 ("encoder.conv_out.weight", (token_size, mid, 1, 1), "c"), ("encoder.conv_out.bias", (token_size,), "cb")]
    block_in = hc * cm[num_resolutions - 1]
    spec += [("decoder.conv_in.weight", (block_in, token_size, 3, 3), "c"), ("decoder.conv_in.bias", (block_in,), "cb")]
    for r in range(num_res_blocks):
        spec += _res_block(f"decoder.mid.res_blocks.{r}.", block_in, block_in)
    icm = cm + (cm[-1],)
    for j, lvl in enumerate(reversed(range(num_resolutions))):
        cin, cout = hc * icm[lvl + 1], hc * icm[lvl]
        for r in range(num_res_blocks):
            spec += _res_block(f"decoder.up.{j}.res_blocks.{r}.", cin if r == 0 else cout, cout)
        if lvl > 0:
            spec += [(f"decoder.up.{j}.upsample_conv.weight", (cout, cout, 3, 3), "c"),
                     (f"decoder.up.{j}.upsample_conv.bias", (cout,), "cb")]
    spec += [("decoder.norm_out.weight", (cout,), "g"), ("decoder.norm_out.bias", (cout,), "gb"),
             ("decoder.conv_out.weight", (num_channels, cout, 3, 3), "c"), ("decoder.conv_out.bias", (num_channels,), "cb"),
             ("quantize.bits_to_indices", (token_size,), "b2i"), ("quantize.codebook", (2 ** token_size, token_size), "cbk")]
    return spec


def synthetic_conv_vq_state_dict(seed=0, **arch):
    """Synthetic ConvVQModel checkpoint: conv weights ~ N(0, 1/fan_in) (kaiming-like), GN affine around (1, 0)."""
    sd = OrderedDict()
    for name, shape, kind in conv_vq_spec(**arch):
        if kind == "c":
            fan_in = shape[1] * shape[2] * shape[3]
            sd[name] = _t(name, shape, 1.0 / math.sqrt(fan_in), seed)
        elif kind == "cb":
            sd[name] = _t(name, shape, 0.05, seed)
        elif kind == "g":
            sd[name] = _t(name, shape, 0.1, seed, mean=1.0)
        elif kind == "gb":
            sd[name] = _t(name, shape, 0.1, seed)
        elif kind == "b2i":
            sd[name] = (2 ** torch.arange(shape[0])).int()
        elif kind == "cbk":
            # lookup_free.py:36-44: all 2^bits codes as +-1 rows, bit k <-> 2^k
            idx = torch.arange(shape[0])
            b2i = (2 ** torch.arange(shape[1]))
            sd[name] = ((idx[:, None] & b2i) != 0).float() * 2.0 - 1.0
    return sd


def trained_like_lfq_bert_state_dict(seed=11, **arch):
    """LFQBert checkpoint with the statistics a TRAINED post-norm transformer shows and a fresh init does not: LayerNorm gains spread
    over 0.1 .. 5 (log-normal) with two outlier channels, LayerNorm biases with a common offset and a few +-3 entries, non-zero
    Linear biases, wider projections (attention logits of several units, output logits of tens).  Used by the parity tests that
    stress the LayerNorm folding and the bf16 pre-norm residual stream (gemm_tcgen05.cuh); values are the counter hash above."""
    sd = OrderedDict()
    for name, shape, kind in lfq_bert_spec(**arch):
        if kind == "b2i":
            sd[name] = (2 ** torch.arange(shape[0])).int()
        elif kind == "g":
            z = hash_normal(name, shape, seed)
            g = np.clip(np.exp(0.5 * z), 0.1, 5.0)
            g[7] = 5.0
            g[300] = 4.0
            sd[name] = torch.from_numpy(g.astype(np.float32))
        elif kind == "b" and ("norm" in name or name.startswith(("first_layer", "last_layer.2"))):
            z = hash_normal(name, shape, seed) * 0.3
            layer_sign = 1.0 if (zlib.crc32(name.encode()) & 1) else -1.0
            b = z + 0.5 * layer_sign
            b[11] = 3.0
            b[500] = -3.0
            sd[name] = torch.from_numpy(b.astype(np.float32))
        elif kind == "b":
            sd[name] = _t(name, shape, 0.1, seed)
        else:
            std = 0.035
            if name.endswith(("out_proj.weight", "net.2.weight")):
                std = 0.02
            elif name.startswith("prediction_layer"):
                std = 0.1
            elif name in ("pos_emb", "class_emb.weight", "input_proj.weight"):
                std = 0.1
            sd[name] = _t(name, shape, std, seed)
            if name.endswith("mha.in_proj_weight"):
                # query / key rows at 0.45 of the value rows: attention logits with a standard deviation of 3-4 and maxima of ~15
                # (sharp but not degenerate; at the full scale the outlier channels drive them to a std of 17, where a softmax row
                # is a near-tie lottery and ANY reduced-precision run -- the reference's own bf16 autocast included -- is chaotic)
                sd[name][: 2 * shape[1]] *= 0.45
    return sd
