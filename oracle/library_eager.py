"""ORACLE-side yardstick -- test / measurement infrastructure, not product code.

The sampling step and the decoder restated with PyTorch's own CUDA library kernels (cuBLASLt GEMMs through F.linear, the fused
scaled_dot_product_attention kernels nn.MultiheadAttention dispatches to in eval mode, cuDNN convolutions, native LayerNorm /
GroupNorm / GELU), run the two ways the reference can run on a GPU: fp32 tensors with TF32 matmuls and convolutions
(scripts/eval_maskbit.py:69-72 sets allow_tf32) and bf16 autocast.  bench.py times it right after its own timed regions and reports
it as `roofline.library_eager`: "the Blackwell library path to beat" (SURVEY.md 8d last row).  Never imported by maskbit_b200/.

Same operator sequence as oracle/maskbit_oracle.py (which is pinned to the reference): lfq_bert_forward (bert.py:440-508) and
conv_decoder (autoencoder.py:399-423); the select step is left out (it is <0.5 % of a step on either side).
"""
import math
import time

import torch
import torch.nn.functional as F


def _forward(sd, tokens, labels, drop, heads=16, splits=2, nclass=1000):
    bits = sd["input_proj.weight"].shape[1]
    eff = bits // splits
    n, seq_len, _ = tokens.shape
    b2i = (2 ** torch.arange(eff, device=tokens.device)).int()
    x_bits = ((tokens[..., None].int() & b2i) != 0).float() * 2.0 - 1.0
    x_bits[tokens == 2 ** eff] = 0.0
    x_bits = x_bits.reshape(n, seq_len, bits)
    cls = labels.clone()
    cls[drop] = nclass
    x = torch.cat([F.linear(x_bits, sd["input_proj.weight"], sd["input_proj.bias"]), sd["class_emb.weight"][cls][:, None, :]], 1) + sd["pos_emb"]
    d = x.shape[-1]
    x = F.layer_norm(x, (d,), sd["first_layer.0.weight"], sd["first_layer.0.bias"], 1e-12)
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    for l in range(depth):
        p = f"transformer.layers.{l}."
        qkv = F.linear(x, sd[p + "0.mha.in_proj_weight"], sd[p + "0.mha.in_proj_bias"])
        q, k, v = (t.view(n, seq_len + 1, heads, d // heads).transpose(1, 2) for t in qkv.split(d, dim=-1))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(n, seq_len + 1, d)
        x = F.layer_norm(F.linear(o, sd[p + "0.mha.out_proj.weight"], sd[p + "0.mha.out_proj.bias"]) + x, (d,),
                         sd[p + "0.norm.weight"], sd[p + "0.norm.bias"], 1e-12)
        h = F.gelu(F.linear(x, sd[p + "1.net.0.weight"], sd[p + "1.net.0.bias"]))
        x = F.layer_norm(F.linear(h, sd[p + "1.net.2.weight"], sd[p + "1.net.2.bias"]) + x, (d,), sd[p + "1.norm.weight"], sd[p + "1.norm.bias"], 1e-12)
    y = F.layer_norm(F.gelu(F.linear(x, sd["last_layer.0.weight"], sd["last_layer.0.bias"])), (d,), sd["last_layer.2.weight"],
                     sd["last_layer.2.bias"], 1e-12)
    logits = F.linear(y, sd["prediction_layer.weight"], sd["prediction_layer.bias"])
    return logits.view(n, seq_len + 1, splits, -1)[:, :seq_len]


def _conv_same(x, w, b=None):
    return F.conv2d(F.pad(x, [1, 1, 1, 1]) if w.shape[-1] == 3 else x, w, b)


def _block(sd, p, x):
    h = _conv_same(F.silu(F.group_norm(x, 32, sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-6)), sd[p + "conv1.weight"])
    h = _conv_same(F.silu(F.group_norm(h, 32, sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-6)), sd[p + "conv2.weight"])
    if (p + "nin_shortcut.weight") in sd:
        return h + _conv_same(h, sd[p + "nin_shortcut.weight"])
    return h + x


def _decode(sd, tokens, bits):
    b2i = (2 ** torch.arange(bits, device=tokens.device)).int()
    z = ((tokens.long()[..., None].int() & b2i) != 0).float() * 2.0 - 1.0
    z = z.reshape(z.shape[0], 16, 16, bits).permute(0, 3, 1, 2).contiguous()
    pre = "decoder."
    h = _conv_same(z, sd[pre + "conv_in.weight"], sd[pre + "conv_in.bias"])
    for r in range(2):
        h = _block(sd, f"{pre}mid.res_blocks.{r}.", h)
    for j in range(5):
        for r in range(2):
            h = _block(sd, f"{pre}up.{j}.res_blocks.{r}.", h)
        if j < 4:
            h = _conv_same(F.interpolate(h, scale_factor=2.0, mode="nearest"), sd[f"{pre}up.{j}.upsample_conv.weight"],
                           sd[f"{pre}up.{j}.upsample_conv.bias"])
    h = F.silu(F.group_norm(h, 32, sd[pre + "norm_out.weight"], sd[pre + "norm_out.bias"], 1e-6))
    return _conv_same(h, sd[pre + "conv_out.weight"], sd[pre + "conv_out.bias"])


@torch.no_grad()
def time_library_eager(gen_sd, dec_sd, bits, batch, sampling_steps, device, fwd_steps=3, dec_chunk=32):
    """images/s of the torch-library path for `sampling_steps` guided steps of `batch` images + decode, from `fwd_steps` timed
    double-batch forwards (every step does identical work) and one timed decode of the whole batch in chunks of `dec_chunk`.
    Returns {"tf32": {...}, "bf16_autocast": {...}}."""
    gsd = {k: v.to(device) for k, v in gen_sd.items() if v.is_floating_point()}
    dsd = {k: v.to(device) for k, v in dec_sd.items() if k.startswith("decoder.")}
    v = 2 ** (bits // 2)
    g = torch.Generator(device="cpu").manual_seed(5)
    tok = torch.randint(0, v + 1, (batch, 256, 2), generator=g).to(device)
    tok2 = torch.cat([tok, tok])
    labels = torch.randint(0, 1000, (batch,), generator=g).to(device)
    lab2 = torch.cat([labels, labels])
    drop = torch.cat([torch.zeros(batch, dtype=torch.bool), torch.ones(batch, dtype=torch.bool)]).to(device)
    codes = torch.randint(0, 2 ** bits, (batch, 256), generator=g).to(device)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    out = {}
    try:
        for name in ("tf32", "bf16_autocast"):
            torch.backends.cuda.matmul.allow_tf32 = True        # eval_maskbit.py:69-72
            torch.backends.cudnn.allow_tf32 = True
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if name == "bf16_autocast" else torch.autocast("cuda", enabled=False)
            with ctx:
                _forward(gsd, tok2, lab2, drop)                  # warm-up (cuBLASLt heuristics, cuDNN algorithm choice)
                _decode(dsd, codes[:dec_chunk], bits)
                torch.cuda.synchronize(device)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                for _ in range(fwd_steps):
                    _forward(gsd, tok2, lab2, drop)
                e[1].record()
                for b0 in range(0, batch, dec_chunk):
                    _decode(dsd, codes[b0:b0 + dec_chunk], bits)
                e[2].record()
                torch.cuda.synchronize(device)
            ms_step = e[0].elapsed_time(e[1]) / fwd_steps
            ms_dec = e[1].elapsed_time(e[2])
            total_ms = sampling_steps * ms_step + ms_dec
            out[name] = {"images_per_s": batch / (total_ms / 1000.0), "ms_per_sampling_step": ms_step, "ms_decode": ms_dec}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["what"] = (f"torch-library eager path on this GPU (F.linear / scaled_dot_product_attention / cuDNN conv), B={batch}: {fwd_steps} timed "
                   f"CFG double-batch forwards scaled to {sampling_steps} steps + one timed decode; select step not included; yardstick only")
    return out
