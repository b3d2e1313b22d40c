"""ORACLE -- test infrastructure, not product code.

CPU restatement (torch fp32 tensor primitives on the host: matmul, erf, exp, conv2d) of the reference's
sampling hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under maskbit_b200/ does.

Each function cites the reference lines it restates (paths relative to the reference repo root).
The functions are *functional*: they take the reference's state_dict (same key names) instead of
nn.Modules, so the same synthetic checkpoint feeds the reference, this oracle and the CUDA path.

Pinning: the reference ships no golden vectors for this path (its only result-pinning tests are the two
integer bit-layout assertion blocks factorization.py:49-67 and lookup_free.py:146-163, restated in
tests/test_oracle.py).  The oracle is therefore pinned against outputs of the reference itself, executed
in the build container: tests/golden/make_golden.py imports /root/reference, runs it on the synthetic
checkpoints and records logits / per-step tokens / pixels into tests/golden/*.npz; tests/test_oracle.py
checks this file against those fixtures (bit-exact for tokens, <=2e-5 abs for floating point).
"""
import math

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# integer bit conventions
# ----------------------------------------------------------------------------------------------
def combine_factorized_tokens(tokens, codebook_size, splits):
    """factorization.py:7-24 -- [B,n,m] group tokens -> [B,n] full index, returned as float32 like the reference."""
    out = torch.zeros((tokens.shape[0], tokens.shape[1]), dtype=torch.float32)
    shift = int(math.log2(codebook_size)) // splits
    for i in range(splits):
        out += (tokens[..., i] << (i * shift))
    return out


def split_factorized_tokens(tokens, codebook_size, splits):
    """factorization.py:27-46."""
    shift = int(math.log2(codebook_size)) // splits
    bm = (1 << shift) - 1
    return torch.stack([(tokens & (bm << (i * shift))) >> (i * shift) for i in range(splits)], dim=2)


def indices_to_bits(indices, bits):
    """lookup_free.py:96-111 get_codebook_entry: bit k <-> 2^k, coded as -1/+1 float."""
    b2i = (2 ** torch.arange(bits)).int()
    return ((indices.long()[..., None].int() & b2i) != 0).float() * 2.0 - 1.0


def bits_to_indices(tokens_pm1):
    """lookup_free.py:113-127 convert_bits_to_indices."""
    bits = tokens_pm1.shape[-1]
    b2i = (2 ** torch.arange(bits)).int()
    return ((tokens_pm1 > 0).int() * b2i).sum(-1)


# ----------------------------------------------------------------------------------------------
# LFQBert forward (bert.py:440-508)
# ----------------------------------------------------------------------------------------------
def layer_norm(x, w, b, eps=1e-12):
    """torch.nn.LayerNorm(eps=1e-12) (bert.py:33,86,394,414): biased variance, rsqrt(var+eps)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) * torch.rsqrt(var + eps) * w + b


def gelu_erf(x):
    """torch.nn.GELU() default = exact erf form (bert.py:29,413)."""
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def preprocess_tokens(img_tokens, bits, splits):
    """bert.py:440-454 -- group tokens -> +-1 bits, masked groups -> 0, (group-major, bit-minor)."""
    eff = bits // splits
    mask_token = 2 ** eff
    b2i = (2 ** torch.arange(eff)).int()
    mask = img_tokens == mask_token
    t = ((img_tokens[..., None].int() & b2i) != 0).float() * 2.0 - 1.0
    t[mask] = 0.0
    return t.reshape(img_tokens.shape[0], img_tokens.shape[1], splits * eff)


def mha(x, w_in, b_in, w_out, b_out, heads, attn_out=None):
    """nn.MultiheadAttention(batch_first, eval) (bert.py:84,137): packed in-proj, softmax(QK^T/sqrt(d))V, out-proj.
    attn_out, if a list, receives the head-averaged weights [N, S, S] (need_weights=True, average_attn_weights=True)."""
    n, s, d = x.shape
    hd = d // heads
    qkv = x @ w_in.t() + b_in
    q, k, v = qkv.split(d, dim=-1)
    q = q.view(n, s, heads, hd).transpose(1, 2)
    k = k.view(n, s, heads, hd).transpose(1, 2)
    v = v.view(n, s, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)), dim=-1)
    if attn_out is not None:
        attn_out.append(att.mean(dim=1))
    o = (att @ v).transpose(1, 2).reshape(n, s, d)
    return o @ w_out.t() + b_out


def _trunk_and_head(sd, x, heads, attn_out=None):
    """first_layer -> TransformerEncoder (post- or pre-norm) -> [norm_after_transformer] -> last_layer, shared by LFQBert
    (bert.py:496-500) and Bert (bert.py:324-328).  Returns (head output, per-layer hidden states)."""
    prenorm = "norm_after_transformer.weight" in sd
    x = layer_norm(x, sd["first_layer.0.weight"], sd["first_layer.0.bias"])
    hidden = [x]
    depth = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    for l in range(depth):
        p = f"transformer.layers.{l}."

        def attn(t):
            return mha(t, sd[p + "0.mha.in_proj_weight"], sd[p + "0.mha.in_proj_bias"],
                       sd[p + "0.mha.out_proj.weight"], sd[p + "0.mha.out_proj.bias"], heads, attn_out)

        def mlp(t):
            h = gelu_erf(t @ sd[p + "1.net.0.weight"].t() + sd[p + "1.net.0.bias"])
            return h @ sd[p + "1.net.2.weight"].t() + sd[p + "1.net.2.bias"]

        if prenorm:
            x = attn(layer_norm(x, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])) + x        # bert.py:118-121
            x = mlp(layer_norm(x, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"])) + x         # bert.py:57-59
        else:
            x = layer_norm(attn(x) + x, sd[p + "0.norm.weight"], sd[p + "0.norm.bias"])        # bert.py:137-139
            x = layer_norm(mlp(x) + x, sd[p + "1.norm.weight"], sd[p + "1.norm.bias"])         # bert.py:69-70
        hidden.append(x)
    if prenorm:
        x = layer_norm(x, sd["norm_after_transformer.weight"], sd["norm_after_transformer.bias"])   # bert.py:498-499
    y = gelu_erf(x @ sd["last_layer.0.weight"].t() + sd["last_layer.0.bias"])
    y = layer_norm(y, sd["last_layer.2.weight"], sd["last_layer.2.bias"])                  # bert.py:500
    return y, hidden


def lfq_bert_forward(sd, img_tokens, class_labels, drop_label_mask, *, heads=16, splits=2, nclass=1000,
                     return_hidden=False, return_attn=False):
    """LFQBert.forward (bert.py:456-508), post-norm, or pre-norm when the checkpoint carries ``norm_after_transformer``
    (use_prenorm=True: bert.py:49-59,106-123,407-408,498-499).  Returns fp32 logits [N, seq_len, splits, V].

    Does not mutate class_labels (the reference mutates a view in place, bert.py:484; harmless in sample()).
    drop_label_mask=None reproduces the reference quirk ``cls_token[None] = 1000`` (drops every label).
    """
    bits = sd["input_proj.weight"].shape[1]
    n, seq_len, _ = img_tokens.shape
    x_bits = preprocess_tokens(img_tokens, bits, splits)
    cls = class_labels.clone().view(n)
    if drop_label_mask is None:
        cls[:] = nclass
    else:
        cls[drop_label_mask] = nclass
    cls_emb = sd["class_emb.weight"][cls][:, None, :]
    proj = x_bits @ sd["input_proj.weight"].t() + sd["input_proj.bias"]
    x = torch.cat([proj, cls_emb], dim=1) + sd["pos_emb"]
    attn = [] if return_attn else None
    y, hidden = _trunk_and_head(sd, x, heads, attn)
    logits = y @ sd["prediction_layer.weight"].t() + sd["prediction_layer.bias"]
    v = logits.shape[-1] // splits
    logits = logits.view(n, seq_len + 1, splits, v)[:, :seq_len]                            # bert.py:502-503
    if return_attn:
        return logits, attn                                                                 # bert.py:505-506
    if return_hidden:
        return logits, hidden
    return logits


def bert_forward(sd, img_tokens, class_labels, drop_label_mask, *, heads=16, nclass=1000):
    """Bert.forward (bert.py:283-340), the embedding-table generator: token embeddings summed over the splits (row V of
    each table = the mask token), the shared trunk, and per split ``x @ tok_emb[i][:V].T + bias[i]`` on the first seq_len
    positions.  Returns fp32 logits [N, seq_len, splits, V]."""
    splits = sum(1 for k in sd if k.startswith("tok_emb_list."))
    n, seq_len, _ = img_tokens.shape
    cls = class_labels.clone().view(n)
    if drop_label_mask is None:
        cls[:] = nclass
    else:
        cls[drop_label_mask] = nclass
    cls_emb = sd["class_emb.weight"][cls][:, None, :]
    tok = sd["tok_emb_list.0.weight"][img_tokens[..., 0]]
    for i in range(1, splits):
        tok = tok + sd[f"tok_emb_list.{i}.weight"][img_tokens[..., i]]                  # bert.py:313-315
    x = torch.cat([tok, cls_emb], dim=1) + sd["pos_emb"]
    y, _ = _trunk_and_head(sd, x, heads)
    logits = []
    for i in range(splits):
        w = sd[f"tok_emb_list.{i}.weight"]
        v = w.shape[0] - 1
        logits.append((y @ w[:v].t())[:, :seq_len] + sd[f"bias.{i}"])                    # bert.py:331-333
    return torch.stack(logits, dim=2)


# ----------------------------------------------------------------------------------------------
# sampler schedules (host scalars)
# ----------------------------------------------------------------------------------------------
def get_masking_ratio(progress, mode="arccos"):
    """masking.py:41-65 -- fp32 0-d tensor, clamp [1e-6, 1]."""
    r = torch.tensor(progress)
    if mode == "root":
        v = 1 - (r ** 0.5)
    elif mode == "square":
        v = 1 - (r ** 2)
    elif mode == "cosine":
        v = torch.cos(r * math.pi * 0.5)
    elif mode == "arccos":
        v = torch.acos(r) / (math.pi * 0.5)
    elif mode == "linear":
        v = 1 - r
    else:
        raise ValueError("Invalid mode. Choose between 'linear','square', 'cosine', 'arccos', 'root'.")
    return torch.clamp(v, 1e-6, 1.0)


def guidance_scale_at(i, num_steps, guidance_scale, guidance_annealing, scale_pow):
    """sampling.py:91-98 -- python float or fp32 shape-[1] tensor, exactly as the reference builds it."""
    if guidance_annealing == "none":
        scale_step = 1.0
    elif guidance_annealing == "linear":
        scale_step = i / num_steps
    elif guidance_annealing == "cosine":
        sp = torch.ones((1)) * scale_pow
        scale_step = (1 - torch.cos(((i / num_steps) ** sp) * torch.pi)) * 1 / 2
    else:
        raise ValueError(f"unknown guidance_annealing {guidance_annealing}")
    return guidance_scale * scale_step


# ----------------------------------------------------------------------------------------------
# one sampling step given logits + noise (sampling.py:90-131)
# ----------------------------------------------------------------------------------------------
def select_step(logits_c, logits_u, scale, softmax_temperature, q_exp, gumbel_noise, noise_mult, mask_len,
                masked_tokens, mask_token):
    """torch restatement of one step's select path with the RNG draws made explicit.

    q_exp        [B*n*m, V] Exp(1) draws (Categorical.sample == argmax(p_hat / q))
    gumbel_noise [B, n, m]  raw Gumbel(0,1) draws; multiplied by noise_mult = randomize_temperature*(1-progress)
    mask_len     fp32 0-d tensor floor(ratio * num_maskable)
    Returns (predicted_tokens, new_masked_tokens).
    """
    mask = masked_tokens == mask_token
    if logits_u is not None:
        logits = logits_c + scale * (logits_c - logits_u)
    else:
        logits = logits_c
    probabilities = torch.softmax(logits / softmax_temperature, dim=-1)
    p_hat = probabilities / probabilities.sum(-1, keepdim=True)                  # Categorical.__init__
    flat = p_hat.reshape(-1, p_hat.shape[-1])
    predicted = torch.argmax(flat / q_exp, dim=-1).view(masked_tokens.shape)     # multinomial n=1 fast path
    num_masked = torch.sum(mask, dim=(1, 2))[0]
    predicted = torch.where(mask, predicted, masked_tokens)
    confidence = torch.gather(probabilities, -1, predicted.unsqueeze(-1)).squeeze(-1)
    confidence = torch.where(mask, confidence, torch.inf)
    confidence = torch.log(confidence) + gumbel_noise * noise_mult
    k = torch.clamp(mask_len, torch.ones_like(num_masked), num_masked - 1).long()
    srt = torch.sort(confidence.view(confidence.shape[0], -1), dim=-1).values
    thr = srt[:, k - 1]
    should_mask = confidence <= thr.unsqueeze(-1).unsqueeze(-1)
    new_masked = torch.where(should_mask, mask_token, predicted)
    return predicted, new_masked


def draw_step_noise(num_samples, n, m, v):
    """The per-step RNG consumption of the reference on CPU (validated bit-exactly):
    first the Categorical exponentials from the default generator, then the Gumbel uniforms."""
    q = torch.empty(num_samples * n * m, v).exponential_(1)
    g = torch.distributions.Gumbel(loc=0.0, scale=1.0).sample((num_samples, n, m))
    return q, g


def sample(gen_sd, dec_sd, num_samples, labels, *, softmax_temperature=1.0, randomize_temperature=4.5,
           mask_schedule_strategy="linear", num_steps=12, guidance_scale=3.0, mask_token=1024, patch_size=16,
           guidance_annealing="none", use_sampling_annealing=False, scale_pow=4.0, codebook_size=1024,
           codebook_splits=1, heads=16, forward_fn=None, decode=True, record=None):
    """sampling.py:12-136 restated; consumes the torch default CPU generator exactly like the reference on CPU.

    forward_fn(tokens[N,n,m], labels[N], drop[N]) -> logits lets tests substitute another forward (teacher forcing).
    record, if a list, receives per-step dicts (tokens_in, logits, q, g, predicted).
    """
    if forward_fn is None:
        def forward_fn(t, y, d):
            return lfq_bert_forward(gen_sd, t, y, d, heads=heads, splits=codebook_splits)
    n = int(patch_size ** 2)
    m = int(codebook_splits)
    drop = torch.ones(num_samples, dtype=torch.bool)
    masked_tokens = torch.full((num_samples, n, m), mask_token)
    num_maskable = n * m
    trace = []
    predicted = None
    for i in range(num_steps):
        progress = (i + 1) / num_steps
        if guidance_scale != 0.0:
            logits = forward_fn(torch.cat([masked_tokens, masked_tokens], 0), torch.cat([labels, labels], 0),
                                torch.cat([~drop, drop], 0))
            lc, lu = torch.chunk(logits, 2, dim=0)
            scale = guidance_scale_at(i, num_steps, guidance_scale, guidance_annealing, scale_pow)
        else:
            lc, lu, scale = forward_fn(masked_tokens, labels, ~drop), None, 0.0
        if use_sampling_annealing:
            softmax_temperature = 0.5 + 0.8 * (1 - progress)
        q, g = draw_step_noise(num_samples, n, m, lc.shape[-1])
        ratio = get_masking_ratio(progress, mode=mask_schedule_strategy)
        mask_len = torch.floor(ratio * num_maskable)
        tokens_in = masked_tokens
        predicted, masked_tokens = select_step(lc, lu, scale, softmax_temperature, q, g,
                                               randomize_temperature * (1 - progress), mask_len, masked_tokens, mask_token)
        if record is not None:
            record.append(dict(tokens_in=tokens_in, logits_c=lc, logits_u=lu, scale=scale, q=q, g=g,
                               predicted=predicted, mask_len=mask_len))
        trace.append(predicted)
    if not decode:
        return None, trace
    combined = combine_factorized_tokens(predicted, codebook_size, codebook_splits)
    return decode_tokens(dec_sd, combined), trace


# ----------------------------------------------------------------------------------------------
# ConvVQModel.decode_tokens (conv_vqgan.py:98-112, autoencoder.py:358-423)
# ----------------------------------------------------------------------------------------------
def conv_same(x, w, b=None, stride=1):
    """Conv2dSame (autoencoder.py:7-36): TF-style SAME padding, extra pixel on the bottom/right."""
    ih, iw = x.shape[-2:]
    k = w.shape[-1]
    ph = max((math.ceil(ih / stride) - 1) * stride + (k - 1) + 1 - ih, 0)
    pw = max((math.ceil(iw / stride) - 1) * stride + (k - 1) + 1 - iw, 0)
    if ph > 0 or pw > 0:
        x = F.pad(x, [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2])
    return F.conv2d(x, w, b, stride=stride)


def group_norm_silu(x, w, b):
    """GroupNorm(32, eps=1e-6, affine) + SiLU (autoencoder.py:39-43,85-90)."""
    return F.silu(F.group_norm(x, 32, w, b, eps=1e-6))


def res_block(sd, p, x):
    """ResidualBlock.forward (autoencoder.py:84-96) incl. the quirk that nin_shortcut is applied to the
    conv2 output (out = h + W_nin h), not to the block input."""
    h = conv_same(group_norm_silu(x, sd[p + "norm1.weight"], sd[p + "norm1.bias"]), sd[p + "conv1.weight"])
    h = conv_same(group_norm_silu(h, sd[p + "norm2.weight"], sd[p + "norm2.bias"]), sd[p + "conv2.weight"])
    if (p + "nin_shortcut.weight") in sd:
        return h + conv_same(h, sd[p + "nin_shortcut.weight"])
    return h + x


def conv_decoder(sd, z, num_resolutions=5, num_res_blocks=2, prefix="decoder."):
    """ConvDecoder.forward (autoencoder.py:399-423)."""
    h = conv_same(z, sd[prefix + "conv_in.weight"], sd[prefix + "conv_in.bias"])
    for r in range(num_res_blocks):
        h = res_block(sd, f"{prefix}mid.res_blocks.{r}.", h)
    for j in range(num_resolutions):
        for r in range(num_res_blocks):
            h = res_block(sd, f"{prefix}up.{j}.res_blocks.{r}.", h)
        if j < num_resolutions - 1:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")                       # autoencoder.py:224
            h = conv_same(h, sd[f"{prefix}up.{j}.upsample_conv.weight"], sd[f"{prefix}up.{j}.upsample_conv.bias"])
    h = group_norm_silu(h, sd[prefix + "norm_out.weight"], sd[prefix + "norm_out.bias"])
    return conv_same(h, sd[prefix + "conv_out.weight"], sd[prefix + "conv_out.bias"])


def decode_tokens(sd, tokens, num_resolutions=5, num_res_blocks=2):
    """ConvVQModel.decode_tokens (conv_vqgan.py:98-112): tokens [B, n] (any int/float dtype) -> fp32 [B,3,H,W]."""
    bits = sd["quantize.bits_to_indices"].shape[0]
    z = indices_to_bits(tokens, bits)
    ss = int(math.sqrt(float(z.shape[1])))
    z = z.reshape(z.shape[0], ss, ss, -1).permute(0, 3, 1, 2).contiguous()
    return conv_decoder(sd, z, num_resolutions, num_res_blocks)


# ----------------------------------------------------------------------------------------------
# ConvVQModel.encode / forward (conv_vqgan.py:71-84,114-132, autoencoder.py:138-184,230-286, lookup_free.py:46-94)
# ----------------------------------------------------------------------------------------------
def conv_encoder(sd, x, num_resolutions=5, num_res_blocks=2, prefix="encoder."):
    """ConvEncoder.forward (autoencoder.py:268-286) with sample_with_conv=True (every shipped config): conv_in (no bias),
    per level num_res_blocks ResidualBlocks then a stride-2 3x3 Conv2dSame (pad 0 top/left, 1 bottom/right), a last
    ResidualStage without downsampling, mid blocks, GroupNorm + SiLU + 1x1 conv_out."""
    h = conv_same(x, sd[prefix + "conv_in.weight"])
    for lvl in range(num_resolutions):
        for r in range(num_res_blocks):
            h = res_block(sd, f"{prefix}down.{lvl}.res_blocks.{r}.", h)
        if lvl < num_resolutions - 1:
            h = conv_same(h, sd[f"{prefix}down.{lvl}.down_conv.weight"], sd[f"{prefix}down.{lvl}.down_conv.bias"], stride=2)
    for r in range(num_res_blocks):
        h = res_block(sd, f"{prefix}mid.res_blocks.{r}.", h)
    h = group_norm_silu(h, sd[prefix + "norm_out.weight"], sd[prefix + "norm_out.bias"])
    return conv_same(h, sd[prefix + "conv_out.weight"], sd[prefix + "conv_out.bias"])


def lfq_quantize(z):
    """LookupFreeQuantizer.forward in eval mode (lookup_free.py:46-94): returns (z_quantized [B,C,H,W], indices [B,H,W]).
    z_quantized = z + (sign - z), i.e. +-1 up to fp32 rounding, exactly as the reference computes it."""
    zt = z.permute(0, 2, 3, 1).contiguous()
    ones = torch.ones_like(zt)
    zq = torch.where(zt > 0.0, ones, -ones)
    idx = bits_to_indices(zq)
    zq = zt + (zq - zt)
    return zq.permute(0, 3, 1, 2).contiguous(), idx


def encode(sd, x, num_resolutions=5, num_res_blocks=2):
    """ConvVQModel.encode (conv_vqgan.py:71-84): returns (z_quantized, indices, z)."""
    z = conv_encoder(sd, x, num_resolutions, num_res_blocks)
    zq, idx = lfq_quantize(z)
    return zq, idx, z


def autoencode(sd, x, num_resolutions=5, num_res_blocks=2):
    """ConvVQModel.forward (conv_vqgan.py:114-132): (reconstruction, indices)."""
    zq, idx, _ = encode(sd, x, num_resolutions, num_res_blocks)
    return conv_decoder(sd, zq, num_resolutions, num_res_blocks), idx


# ----------------------------------------------------------------------------------------------
# forward half of the training step (scripts/train_maskbit.py:362-380)
# ----------------------------------------------------------------------------------------------
def get_mask_tokens(tokens, mask_token, mode="arccos", min_masking_ratio=0.0):
    """masking.py:7-38: per-sample ratio from rand(B), per-slot rand(tokens.size()) < ratio, both from the default CPU generator."""
    r = torch.rand(tokens.size(0)) * (1 - min_masking_ratio)
    if mode == "linear":
        val = 1 - r
    elif mode == "square":
        val = 1 - (r ** 2)
    elif mode == "cosine":
        val = torch.cos(r * math.pi * 0.5)
    elif mode == "arccos":
        val = torch.acos(r) / (math.pi * 0.5)
    else:
        raise ValueError("Invalid mode. Choose between 'linear','square', 'cosine', 'arccos'.")
    mask = torch.rand(tokens.size()) < val.view(-1, 1, 1)
    masked = tokens.detach().clone()
    masked[mask] = mask_token
    return masked, mask


def mlm_loss(inputs, targets, masks, label_smoothing=0.1, sum_splits=False):
    """losses.py:289-339 MLMLoss.forward written out: label-smoothed cross entropy (1-eps) * nll + eps * mean_j(-logp_j), mean over
    rows; accuracies as mean(argmax == target) ** m; the same over the masked rows.  Returns the four scalars as python floats."""
    b, n, m, v = inputs.shape
    logp = torch.log_softmax(inputs.reshape(-1, v).double(), dim=-1)
    t = targets.reshape(-1)
    nll = -logp.gather(1, t[:, None]).squeeze(1)
    smooth = -logp.mean(dim=1)
    row = (1 - label_smoothing) * nll + label_smoothing * smooth
    hit = (inputs.reshape(-1, v).argmax(-1) == t).double()
    mk = masks.reshape(-1)
    scale = m if sum_splits else 1
    return (float(row.mean()) * scale, float(hit.mean()) ** m, float(row[mk].mean()) * scale, float(hit[mk].mean()) ** m)
