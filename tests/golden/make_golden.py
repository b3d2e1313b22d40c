"""Generate the golden fixtures in this directory by running the UNMODIFIED reference (imported from
/root/reference, CPU, fp32) on the deterministic synthetic checkpoints of maskbit_b200.weights.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Fixtures written (all small enough to commit):
  forward_12bit.npz   LFQBert.forward logits for 4 sequences (cond/uncond halves, partially masked tokens)
  forward_14bit.npz   same for the 14-bit model (V=128), 2 sequences
  select_12bit.npz    sample() B=2, 4 steps: per-step generator logits (forward hook), replayed RNG draws
                      (q ~ Exp(1), g ~ Gumbel) and the reference's per-step predicted tokens
  sample_12bit.npz    BASELINE config #1: sample() B=4, 8 steps, CFG cosine: per-step tokens + pixels
  forward_prenorm_12bit.npz  LFQBert(use_prenorm=True, depth=2).forward, 2 sequences (no shipped config uses pre-norm; the
                      branch is bert.py:49-59,106-123,498-499)
  forward_bert_12bit.npz  Bert(depth=2).forward (embedding-table generator, bert.py:184-340), 2 sequences
  decode_12bit.npz    ConvVQModel.decode_tokens on random tokens, B=2
  encode_12bit.npz    ConvVQModel.forward (encode -> LFQ -> decode) on seeded images, B=2: latents z, indices, reconstruction
"""
import os
import sys
import time

import numpy as np
import torch

REF = os.environ.get("MASKBIT_REF", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)

from modeling.bert import Bert, LFQBert  # noqa: E402  (reference)
from modeling.conv_vqgan import ConvVQModel  # noqa: E402  (reference)
from modeling.modules import sample as ref_sample  # noqa: E402  (reference)

from maskbit_b200.config import load_config, sampler_kwargs  # noqa: E402
from maskbit_b200.weights import synthetic_bert_state_dict, synthetic_conv_vq_state_dict, synthetic_lfq_bert_state_dict  # noqa: E402


def build_reference(bits):
    cfg = load_config(f"maskbit_generator_{bits}bit")
    kw = sampler_kwargs(cfg)
    vq = ConvVQModel(cfg.model.vq_model, legacy=False)
    vq.load_state_dict(synthetic_conv_vq_state_dict(seed=0, token_size=bits), strict=True)
    vq.eval().requires_grad_(False)
    mlm = cfg.model.mlm_model
    gen = LFQBert(img_size=256, hidden_dim=mlm.hidden_dim, codebook_size=cfg.model.vq_model.codebook_size,
                  codebook_splits=mlm.codebook_splits, depth=mlm.depth, heads=mlm.heads, mlp_dim=mlm.mlp_dim,
                  dropout=mlm.dropout, use_prenorm=mlm.use_prenorm, input_stride=16)
    gen.load_state_dict(synthetic_lfq_bert_state_dict(seed=0, codebook_size=2 ** bits), strict=True)
    gen.eval().requires_grad_(False)
    return cfg, kw, vq, gen


def partially_masked_tokens(n, v, seed):
    g = torch.Generator().manual_seed(seed)
    tok = torch.randint(0, v, (n, 256, 2), generator=g)
    frac = torch.linspace(0.1, 0.9, n).view(n, 1, 1)
    m = torch.rand((n, 256, 2), generator=g) < frac
    tok[m] = v
    return tok


def golden_forward(gen, bits, n, path):
    v = 2 ** (bits // 2)
    tok = partially_masked_tokens(n // 2, v, seed=100 + bits)
    tok = torch.cat([tok, tok], 0)
    g = torch.Generator().manual_seed(7)
    labels = torch.randint(0, 1000, (n // 2,), generator=g)
    labels = torch.cat([labels, labels], 0)
    drop = torch.cat([torch.zeros(n // 2, dtype=torch.bool), torch.ones(n // 2, dtype=torch.bool)])
    with torch.no_grad():
        logits = gen(tok.clone(), labels.clone(), drop)
    np.savez_compressed(path, tokens=tok.numpy().astype(np.int16), labels=labels.numpy(), drop=drop.numpy(),
                        logits=logits.numpy())
    print(path, logits.shape, float(logits.abs().max()))


def golden_forward_prenorm(path):
    """Pre-norm branch on a 2-layer generator of the shipped width (hidden 1024, 16 heads, MLP 4096, 12-bit)."""
    gen = LFQBert(img_size=256, hidden_dim=1024, codebook_size=4096, codebook_splits=2, depth=2, heads=16, mlp_dim=4096,
                  dropout=0.1, use_prenorm=True, input_stride=16)
    gen.load_state_dict(synthetic_lfq_bert_state_dict(seed=3, codebook_size=4096, depth=2, use_prenorm=True), strict=True)
    gen.eval().requires_grad_(False)
    golden_forward(gen, 12, 2, path)


def golden_forward_bert(path):
    """Embedding-table generator on a 2-layer model of the shipped width (hidden 1024, 16 heads, MLP 4096, 4096 codes in 2 splits)."""
    gen = Bert(img_size=256, hidden_dim=1024, codebook_size=4096, codebook_splits=2, depth=2, heads=16, mlp_dim=4096,
               dropout=0.1, use_prenorm=False, input_stride=16)
    gen.load_state_dict(synthetic_bert_state_dict(seed=5, codebook_size=4096, depth=2), strict=True)
    gen.eval().requires_grad_(False)
    golden_forward(gen, 12, 2, path)


def golden_sample(kw, vq, gen, b, steps, path, with_logits):
    g = torch.Generator().manual_seed(1234)
    labels = torch.randint(0, 1000, (b,), generator=g)
    kw = dict(kw, num_steps=steps)
    rec = []
    hook = gen.register_forward_hook(lambda mod, inp, out: rec.append(out.detach().clone()))
    torch.manual_seed(1234)
    t0 = time.time()
    with torch.no_grad():
        imgs, trace = ref_sample(gen, vq, num_samples=b, labels=labels, use_tqdm=False, **kw)
    dt = time.time() - t0
    hook.remove()
    out = dict(labels=labels.numpy(), tokens=torch.stack(trace).numpy().astype(np.int16), seconds=np.float64(dt),
               threads=np.int64(torch.get_num_threads()))
    if with_logits:
        # replay the RNG stream: per step exponentials first, then Gumbel
        torch.manual_seed(1234)
        v = rec[0].shape[-1]
        qs, gs = [], []
        for _ in range(steps):
            qs.append(torch.empty(b * 512, v).exponential_(1))
            gs.append(torch.distributions.Gumbel(0.0, 1.0).sample((b, 256, 2)))
        out.update(logits=torch.stack(rec).numpy(), q=torch.stack(qs).numpy(), g=torch.stack(gs).numpy())
    else:
        out.update(image0=imgs[0].numpy(), image_sub=imgs[:, :, ::4, ::4].contiguous().numpy())
    np.savez_compressed(path, **out)
    print(path, f"{dt:.1f}s", imgs.shape)


def golden_decode(vq, bits, path):
    g = torch.Generator().manual_seed(99)
    tok = torch.randint(0, 2 ** bits, (2, 256), generator=g)
    with torch.no_grad():
        img = vq.decode_tokens(tok)
    np.savez_compressed(path, tokens=tok.numpy().astype(np.int32), image0=img[0].numpy(),
                        image_sub=img[:, :, ::4, ::4].contiguous().numpy())
    print(path, img.shape, float(img.min()), float(img.max()))


def golden_encode(vq, path):
    """BASELINE config #4 shape at B=2: images in [0,1] (data/webdataset_reader.py:83) -> encoder latents, LFQ indices, reconstruction."""
    g = torch.Generator().manual_seed(4242)
    # smooth-ish random images: low-resolution noise upsampled, so that latents are not all near 0
    x = torch.nn.functional.interpolate(torch.rand((2, 3, 32, 32), generator=g), size=(256, 256), mode="bilinear", align_corners=False)
    x = (x + 0.1 * torch.rand((2, 3, 256, 256), generator=g)).clamp(0, 1)
    with torch.no_grad():
        z = vq.encoder(x)
        recon, d = vq(x)
    np.savez_compressed(path, x_seed=np.int64(4242), z=z.numpy(), indices=d["min_encoding_indices"].numpy().astype(np.int32),
                        recon0=recon[0].numpy(), recon_sub=recon[:, :, ::4, ::4].contiguous().numpy())
    print(path, z.shape, float(z.abs().min()), float(z.abs().mean()))


def golden_encode_input():
    g = torch.Generator().manual_seed(4242)
    x = torch.nn.functional.interpolate(torch.rand((2, 3, 32, 32), generator=g), size=(256, 256), mode="bilinear", align_corners=False)
    return (x + 0.1 * torch.rand((2, 3, 256, 256), generator=g)).clamp(0, 1)


def main():
    torch.set_num_threads(os.cpu_count())
    cfg, kw, vq, gen = build_reference(12)
    golden_forward(gen, 12, 4, os.path.join(HERE, "forward_12bit.npz"))
    golden_forward_prenorm(os.path.join(HERE, "forward_prenorm_12bit.npz"))
    golden_forward_bert(os.path.join(HERE, "forward_bert_12bit.npz"))
    golden_decode(vq, 12, os.path.join(HERE, "decode_12bit.npz"))
    golden_encode(vq, os.path.join(HERE, "encode_12bit.npz"))
    golden_sample(kw, vq, gen, 2, 4, os.path.join(HERE, "select_12bit.npz"), with_logits=True)
    golden_sample(kw, vq, gen, 4, 8, os.path.join(HERE, "sample_12bit.npz"), with_logits=False)
    del vq, gen
    cfg, kw, vq, gen = build_reference(14)
    golden_forward(gen, 14, 2, os.path.join(HERE, "forward_14bit.npz"))


if __name__ == "__main__":
    main()
