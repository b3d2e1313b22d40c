"""maskbit_b200 -- B200-native (sm_100a) implementation of MaskBit's sampling hot path.

Drop-in for the reference's `sample` / `LFQBert` / `ConvVQModel.decode_tokens` (same names, arguments, checkpoint
layout); all compute runs in hand-written CUDA behind the C ABI of include/maskbit_b200.h.
"""
from .config import load_config, sampler_kwargs, derive_sampling_config  # noqa: F401
from .bert import Bert, LFQBert  # noqa: F401
from .conv_vqgan import ConvVQModel  # noqa: F401
from .sampling import sample  # noqa: F401
from .factorization import combine_factorized_tokens, split_factorized_tokens  # noqa: F401
from .masking import get_mask_tokens, get_masking_ratio  # noqa: F401
from .losses import MLMLoss  # noqa: F401


def build_models(config, device="cuda", generator_path=None, tokenizer_path=None):
    """Model construction of scripts/eval_maskbit.py:24-56,82-83 (get_tokenizer / get_generator).  With no checkpoint
    paths the deterministic synthetic weights of maskbit_b200.weights are used (there is no network for real ones)."""
    derive_sampling_config(config)
    tokenizer = ConvVQModel(config.model.vq_model, legacy=False)
    if tokenizer_path:
        tokenizer.load_pretrained(tokenizer_path)
    tokenizer.eval().requires_grad_(False)
    mlm = config.model.mlm_model
    if mlm.model_cls not in ("lfq_bert", "bert"):      # train_maskbit.py:128-131 knows exactly these two
        raise ValueError(f"model_cls {mlm.model_cls!r}: expected 'lfq_bert' or 'bert'")
    generator = (LFQBert if mlm.model_cls == "lfq_bert" else Bert)(
        img_size=config.dataset.preprocessing.resolution, hidden_dim=mlm.hidden_dim,
        codebook_size=config.model.vq_model.codebook_size, codebook_splits=mlm.codebook_splits, depth=mlm.depth,
        heads=mlm.heads, mlp_dim=mlm.mlp_dim, dropout=mlm.dropout, use_prenorm=mlm.use_prenorm,
        input_stride=2 ** (config.model.vq_model.num_resolutions - 1))
    if generator_path:
        generator.load_pretrained(generator_path, rename_keys={"token_emb": "input_proj"})
    generator.eval().requires_grad_(False)
    return tokenizer.to(device), generator.to(device)
