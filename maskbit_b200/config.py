"""YAML config loading for the sampling hot path.

Mirrors what scripts/eval_maskbit.py:67,74-80 does with OmegaConf: load the YAML, then derive
``codebook_size = 2**token_size`` and ``mask_token = 2**(log2(codebook_size)//codebook_splits)``.
OmegaConf is optional: ``yaml.safe_load`` plus an attribute dict with ``.get`` covers every key the
hot path reads (autoencoder.py:371 and conv_vqgan.py:23 use ``config.get``).
"""
import math
import os

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


class AttrDict(dict):
    """dict with attribute access and OmegaConf-style ``.get``."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return AttrDict({k: AttrDict.wrap(v) for k, v in obj.items()})
        if isinstance(obj, list):
            return [AttrDict.wrap(v) for v in obj]
        return obj


def load_config(path_or_name):
    """Load a generator YAML. ``path_or_name`` is a file path or a shipped name such as
    ``"maskbit_generator_12bit"``."""
    path = path_or_name
    if not os.path.isfile(path):
        cand = os.path.join(CONFIG_DIR, path_or_name if path_or_name.endswith(".yaml") else path_or_name + ".yaml")
        if not os.path.isfile(cand):
            raise ValueError(f"{path_or_name} does not exist")
        path = cand
    with open(path) as f:
        return AttrDict.wrap(yaml.safe_load(f))


def derive_sampling_config(config):
    """eval_maskbit.py:74-80: mutate the config with codebook_size and mask_token; returns (codebook_size, mask_token)."""
    vq = config.model.vq_model
    mlm = config.model.mlm_model
    codebook_size = 2 ** int(vq.token_size)
    vq.codebook_size = codebook_size
    splits = int(mlm.codebook_splits)
    mask_token = int(2 ** (int(math.log2(codebook_size)) // splits))
    mlm.mask_token = mask_token
    return codebook_size, mask_token


def sampler_kwargs(config, res=256):
    """The kwarg mapping of eval_maskbit.py:114-132 (everything but model/vqgan_model/num_samples/labels)."""
    codebook_size, mask_token = derive_sampling_config(config)
    mlm = config.model.mlm_model
    vq = config.model.vq_model
    return dict(
        softmax_temperature=mlm.softmax_temperature,
        randomize_temperature=mlm.randomize_temperature,
        mask_schedule_strategy=mlm.gen_mask_schedule_strategy,
        num_steps=mlm.num_steps,
        guidance_scale=mlm.guidance_scale,
        mask_token=mask_token,
        patch_size=res // (2 ** (vq.num_resolutions - 1)),
        guidance_annealing=mlm.guidance_annealing,
        use_sampling_annealing=mlm.use_sampling_annealing,
        scale_pow=mlm.scale_pow,
        codebook_size=codebook_size,
        codebook_splits=mlm.codebook_splits,
    )
