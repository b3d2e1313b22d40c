"""Host-side schedule scalars of the sampler, computed with the same torch fp32 ops as the reference so that the
per-step tables handed to the device loop are bit-identical to what the reference computes on the fly."""
import math

import torch


def get_masking_ratio(progress: float, mode: str = "arccos") -> torch.Tensor:
    """modeling/modules/masking.py:41-65."""
    r = torch.tensor(progress)
    if mode == "root":
        val_to_mask = 1 - (r ** 0.5)
    elif mode == "square":
        val_to_mask = 1 - (r ** 2)
    elif mode == "cosine":
        val_to_mask = torch.cos(r * math.pi * 0.5)
    elif mode == "arccos":
        val_to_mask = torch.acos(r) / (math.pi * 0.5)
    elif mode == "linear":
        val_to_mask = 1 - r
    else:
        raise ValueError("Invalid mode. Choose between 'linear','square', 'cosine', 'arccos', 'root'.")
    return torch.clamp(val_to_mask, 1e-6, 1.0)


def guidance_scale_at(i, num_steps, guidance_scale, guidance_annealing, scale_pow):
    """modeling/modules/sampling.py:91-98; returns the value the reference multiplies (lc - lu) with, as fp32."""
    if guidance_annealing == "none":
        scale_step = 1.0
    elif guidance_annealing == "linear":
        scale_step = i / num_steps
    elif guidance_annealing == "cosine":
        sp = torch.ones((1)) * scale_pow
        scale_step = (1 - torch.cos(((i / num_steps) ** sp) * torch.pi)) * 1 / 2
    else:
        # the reference leaves scale_step undefined here (NameError); fail with a clear message instead
        raise ValueError(f"Invalid guidance_annealing {guidance_annealing!r}. Choose between 'none', 'linear', 'cosine'.")
    scale = guidance_scale * scale_step
    return float(torch.as_tensor(scale, dtype=torch.float32).reshape(-1)[0])


def step_tables(num_steps, num_maskable, *, softmax_temperature, mask_schedule_strategy, guidance_scale, guidance_annealing,
                scale_pow, use_sampling_annealing):
    """Per-step (scale, temperature, 1-progress, mask_len) as the reference computes them (sampling.py:82-124)."""
    scale, temp, omp, mask_len = [], [], [], []
    for i in range(num_steps):
        progress = (i + 1) / num_steps
        scale.append(guidance_scale_at(i, num_steps, guidance_scale, guidance_annealing, scale_pow) if guidance_scale != 0.0 else 0.0)
        temp.append(0.5 + 0.8 * (1 - progress) if use_sampling_annealing else softmax_temperature)
        omp.append(1 - progress)
        ratio = get_masking_ratio(progress, mode=mask_schedule_strategy)
        mask_len.append(float(torch.floor(ratio * num_maskable)))
    return scale, temp, omp, mask_len
