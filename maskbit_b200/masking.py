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


def get_mask_tokens(tokens, mask_token, mode="arccos", min_masking_ratio=0.0, *, noise="reference_cpu", generator=None):
    """Drop-in mirror of modeling/modules/masking.py:7-38 (training-time random masking): returns (masked_tokens, mask).

    The per-sample ratio and the per-slot uniforms are drawn exactly like the reference draws them -- on the CPU, from torch's
    default generator (or `generator`), first ``rand(B)`` then ``rand(tokens.size())`` -- so the same seed masks the same slots;
    the compare-and-replace runs on the device (mb_mask_tokens).  ``noise="device"`` draws the per-slot uniforms with torch's CUDA
    generator instead (no host->device copy; not reproducible against the reference).  ``mask`` is returned on the tokens' device
    (the reference leaves it on the CPU, masking.py:35)."""
    import ctypes

    from . import _lib
    if tokens.device.type != "cuda":
        raise _lib.MaskbitError("get_mask_tokens runs on a CUDA device (tokens must be a CUDA tensor); there is no CPU fallback")
    b = tokens.size(0)
    r = torch.rand(b, generator=generator) * (1 - min_masking_ratio)
    if mode == "linear":
        val_to_mask = 1 - r
    elif mode == "square":
        val_to_mask = 1 - (r ** 2)
    elif mode == "cosine":
        val_to_mask = torch.cos(r * math.pi * 0.5)
    elif mode == "arccos":
        val_to_mask = torch.acos(r) / (math.pi * 0.5)
    else:
        raise ValueError("Invalid mode. Choose between 'linear','square', 'cosine', 'arccos'.")
    if noise == "reference_cpu":
        u = torch.rand(tokens.size(), generator=generator).to(tokens.device, non_blocking=True)
    elif noise == "device":
        u = torch.rand(tokens.size(), device=tokens.device)
    else:
        raise ValueError("noise must be 'reference_cpu' or 'device'")
    tok = tokens.detach().to(torch.int64).contiguous()
    val = val_to_mask.to(device=tokens.device, dtype=torch.float32).contiguous()
    masked = torch.empty_like(tok)
    mask = torch.empty(tok.shape, dtype=torch.bool, device=tok.device)
    with torch.cuda.device(tok.device):
        _lib.check(_lib.lib().mb_mask_tokens(ctypes.c_void_p(tok.data_ptr()), ctypes.c_void_p(u.data_ptr()), ctypes.c_void_p(val.data_ptr()),
                                             int(mask_token), ctypes.c_void_p(masked.data_ptr()), ctypes.c_void_p(mask.data_ptr()),
                                             b, tok.numel() // b, _lib.current_stream()))
    return masked, mask
