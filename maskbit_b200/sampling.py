"""sample() -- drop-in mirror of modeling/modules/sampling.py:12-136.

Same signature and return value ((images fp32 [B,3,H,W], list of num_steps int64 [B,n,m] tensors)).  The step loop
(CFG double-batch forward, softmax / categorical sample, Gumbel confidence, k-th-smallest re-mask) runs resident on the
device inside libmaskbit_b200 (mb_sample): no host synchronisation and no host<->device copies between steps, where the
reference does two H2D copies and ~20 small launches per step.

``noise`` selects where the per-step random draws come from (extension, keyword-only):
  "device"        (default) Philox4x32-10 on the device, keyed by ``seed`` (taken from torch's default generator when None)
  "reference_cpu" the reference's own draws when it runs on CPU: per step, first ``B*n*m*V`` exponentials, then the Gumbel
                  uniforms, both from torch's default CPU generator -- used by the parity tests
  (q, g) tensors  explicit draws, shapes [num_steps, B*n*m, V] and [num_steps, B, n, m]
"""
import ctypes
from typing import List, Optional, Text, Tuple

import torch

from . import _lib
from .masking import step_tables


def _draw_reference_cpu_noise(num_steps, b, n, m, v):
    qs, gs = [], []
    gumbel = torch.distributions.Gumbel(loc=0.0, scale=1.0)
    for _ in range(num_steps):
        qs.append(torch.empty(b * n * m, v).exponential_(1))
        gs.append(gumbel.sample((b, n, m)))
    return torch.stack(qs), torch.stack(gs)


@torch.no_grad()
def sample(
    model,
    vqgan_model,
    num_samples: int = 10,
    labels: Optional[torch.Tensor] = None,
    softmax_temperature: float = 1.0,
    randomize_temperature: float = 4.5,
    mask_schedule_strategy: Text = "linear",
    num_steps: int = 12,
    guidance_scale: float = 3.0,
    mask_token: int = 1024,
    patch_size: int = 16,
    guidance_annealing: Text = "none",
    use_sampling_annealing: bool = False,
    scale_pow: float = 4.0,
    codebook_size: int = 1024,
    codebook_splits: int = 1,
    use_tqdm: bool = False,
    *,
    noise="device",
    seed: Optional[int] = None,
    skip_zero_scale_uncond: bool = False,
    return_trace: bool = True,
) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    device = model.device
    model.eval()
    vqgan_model.eval()
    if labels is None:
        # sampling.py:60-63
        labels = [1, 7, 282, 604, 724, 179, 751, 404, 850, torch.randint(0, 999, size=(1,))] * (num_samples // 10)
        labels = torch.LongTensor(labels)
    n = int(patch_size ** 2)
    m = int(codebook_splits)
    if n != model.seq_len or m != model.splits or int(mask_token) != model.mask_token or int(codebook_size) != model.codebook_size:
        raise ValueError(f"sampler arguments (patch_size={patch_size}, codebook_splits={codebook_splits}, mask_token={mask_token}, "
                         f"codebook_size={codebook_size}) do not match the generator (seq_len={model.seq_len}, splits={model.splits}, "
                         f"mask_token={model.mask_token}, codebook_size={model.codebook_size})")
    labels = labels.to(device=device, dtype=torch.int64).contiguous().view(-1)
    if labels.numel() != num_samples:
        raise ValueError(f"labels has {labels.numel()} entries for num_samples={num_samples}")
    v = model.effective_codebook_size
    scale, temp, omp, mask_len = step_tables(
        num_steps, n * m, softmax_temperature=softmax_temperature, mask_schedule_strategy=mask_schedule_strategy,
        guidance_scale=guidance_scale, guidance_annealing=guidance_annealing, scale_pow=scale_pow,
        use_sampling_annealing=use_sampling_annealing)

    q_dev = g_dev = None
    if isinstance(noise, str):
        if noise == "reference_cpu":
            q, g = _draw_reference_cpu_noise(num_steps, num_samples, n, m, v)
            q_dev, g_dev = q.to(device), g.to(device)
        elif noise != "device":
            raise ValueError("noise must be 'device', 'reference_cpu' or a (q, g) tuple")
    else:
        q, g = noise
        q_dev = q.to(device=device, dtype=torch.float32).contiguous()
        g_dev = g.to(device=device, dtype=torch.float32).contiguous()
        if tuple(q_dev.shape) != (num_steps, num_samples * n * m, v) or g_dev.numel() != num_steps * num_samples * n * m:
            raise ValueError("injected noise has the wrong shape")
    if seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if q_dev is None else 0

    h = model._engine()
    FA = ctypes.c_float * num_steps
    args = _lib.MBSampleArgs()
    tables = [FA(*scale), FA(*temp), FA(*omp), FA(*mask_len)]
    args.labels = labels.data_ptr()
    args.B, args.num_steps = num_samples, num_steps
    args.use_guidance = int(guidance_scale != 0.0)
    args.skip_zero_scale_uncond = int(bool(skip_zero_scale_uncond))
    args.scale, args.temperature, args.one_minus_progress, args.mask_len = tables
    args.randomize_temperature = randomize_temperature
    args.q = q_dev.data_ptr() if q_dev is not None else None
    args.gumbel = g_dev.data_ptr() if g_dev is not None else None
    args.seed = seed
    with torch.cuda.device(device):
        trace = torch.empty((num_steps, num_samples, n, m), dtype=torch.int64, device=device) if return_trace else None
        final_tokens = torch.empty((num_samples, n), dtype=torch.int64, device=device)
        args.images = None                      # decode through the tokenizer model's own handle below
        args.trace = trace.data_ptr() if trace is not None else None
        args.final_tokens = final_tokens.data_ptr()
        _lib.check(_lib.lib().mb_sample(h, ctypes.byref(args), _lib.current_stream()))
    generated_image = vqgan_model.decode_tokens(final_tokens)     # sampling.py:133-135
    l_full_tokens = list(trace.unbind(0)) if trace is not None else []
    return generated_image, l_full_tokens
