"""Generation loop of scripts/eval_maskbit.py:58-137 (the caller of the hot path), minus the TensorFlow FID / IS evaluator.

    python -m maskbit_b200.eval_driver --config maskbit_generator_12bit --total-samples 50000 --batchsize 256 --out samples.npz
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 -m maskbit_b200.eval_driver ...   (one rank per GPU)

What is kept from the reference, line by line:
  * label schedule: ``randperm(1000).repeat(50)`` cut into consecutive batches (eval_maskbit.py:107-112) -- generalised to any
    ``total_samples`` (``repeat(ceil(total / 1000))[:total]``); drawn on the CPU from ``label_seed`` so every rank sees the same
    global schedule;
  * per batch ``sample(...)`` with the driver's kwarg mapping (eval_maskbit.py:114-132; ``softmax_temperature`` is the literal
    1.0 of line 118, not the YAML value);
  * post-processing ``clamp(0,1) * 255 -> NHWC -> uint8`` by truncation (eval_maskbit.py:134-135), here on the device
    (``mb_postprocess_u8``) before the copy to the host, so 4x fewer bytes cross PCIe / NVLink.
What is new: with ``torch.distributed`` initialised each global batch is split contiguously over the ranks
(``sharding.shard_labels``) and the finished uint8 images are collected with one all-gather per batch.
The output array [total_samples, H, W, 3] uint8 is what the reference hands to ``Evaluator.read_activations`` (and what the
ADM evaluation suite reads from an ``.npz`` as ``arr_0``).
"""
import argparse
import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import build_models, load_config, sample, sampler_kwargs
from .sharding import gather_images, rank_seed, shard_labels


def label_schedule(total_samples, nclass=1000, label_seed=0):
    """eval_maskbit.py:107-108: a random permutation of the classes repeated, so every batch mixes classes."""
    g = torch.Generator().manual_seed(int(label_seed))
    perm = torch.randperm(nclass, dtype=torch.int32, generator=g)
    return perm.repeat(math.ceil(total_samples / nclass))[:total_samples].long()


@torch.no_grad()
def generate_samples(config, total_samples=50_000, batchsize=100, res=256, tokenizer_path="", generator_path="", device="cuda:0",
                     label_seed=0, noise_seed=0, models=None, progress=False):
    """Returns (images uint8 numpy [total_samples, res, res, 3] on rank 0 -- None on the other ranks --, labels int64 tensor
    [total_samples]).  Every batch is all-gathered over NCCL (the path's one collective); only rank 0 stages it to the host, so
    host memory and PCIe traffic do not grow with the number of ranks (50 000 images are 9.8 GB).

    ``config`` is a YAML path / shipped config name / loaded config.  ``batchsize`` is the GLOBAL batch of one sample() round
    (the reference's ``batchsize``); with W ranks each samples ``batchsize / W`` of it.  The last batch may be ragged."""
    if isinstance(config, str):
        config = load_config(config)
    if res != 256:
        raise ValueError("res must be 256 (the shipped tokenizers are 256x256; eval_maskbit.py:137-142 also lists 512)")
    kw = dict(sampler_kwargs(config, res=res), softmax_temperature=1.0)          # eval_maskbit.py:118
    tokenizer, generator = models if models is not None else build_models(config, device=device, generator_path=generator_path or None,
                                                                            tokenizer_path=tokenizer_path or None)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    labels = label_schedule(total_samples, label_seed=label_seed)
    keep = rank == 0
    out = np.empty((total_samples, res, res, 3), dtype=np.uint8) if keep else None
    staging = torch.empty((min(batchsize, total_samples), res, res, 3), dtype=torch.uint8).pin_memory() if keep else None
    n_batches = math.ceil(total_samples / batchsize)
    for i in range(n_batches):
        y = labels[batchsize * i: batchsize * (i + 1)]
        nb = y.shape[0]
        y_local = shard_labels(y, rank, world) if world > 1 else y
        if y_local.shape[0] > 0:
            imgs, _ = sample(generator, tokenizer, num_samples=y_local.shape[0], labels=y_local, noise="device",
                             seed=rank_seed(noise_seed + i, rank), return_trace=False, skip_zero_scale_uncond=True, **kw)
            u8 = tokenizer.postprocess_uint8(imgs)
        else:
            u8 = torch.empty((0, res, res, 3), dtype=torch.uint8, device=generator.device)
        if world > 1:
            u8 = gather_images(u8, nb)
        if keep:
            staging[:nb].copy_(u8, non_blocking=True)
            torch.cuda.current_stream(u8.device).synchronize()
            out[batchsize * i: batchsize * i + nb] = staging[:nb].numpy()
        if progress and rank == 0:
            print(f"batch {i + 1}/{n_batches}", flush=True)
    return out, labels


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--config", default="maskbit_generator_12bit")
    ap.add_argument("--total-samples", type=int, default=50_000)
    ap.add_argument("--batchsize", type=int, default=256, help="global batch per sample() round (split over the ranks)")
    ap.add_argument("--res", type=int, default=256)
    ap.add_argument("--tokenizer-path", default="")
    ap.add_argument("--generator-path", default="")
    ap.add_argument("--label-seed", type=int, default=0)
    ap.add_argument("--noise-seed", type=int, default=0)
    ap.add_argument("--out", default="", help=".npz written by rank 0 (arr_0 = images uint8 NHWC, labels)")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    t0 = time.perf_counter()
    imgs, labels = generate_samples(a.config, a.total_samples, a.batchsize, a.res, a.tokenizer_path, a.generator_path,
                                    device=f"cuda:{local}", label_seed=a.label_seed, noise_seed=a.noise_seed, progress=True)
    dt = time.perf_counter() - t0
    if (not dist.is_initialized()) or dist.get_rank() == 0:
        print(f"{imgs.shape[0]} images in {dt:.1f} s ({imgs.shape[0] / dt:.1f} images/s incl. model load) on {world} GPU(s)")
        if a.out:
            np.savez(a.out, arr_0=imgs, labels=labels.numpy())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
