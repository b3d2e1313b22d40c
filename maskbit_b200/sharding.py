"""Multi-GPU sharding of the sampling path: one process per GPU, independent images, weights replicated.

The reference has no multi-GPU inference (scripts/eval_maskbit.py:65,183 takes one --device); its outer loop over label
batches (eval_maskbit.py:107-112) is the natural unit to shard: rank r of W takes the contiguous slice
[r*B/W, (r+1)*B/W) of every global batch, samples it with its own noise stream, and the finished uint8 images are collected
with ONE all-gather per batch (NCCL over NVLink on the GPU box, gloo in the CPU tests).  No collective runs inside the
decoding loop or the decoder.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous [lo, hi) slice of n items for `rank`; the first n % world ranks take one extra item."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_labels(labels, rank=None, world=None):
    """This rank's slice of a global label batch (eval_maskbit.py:112 `labels[i*bs:(i+1)*bs]`, split once more per rank)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(labels.shape[0], rank, world)
    return labels[lo:hi]


def rank_seed(seed, rank):
    """Per-rank Philox key for the device noise stream: ranks never share draws."""
    return (int(seed) * 0x9E3779B1 + 0x85EBCA6B * (rank + 1)) & ((1 << 62) - 1)


def gather_images(local, global_count=None, group=None):
    """All-gather the finished images ([b_r, H, W, 3] uint8 or [b_r, 3, H, W] float) of every rank into rank order.

    Equal shards use one all_gather_into_tensor (the single collective of the path); ragged shards (global batch not
    divisible by the world size) are padded to the largest shard and trimmed after the gather."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    rank = dist.get_rank(group)
    n = global_count if global_count is not None else local.shape[0] * world
    sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
    if sizes[rank] != local.shape[0]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} images, expected {sizes[rank]} of {n}")
    big = max(sizes)
    send = local
    if local.shape[0] != big:
        send = torch.zeros((big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    out = torch.empty((world * big,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send.contiguous(), group=group)
    if all(s == big for s in sizes):
        return out
    return torch.cat([out[r * big: r * big + sizes[r]] for r in range(world)], dim=0)
