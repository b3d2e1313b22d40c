"""Worker of tests/test_gpu_multi.py: launched by torchrun with one rank per GPU (NCCL).

Every rank samples ITS contiguous shard of a global batch with ITS slice of one globally drawn, injected noise stream
("compare each rank against the run on that rank's shard (labels + noise slices)"), the ranks all-gather tokens and
uint8 images over NCCL, and rank 0 re-runs every shard alone on its own GPU: tokens and images must be identical, in rank order."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maskbit_b200 import build_models, load_config, sample, sampler_kwargs  # noqa: E402
from maskbit_b200.sharding import gather_images, shard_bounds, shard_labels  # noqa: E402


def main():
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    cfg = load_config("maskbit_generator_12bit")
    steps, n_global = 6, 2 * world + 1                       # ragged: the first rank takes one image more
    kw = dict(sampler_kwargs(cfg), num_steps=steps)
    tokenizer, gen = build_models(cfg, device=dev)
    g = torch.Generator().manual_seed(77)
    labels = torch.randint(0, 1000, (n_global,), generator=g)
    q = torch.empty((steps, n_global, 512, 64)).exponential_(1, generator=g)
    gum = -torch.log(-torch.log(torch.rand((steps, n_global, 256, 2), generator=g).clamp_min(1e-20)))

    def run_shard(r):
        lo, hi = shard_bounds(n_global, r, world)
        img, trace = sample(gen, tokenizer, num_samples=hi - lo, labels=shard_labels(labels, r, world),
                            noise=(q[:, lo:hi].reshape(steps, (hi - lo) * 512, 64), gum[:, lo:hi]), **kw)
        return tokenizer.postprocess_uint8(img), torch.stack(trace, 1)          # [b, 256, 256, 3] uint8, [b, steps, 256, 2]

    u8, tok = run_shard(rank)
    all_u8 = gather_images(u8, n_global)
    all_tok = gather_images(tok, n_global)
    assert all_u8.shape == (n_global, 256, 256, 3) and all_tok.shape == (n_global, steps, 256, 2)
    if rank == 0:
        for r in range(world):
            lo, hi = shard_bounds(n_global, r, world)
            ref_u8, ref_tok = run_shard(r)
            assert torch.equal(all_tok[lo:hi], ref_tok), f"tokens of rank {r} differ from the single-GPU run of its shard"
            assert torch.equal(all_u8[lo:hi], ref_u8), f"images of rank {r} differ from the single-GPU run of its shard"
        print(f"multi-gpu parity ok: {world} ranks, {n_global} images, {steps} steps, tokens and uint8 images identical per shard")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
