"""CPU tests of the multi-GPU host logic with a real 2-process gloo group (the GPU path uses the same code over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from maskbit_b200.sharding import gather_images, rank_seed, shard_bounds, shard_labels


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        labels = torch.randint(0, 1000, (n_global,), generator=torch.Generator().manual_seed(1234))
        mine = shard_labels(labels)
        # stand-in for sample(): an "image" that encodes (label, position in the global batch)
        lo, hi = shard_bounds(n_global, rank, world)
        img = torch.zeros((hi - lo, 4, 4, 3), dtype=torch.uint8)
        img[:, 0, 0, 0] = (mine % 256).to(torch.uint8)
        img[:, 0, 0, 1] = torch.arange(lo, hi).to(torch.uint8)
        allimg = gather_images(img, n_global)
        assert allimg.shape == (n_global, 4, 4, 3)
        assert torch.equal(allimg[:, 0, 0, 0], (labels % 256).to(torch.uint8))
        assert torch.equal(allimg[:, 0, 0, 1], torch.arange(n_global).to(torch.uint8))
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([10.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert t.item() == 10.0 + world - 1
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_global", [8, 7])
def test_two_rank_shard_and_gather(tmp_path, n_global):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n_global, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_bounds_cover_and_seeds_differ():
    for n in (0, 1, 7, 256, 1024):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert len({rank_seed(7, r) for r in range(8)}) == 8
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)
    assert gather_images.__doc__
