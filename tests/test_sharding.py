"""Host-side multi-GPU logic on CPU with the gloo backend, world_size 2 (the N>1 path of bench.py)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from _pkgload import load_package

    load_package()
    from grl_image_restoration_b200 import sharding

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_items = 7  # uneven on purpose
    lo, hi = sharding.shard_range(n_items, rank, world)
    idx = torch.arange(lo, hi)
    vals = idx.float() * 1.5 + 30.0  # stand-in for per-image PSNR
    if rank == 1:  # DistributedSampler-style duplicate of image 0 on the last rank
        idx = torch.cat([idx, torch.tensor([0])])
        vals = torch.cat([vals, torch.tensor([30.0])])
    gv, gi = sharding.gather_metric(vals, idx)
    mean = sharding.average_metric(gv, gi)
    if rank == 0:
        torch.save(dict(mean=mean, gi=gi, gv=gv, lohi=(lo, hi)), out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range(pkg):
    from grl_image_restoration_b200 import sharding

    for n in (0, 1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            parts = [sharding.shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_gather_world2_gloo(pkg, tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out)
    assert sorted(r["gi"].tolist()) == [0, 0, 1, 2, 3, 4, 5, 6]
    expect = sum(30.0 + 1.5 * i for i in range(7)) / 7
    assert abs(r["mean"] - expect) < 1e-6


def test_psnr_matches_oracle(pkg, oracle, golden_loader):
    from grl_image_restoration_b200 import metrics

    g = golden_loader("model_micro_cab_x2.npz")
    assert torch.equal(metrics.psnr(g["psnr/a"], g["psnr/b"], 4), g["psnr/value_border4"])
    assert torch.equal(metrics.psnr(g["psnr/a"], g["psnr/b"], 0), oracle.psnr(g["psnr/a"], g["psnr/b"], 0))


class _FakeModel:
    """Stand-in for GRL on the CPU: a per-tile function with a tile-global term (like the CAB pool), x2 upscale."""
    upscale, out_channels = 2, 3

    def __call__(self, patches):
        up = torch.nn.functional.interpolate(patches, scale_factor=2, mode="nearest")
        return up * 0.5 + patches.mean(dim=(1, 2, 3), keepdim=True)


def _tile_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    from _pkgload import load_package

    load_package()
    from grl_image_restoration_b200 import tiling

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x = torch.rand(1, 3, 72, 128, generator=torch.Generator().manual_seed(3))  # 2 x 3 = 6 tiles of 48 (overlap 8)
    y = tiling.forward_tile_sharded(_FakeModel(), x, 48, 8, max_batch=2)
    torch.save(y, out + f".{rank}")
    dist.barrier()
    dist.destroy_process_group()


def test_forward_tile_sharded_world2_gloo(pkg, tmp_path):
    """cfg5's N>1 path: tiles of one frame round-robin over ranks, all-gather, E / W on every rank == single-process."""
    from grl_image_restoration_b200 import tiling

    out = str(tmp_path / "tiles.pt")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_tile_worker, args=(2, port, out), nprocs=2, join=True)
    x = torch.rand(1, 3, 72, 128, generator=torch.Generator().manual_seed(3))
    ref = tiling.forward_tile(_FakeModel(), x, 48, 8, max_batch=4)
    assert tiling.tile_origins(72, 48, 8) == [0, 24] and tiling.tile_origins(128, 48, 8) == [0, 40, 80]
    assert tiling.shard_tiles(6, 0, 2) == [0, 2, 4] and tiling.shard_tiles(6, 1, 2) == [1, 3, 5]
    for r in range(2):
        y = torch.load(out + f".{r}")
        assert y.shape == (1, 3, 144, 256)
        assert torch.allclose(y, ref, atol=1e-6, rtol=0)
