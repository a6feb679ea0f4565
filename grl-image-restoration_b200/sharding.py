"""Multi-GPU plumbing for tile/batch-sharded inference.

The hot path has no collective inside it: tiles are independent, weights are replicated, each rank (one process
per GPU) restores a contiguous slice of the batch.  The only exchange is the one the reference's metrics do at
`compute()` time -- torchmetrics states with dist_reduce_fx="cat" (utils/metrics/psnr.py:71-72) -- i.e. an
all-gather of per-image (index, value) pairs followed by de-duplication by index (average_metric, psnr.py:19-41).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous [lo, hi) slice of `n_items` for `rank`; earlier ranks take the remainder (sizes differ by <= 1)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_metric(values, indices, group=None):
    """All-gather per-image metric values and their dataset indices from every rank (uneven counts allowed),
    returns (values, indices) concatenated in rank order on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return values, indices
    world = dist.get_world_size(group)
    n = torch.tensor([values.numel()], device=values.device, dtype=torch.int64)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    m = int(max(c.item() for c in counts))
    pv = torch.zeros(m, device=values.device, dtype=values.dtype)
    pi = torch.full((m,), -1, device=values.device, dtype=torch.int64)
    pv[: values.numel()] = values.flatten()
    pi[: indices.numel()] = indices.flatten().to(torch.int64)
    gv = [torch.zeros_like(pv) for _ in range(world)]
    gi = [torch.zeros_like(pi) for _ in range(world)]
    dist.all_gather(gv, pv, group=group)
    dist.all_gather(gi, pi, group=group)
    vs = torch.cat([g[: int(c.item())] for g, c in zip(gv, counts)])
    ids = torch.cat([g[: int(c.item())] for g, c in zip(gi, counts)])
    return vs, ids


def average_metric(values, indices):
    """Mean over unique indices, first occurrence wins (utils/metrics/psnr.py:19-41: DistributedSampler pads the
    last batch with duplicates)."""
    seen, acc = set(), []
    for v, i in zip(values.tolist(), indices.tolist()):
        if i not in seen:
            seen.add(i)
            acc.append(v)
    return sum(acc) / max(len(acc), 1)
