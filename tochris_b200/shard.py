"""Multi-GPU plumbing of ref_cuda: views are independent, so a batch is cut into
contiguous shards, one per rank (one process per GPU), the world is replicated, and the only
collective is the gather of finished framebuffers to the root rank.  torch.distributed is used for
that gather (NCCL over NVLink on the GPU box, gloo in the CPU tests); nothing here renders."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """views[lo:hi) of an n-view batch belong to `rank` (contiguous blocks, sizes differ by at most one)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_frames(local: torch.Tensor, n_total: int, rank: int, world: int, dst: int = 0, bufs=None, concat: bool = True):
    """Gathers the per-rank framebuffers [n_local, h, w] to `dst` in view order.  Returns the
    [n_total, h, w] tensor on `dst`, None elsewhere.  Shards may differ in size by one view:
    every rank pads to the largest shard so that one gather call moves everything.
    `bufs`: preallocated receive buffers on `dst` (a timed loop reuses them); `concat=False` returns
    the list of per-rank buffers instead of one tensor (no extra copy)."""
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    biggest = max(hi - lo for lo, hi in sizes)
    if local.shape[0] < biggest:
        pad = torch.zeros((biggest - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], 0)
    local = local.contiguous()
    if rank == dst and bufs is None:
        bufs = [torch.empty_like(local) for _ in range(world)]
    if rank != dst:
        bufs = None
    dist.gather(local, bufs, dst=dst)
    if rank != dst:
        return None
    if not concat:
        return bufs
    return torch.cat([bufs[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)], 0)
