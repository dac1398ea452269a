"""Multi-GPU sharding of a centroid list: one process per GPU (torch.distributed), every rank holds
the whole grid (4096^2 float64 = 134 MB; the EMAG2-shaped 21600 x 10800 grid = 1.87 GB) and owns a
contiguous block of windows; the per-window results ((n,4) fits or (n,10) moments) are brought
together by a single gather over NCCL (NVLink).  Windows are independent: there is no collective on
the data path.  The same functions run with the gloo backend on CPU tensors,
which is how the host logic is tested without GPUs.
"""
from __future__ import annotations

import numpy as np


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: floor(r*n/R) .. floor((r+1)*n/R) keeps the
    reference's output order when blocks are concatenated."""
    return (rank * n) // world, ((rank + 1) * n) // world


def is_distributed():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def gather_blocks(local, n_total, dst=0, group=None):
    """Gather the ranks' (n_r, ...) blocks into an (n_total, ...) tensor on rank `dst`
    (all ranks when dst is None).  Blocks are padded to the largest block so that one
    fixed-size gather suffices.  Returns None on the other ranks."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nmax = max(shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world))
    tail = tuple(local.shape[1:])
    pad = torch.zeros((nmax,) + tail, dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if dst is None:
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
    else:
        bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, bufs, dst=dst, group=group)
        if rank != dst:
            return None
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts, dim=0)


def optimise_routine_sharded(grid, window, xc_list, yc_list, dst=0, **kwargs):
    """`grid.optimise_routine` over all ranks: each rank fits its block, rank `dst` receives
    [beta(n,), zt(n,), dz(n,), C(n,)] (None elsewhere; every rank if dst is None)."""
    import torch.distributed as dist
    n = len(xc_list)
    lo, hi = shard_range(n, dist.get_rank(), dist.get_world_size())
    local = grid.optimise_batch(window, np.asarray(xc_list)[lo:hi], np.asarray(yc_list)[lo:hi],
                                as_numpy=False, **kwargs)
    full = gather_blocks(local, n, dst=dst)
    if full is None:
        return None
    return list(full.t().contiguous().cpu().numpy())           # transposed on the device: four (n,) rows


def metropolis_hastings_sharded(grid, window, xc_list, yc_list, nsim, burnin, dst=0, **kwargs):
    """Posterior moments (n, 10) for every centroid, chains sharded over the ranks.  The chain id of
    a window is its global index, so the result does not depend on the number of ranks."""
    import torch.distributed as dist
    n = len(xc_list)
    lo, hi = shard_range(n, dist.get_rank(), dist.get_world_size())
    x0 = kwargs.pop("x0", None)
    if x0 is not None:
        x0 = x0[lo:hi]
    local = grid.metropolis_hastings_batch(window, np.asarray(xc_list)[lo:hi], np.asarray(yc_list)[lo:hi],
                                           nsim, burnin, reduce="moments", chain_offset=lo, x0=x0,
                                           as_numpy=False, **kwargs)
    full = gather_blocks(local, n, dst=dst)
    return None if full is None else full.cpu().numpy()
