"""Multi-GPU sharding of a centroid list (the reference's fan-out is the fork/Queue
pool of pycurious/optimise.py:386-444).

Windows are independent, so there is no collective on the data path: every worker owns a contiguous block of
the centroid list (concatenating the blocks keeps the reference's output order), uploads only the grid rows its
windows cover, and the per-window results ((n,4) fits or (n,10) moments) are brought together by ONE collective.
Two ways to get the workers:

  * one process per GPU under `torch.distributed` (torchrun): `CurieOptimise.optimise_routine`,
    `parallelise_routine`, `metropolis_hastings_batch`, `sensitivity_batch` and `curie_depth_maps` shard by
    themselves when a process group with more than one rank is initialised (every rank calls them with the
    same arguments and every rank gets the full result: one NCCL all-gather over NVLink); the `*_sharded`
    functions below are the explicit forms with a destination rank (one NCCL gather);
  * one process, several GPUs: `CurieOptimise(..., max_processors=N)` -- the reference's own knob
    (optimise.py:116) -- fans out over min(N, visible GPUs) host threads, one per device (`fan_out_devices`).

The same functions run with the gloo backend on CPU tensors, which is how the host logic is tested without
GPUs (tests/test_distributed_gloo.py).
"""
from __future__ import annotations

import threading

import numpy as np


def shard_range(n, rank, world):
    """Contiguous block [lo, hi) of rank `rank`: floor(r*n/R) .. floor((r+1)*n/R) keeps the
    reference's output order when blocks are concatenated."""
    return (rank * n) // world, ((rank + 1) * n) // world


def is_distributed():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


_gather_cache = {}


def _gather_plan(n_total, world, tail, dtype, device):
    """Preallocated receive buffer [world, nmax, ...] and, when the blocks are not all the same size, the row
    index that compacts it into reference order -- built once per (n, world, shape, device)."""
    import torch
    key = (int(n_total), int(world), tuple(tail), dtype, str(device))
    hit = _gather_cache.get(key)
    if hit is None:
        if len(_gather_cache) > 16:
            _gather_cache.clear()
        sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
        nmax = max(sizes) if sizes else 0
        buf = torch.empty((world, nmax) + tuple(tail), dtype=dtype, device=device)
        pick = None
        if any(s != nmax for s in sizes):
            pick = torch.from_numpy(np.concatenate([r * nmax + np.arange(s, dtype=np.int64)
                                                    for r, s in enumerate(sizes)])).to(device)
        hit = (buf, pick, nmax)
        _gather_cache[key] = hit
    return hit


def gather_blocks(local, n_total, dst=0, group=None):
    """Bring the ranks' (n_r, ...) blocks together as one (n_total, ...) tensor on rank `dst` (one gather) or on
    every rank (dst=None: one all-gather into a preallocated tensor).  Blocks are padded to the largest block
    (they differ by at most one row).  Returns None on the ranks that receive nothing."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    tail = tuple(local.shape[1:])
    buf, pick, nmax = _gather_plan(n_total, world, tail, local.dtype, local.device)
    if local.shape[0] == nmax:
        send = local.contiguous()
    else:
        send = buf.new_empty((nmax,) + tail)
        send[: local.shape[0]] = local
        send[local.shape[0]:] = 0
    if dst is None:
        dist.all_gather_into_tensor(buf.view((world * nmax,) + tail), send, group=group)
    else:
        dist.gather(send, list(buf.unbind(0)) if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    flat = buf.view((world * nmax,) + tail)
    if pick is None:
        return flat.clone()                    # the receive buffer is reused by the next call
    return flat.index_select(0, pick)


def device_workers(max_processors, n_items, min_per_worker):
    """How many GPUs of this process a batch of n_items fans out over: min(max_processors, visible GPUs), but
    never so many that a device gets fewer than `min_per_worker` items.  1 inside torch.distributed (one
    process per GPU there)."""
    import torch
    if is_distributed() or not torch.cuda.is_available():
        return 1
    ndev = min(int(max_processors), torch.cuda.device_count())
    return max(1, min(ndev, int(n_items) // max(int(min_per_worker), 1)))


def fan_out_devices(ndev, n_items, work):
    """Run `work(lo, hi, device_index)` for the ndev contiguous blocks of range(n_items), one host thread per
    device (the calls release the GIL inside CUDA / ctypes).  A call either places its block itself and returns
    None, or returns a host array (hi-lo, ...); returned arrays are concatenated in block order.  The first
    worker exception is re-raised."""
    import torch
    out, err = [None] * ndev, []

    def run(r):
        try:
            lo, hi = shard_range(n_items, r, ndev)
            with torch.cuda.device(r):
                out[r] = work(lo, hi, r)
        except BaseException as e:          # noqa: BLE001  (re-raised in the caller's thread)
            err.append(e)

    threads = [threading.Thread(target=run, args=(r,), name="curie-b200-dev%d" % r) for r in range(ndev)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if err:
        raise err[0]
    if any(o is None for o in out):
        return None
    return np.concatenate(out, axis=0)


def optimise_routine_sharded(grid, window, xc_list, yc_list, dst=0, **kwargs):
    """`grid.optimise_routine` over all ranks: each rank fits its block, rank `dst` receives
    [beta(n,), zt(n,), dz(n,), C(n,)] (None elsewhere; every rank if dst is None)."""
    import torch.distributed as dist
    n = len(xc_list)
    lo, hi = shard_range(n, dist.get_rank(), dist.get_world_size())
    local = grid.optimise_batch(window, np.asarray(xc_list)[lo:hi], np.asarray(yc_list)[lo:hi],
                                as_numpy=False, shard=False, **kwargs)
    full = gather_blocks(local, n, dst=dst)
    if full is None:
        return None
    return list(grid._to_host(full.t()))                    # transposed on the device: four (n,) rows


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
                                           as_numpy=False, shard=False, **kwargs)
    full = gather_blocks(local, n, dst=dst)
    return None if full is None else grid._to_host(full)
