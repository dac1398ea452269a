"""N > 1 host logic on CPU: contiguous block sharding + the single gather, world_size 2 over gloo."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from pycurious_b200 import distributed as pd
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    lo, hi = pd.shard_range(n, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([1.0, 10.0, 100.0, 1000.0], dtype=torch.float64)
    full = pd.gather_blocks(local, n, dst=0)
    allr = pd.gather_blocks(local, n, dst=None)
    q.put((rank, None if full is None else full.numpy(), allr.numpy(), pd.is_distributed()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 152100 // 50])
def test_shard_and_gather_world2(n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        r, full, allr, flag = q.get(timeout=120)
        got[r] = (full, allr, flag)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(n, dtype=np.float64)[:, None] * np.array([1.0, 10.0, 100.0, 1000.0])
    assert got[1][0] is None and np.array_equal(got[0][0], want)
    assert np.array_equal(got[0][1], want) and np.array_equal(got[1][1], want)
    assert got[0][2] and got[1][2]


def test_shard_ranges_cover_and_keep_order():
    from pycurious_b200.distributed import shard_range
    for n in (0, 1, 7, 152100, 246016):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b[1] - b[0] for b in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker_run_blocks(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import torch.distributed as dist
    import pycurious_b200 as pc
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    G = pc.CurieOptimise(np.zeros((16, 16)), 0.0, 15e3, 0.0, 15e3)
    seen = []

    def block(o, lo, hi):
        seen.append((lo, hi))
        idx = torch.arange(lo, hi, dtype=torch.float64)
        return torch.stack([idx, 10.0 * idx, -idx, idx * idx], dim=1)

    full = G._run_blocks(n, block)
    rows = G._run_blocks(n, block, transpose=True)
    alone = G._run_blocks(n, block, shard=False)
    np.random.seed(100 + rank)                       # different global RNG state on every rank
    seed = G._shared_seed(None)
    q.put((rank, full, rows, alone, seen, seed, G._shared_seed(12345)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [9, 3042])
def test_run_blocks_shards_over_ranks_world2(n):
    """CurieOptimise._run_blocks inside a process group: each rank evaluates only its contiguous block, every
    rank gets the full result in reference order; a seed drawn from numpy's global RNG is rank 0's on all ranks."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_run_blocks, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        item = q.get(timeout=120)
        got[item[0]] = item[1:]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    idx = np.arange(n, dtype=np.float64)
    want = np.stack([idx, 10.0 * idx, -idx, idx * idx], axis=1)
    from pycurious_b200.distributed import shard_range
    for r in range(2):
        full, rows, alone, seen, seed, fixed = got[r]
        assert np.array_equal(full, want) and np.array_equal(rows, want.T) and np.array_equal(alone, want)
        lo, hi = shard_range(n, r, 2)
        assert seen == [(lo, hi), (lo, hi), (0, n)]
        assert fixed == 12345
    assert got[0][4] == got[1][4]
