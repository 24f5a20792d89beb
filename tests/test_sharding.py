"""Host-side multi-rank logic on CPU (gloo, world_size 2 and 3): record cuts, triangle-balanced
block rows, and the exchange of transposed off-diagonal blocks.  The per-block numbers come
from the oracle here -- the point is the partition and the assembly, not the kernel."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import phylo_oracle as po
from phyloligo_b200 import sharding


def test_record_cuts_balance_and_cover():
    rng = np.random.default_rng(0)
    lengths = rng.integers(100, 50_000, size=1000)
    for world in (1, 2, 3, 8):
        cuts = sharding.record_cuts(lengths, world)
        assert cuts[0] == 0 and cuts[-1] == 1000 and all(b >= a for a, b in zip(cuts, cuts[1:]))
        loads = [lengths[a:b].sum() for a, b in zip(cuts, cuts[1:])]
        assert max(loads) - min(loads) <= 2 * lengths.max()
    assert sharding.record_cuts([], 4) == [0, 0, 0, 0, 0]


@pytest.mark.parametrize("n,world", [(100_000, 8), (30_000, 2), (1000, 4), (130, 8), (64, 3), (1, 2)])
def test_triangle_row_ranges(n, world):
    b = sharding.triangle_row_ranges(n, world)
    assert b[0] == 0 and b[-1] == n and len(b) == world + 1
    assert all(y >= x for x, y in zip(b, b[1:]))
    assert all(x % 128 == 0 for x in b[1:-1])
    assert sum(sharding.upper_area(b, s, n) for s in range(world)) == n * (n + 1) // 2
    if n >= 10_000:
        areas = [sharding.upper_area(b, s, n) for s in range(world)]
        assert max(areas) / (n * (n + 1) / 2 / world) < 1.05


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        X = rng.dirichlet(np.ones(dim), size=n)
        full = po.pairwise_np(X, "Eucl")
        bounds = sharding.triangle_row_ranges(n, world)
        a, b = bounds[rank], bounds[rank + 1]
        rows = b - a
        out_rows = torch.full((rows, n), float("nan"), dtype=torch.float64)
        # what the rank's two kernel launches produce: its rows right of its first row
        # (diagonal block mirrored locally) and the transposed off-diagonal blocks
        out_rows[:, a:] = torch.from_numpy(full[a:b, a:])
        T = torch.from_numpy(np.ascontiguousarray(full[a:b, b:].T))
        sharding.exchange_transposed(T, bounds, rank, world, out_rows)
        assert np.array_equal(out_rows.numpy(), full[a:b]), "rank %d assembled a wrong block row" % rank
        # every rank's rows together are the whole matrix
        sizes = [bounds[s + 1] - bounds[s] for s in range(world)]
        gathered = [torch.empty((sizes[s], n), dtype=torch.float64) for s in range(world)]
        dist.all_gather(gathered, out_rows) if len(set(sizes)) == 1 else None
        if len(set(sizes)) == 1:
            assert np.array_equal(torch.cat(gathered).numpy(), full)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 600), (3, 1000), (2, 64)])
def test_exchange_transposed_gloo(world, n):
    mp.spawn(_worker, args=(world, _free_port(), n, 16), nprocs=world, join=True)
