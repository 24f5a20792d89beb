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


@pytest.mark.parametrize("n,world", [(100_000, 8), (30_000, 2), (1000, 4), (130, 8), (64, 3), (1, 2), (100_000, 1)])
def test_paired_row_ranges(n, world):
    rg = sharding.paired_row_ranges(n, world)
    assert len(rg) == 2 * world and rg[0][0] == 0 and rg[-1][1] == n
    assert all(a <= b for a, b in rg) and all(rg[i][1] == rg[i + 1][0] for i in range(len(rg) - 1))
    assert all(a % 128 == 0 for a, _ in rg)
    owners = [sharding.range_owner(i, world) for i in range(2 * world)]
    assert sorted(owners) == sorted(list(range(world)) * 2)
    assert sum(sharding.upper_area(rg, r, world, n) for r in range(world)) == n * (n + 1) // 2
    if n >= 10_000:
        areas = [sharding.upper_area(rg, r, world, n) for r in range(world)]
        rows = [sum(rg[i][1] - rg[i][0] for i in sharding.owned_ranges(rg, r, world)) for r in range(world)]
        assert max(areas) / (n * (n + 1) / 2 / world) < 1.03       # equal compute
        assert max(rows) - min(rows) <= 256                        # equal rows (D2H volume)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim, align):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        X = rng.dirichlet(np.ones(dim), size=n)
        full = po.pairwise_np(X, "Eucl")
        ranges = sharding.paired_row_ranges(n, world, align=align)
        out_rows, T = {}, {}
        for i in sharding.owned_ranges(ranges, rank, world):
            a, b = ranges[i]
            # what the rank's two kernel launches per range produce: its rows from the diagonal block
            # rightwards (diagonal block mirrored locally) and the transposed off-diagonal blocks
            out_rows[i] = torch.full((b - a, n), float("nan"), dtype=torch.float64)
            out_rows[i][:, a:] = torch.from_numpy(full[a:b, a:])
            T[i] = torch.from_numpy(np.ascontiguousarray(full[a:b, b:].T))
        sharding.exchange_transposed(T, ranges, rank, world, out_rows)
        for i, rows in out_rows.items():
            a, b = ranges[i]
            assert np.array_equal(rows.numpy(), full[a:b]), "rank %d assembled a wrong block row %d" % (rank, i)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,align", [(2, 600, 128), (3, 1000, 128), (2, 64, 128), (3, 333, 16)])
def test_exchange_transposed_gloo(world, n, align):
    mp.spawn(_worker, args=(world, _free_port(), n, 16, align), nprocs=world, join=True)


@pytest.mark.parametrize("n,world", [(100_000, 8), (3000, 2), (130, 8), (1, 2)])
def test_range_offsets_stack_the_owned_ranges(n, world):
    rg = sharding.paired_row_ranges(n, world)
    off = sharding.range_offsets(rg, world)
    for rank in range(world):
        expect = 0
        for i in sharding.owned_ranges(rg, rank, world):
            assert off[i] == expect
            expect += rg[i][1] - rg[i][0]


def _block_rows_worker(rank, world, port, n, dim):
    """multigpu.BlockRows end to end under gloo: the tile launches are replaced by a CPU stand-in that
    fills `out` and the mirror target exactly as po_distance_block_ex addresses them."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from phyloligo_b200 import engine, multigpu
        from phyloligo_b200._lib import FLAG_MIRROR, FLAG_SKIP_LOWER

        rng = np.random.default_rng(5)
        X = rng.dirichlet(np.ones(dim), size=n)
        full = torch.from_numpy(po.pairwise_np(X, "Eucl"))
        calls = []

        def fake_block(metric, P, aux, d, row0, row1, col0, col1, out, out_row0, out_col0, flags=0, mirror=None,
                       mirror_row0=0, mirror_col0=0, mirror_ld=None):
            calls.append((row0, row1, col0, col1))
            blk = full[row0:row1, col0:col1]
            if flags & FLAG_SKIP_LOWER:  # (a row panel of) the diagonal block: upper triangle computed, mirrored in place
                assert row0 == col0 and row1 <= col1 and mirror is None and (flags & FLAG_MIRROR)
                out[col0 - out_row0:col1 - out_row0, row0 - out_col0:row1 - out_col0] = blk.T
            out[row0 - out_row0:row1 - out_row0, col0 - out_col0:col1 - out_col0] = blk
            if (flags & FLAG_MIRROR) and mirror is not None:
                assert isinstance(mirror, torch.Tensor)  # the NCCL-exchange path stages locally
                mirror[col0 - mirror_row0:col1 - mirror_row0, row0 - mirror_col0:row1 - mirror_col0] = blk.T

        engine.distance_block, keep = fake_block, engine.distance_block
        try:
            job = multigpu.BlockRows(n, torch.float64, rank, world, exchange="nccl", device=torch.device("cpu"))
            job.matrix.fill_(float("nan"))
            job.compute("Eucl", None, None, dim)
            first_pass = list(calls)
            # the command line's sink: every finished block is handed to `ship` exactly once, right parts of a
            # block row before the closing exchange, left parts after it; together they tile the rank's rows
            shipped = torch.full((n, n), float("nan"), dtype=torch.float64)
            order = []

            def ship(block, row0, col0):
                assert torch.isnan(shipped[row0:row0 + block.shape[0], col0:col0 + block.shape[1]]).all()
                shipped[row0:row0 + block.shape[0], col0:col0 + block.shape[1]] = block
                order.append((row0, col0))

            job.compute("Eucl", None, None, dim, ship=ship)
        finally:
            engine.distance_block = keep
        for i in job.my_ranges:
            a, b = job.ranges[i]
            assert torch.equal(shipped[a:b], full[a:b]), "rank %d shipped rows of block row %d" % (rank, i)
        owned = torch.zeros(n, dtype=torch.bool)
        for i in job.my_ranges:
            owned[job.ranges[i][0]:job.ranges[i][1]] = True
        assert torch.isnan(shipped[~owned]).all()
        rights = [k for k, (r0, c0) in enumerate(order) if c0 == r0]
        lefts = [k for k, (r0, c0) in enumerate(order) if c0 == 0 and r0 > 0]
        assert len(rights) == len(job.my_ranges) and (not lefts or max(rights) < min(lefts))
        for i, rows in job.out_rows.items():
            a, b = job.ranges[i]
            assert torch.equal(rows, full[a:b]), "rank %d block row %d" % (rank, i)
        # every unordered pair of blocks is computed by exactly one rank: upper-triangle area only
        assert calls[len(first_pass):] == first_pass  # shipping does not change what is launched
        calls = first_pass
        mine = sum((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in calls)
        diag = sum((r1 - r0) ** 2 for r0, r1, c0, c1 in calls if (r0, r1) == (c0, c1))
        assert mine - diag + (diag + sum(r1 - r0 for r0, r1, c0, c1 in calls if (r0, r1) == (c0, c1))) // 2 == job.upper_area()
        # upper triangle only across "PCIe": every rank hands the right parts of its block rows to a
        # MirroredHostSink over ONE matrix in shared memory and transposes them into the rows below; no
        # rank ships a left part, and together the ranks fill the whole matrix
        import tempfile
        from phyloligo_b200 import hostsink
        path = os.path.join(tempfile.gettempdir(), "po_test_shared_%d.mat" % port)
        if rank == 0:
            fm = hostsink.FileMatrix(path, n, n, np.float32, create=True)
            fm.array[:] = np.nan
        dist.barrier()
        if rank != 0:
            fm = hostsink.FileMatrix(path, n, n, np.float32, create=False)
        host = torch.from_numpy(fm.array)
        pool = engine.HostMirror(2)
        sink = multigpu.MirroredHostSink(host, pool, sub_rows=100)
        engine.distance_block = fake_block
        try:
            job.compute("Eucl", None, None, dim, ship=sink.ship, left_parts=False)
            sink.finish()
            dist.barrier()
            assert torch.equal(host, full.float()), "rank %d: shared host matrix" % rank
            tri = sum((job.ranges[i][1] - job.ranges[i][0]) * (n - job.ranges[i][0]) for i in job.my_ranges) * 4
            low = sum((job.ranges[i][1] - job.ranges[i][0]) * (n - job.ranges[i][1]) for i in job.my_ranges) * 4
            assert sink.dma_bytes == tri and sink.mirrored_bytes == low
            dist.barrier()
            # the same with the block rows launched and shipped in row panels: a panel's part right of its own
            # square is mirrored below it, inside the block row as well
            if rank == 0:
                fm.array[:] = np.nan
            job.matrix.fill_(float("nan"))
            dist.barrier()
            sink.reset()
            job.compute("Eucl", None, None, dim, ship=sink.ship, left_parts=False, panel_rows=128)
            sink.finish()
            dist.barrier()
            assert torch.equal(host, full.float()), "rank %d: shared host matrix, row panels" % rank
            for i, rows in job.out_rows.items():
                a, b = job.ranges[i]
                assert torch.equal(rows, full[a:b]), "rank %d block row %d, row panels" % (rank, i)
            # the lower halves of the diagonal blocks are mirrored too now: fewer bytes by "DMA", as many more by the host
            assert sink.dma_bytes <= tri and sink.dma_bytes + sink.mirrored_bytes == tri + low
            # row panels through the plain callback: right parts panel by panel, left parts at the end, no overlap
            shipped.fill_(float("nan"))
            del order[:]
            job.compute("Eucl", None, None, dim, ship=ship, panel_rows=256)
            for i in job.my_ranges:
                a, b = job.ranges[i]
                assert torch.equal(shipped[a:b], full[a:b]), "rank %d shipped rows of block row %d, row panels" % (rank, i)
        finally:
            engine.distance_block = keep
        dist.barrier()
        pool.close()
        del host, sink
        fm.close()
        if rank == 0:
            os.unlink(path)
        job.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 700), (3, 1000), (4, 2100)])
def test_block_rows_orchestration_gloo(world, n):
    mp.spawn(_block_rows_worker, args=(world, _free_port(), n, 12), nprocs=world, join=True)
