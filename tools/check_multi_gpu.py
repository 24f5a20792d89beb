#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_multi_gpu.py : every rank computes its paired block rows
of a distance matrix (diagonal block mirrored locally, off-diagonal blocks exchanged transposed) and compares it bit
for bit with the same rows of a single-GPU run of the same kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from phyloligo_b200 import engine, sharding
from phyloligo_b200._lib import FLAG_MIRROR, FLAG_SKIP_LOWER

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for metric, n, dim in (("JSD", 3000, 256), ("Eucl", 1111, 64), ("EuclGram", 1500, 512), ("BC", 777, 256)):
    rng = np.random.default_rng(11)
    X = torch.from_numpy(rng.dirichlet(np.ones(dim), size=n).astype(np.float32)).cuda()
    full = engine.distance_matrix_device(X, metric, torch.float32, symmetric=True)
    ranges = sharding.paired_row_ranges(n, world)
    P, aux, d = engine.prepare(X, metric)
    rows, T = {}, {}
    for i in sharding.owned_ranges(ranges, rank, world):
        a, b = ranges[i]
        if b <= a:
            continue
        rows[i] = torch.full((b - a, n), float("nan"), dtype=torch.float32, device="cuda")
        T[i] = torch.empty((max(1, n - b), b - a), dtype=torch.float32, device="cuda")
        engine.distance_block(metric, P, aux, d, a, b, a, b, rows[i], a, 0, FLAG_SKIP_LOWER | FLAG_MIRROR)
        if b < n:
            engine.distance_block(metric, P, aux, d, a, b, b, n, rows[i], a, 0, FLAG_MIRROR, mirror=T[i], mirror_row0=b, mirror_col0=a)
    sharding.exchange_transposed(T, ranges, rank, world, rows)
    ok = all(torch.equal(rows[i], full[ranges[i][0]:ranges[i][1]]) for i in rows)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("%s n=%d world=%d rows identical to the single-GPU matrix: %s" % (metric, n, world, bool(flag.item())))
    assert ok, "rank %d rows differ" % rank
dist.destroy_process_group()
