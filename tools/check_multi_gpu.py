#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_multi_gpu.py : every rank computes its paired block rows
of a distance matrix (multigpu.BlockRows: diagonal blocks mirrored locally, off-diagonal tiles stored
transposed into the owner's rows -- over NVLink peer memory and, for comparison, through the NCCL
exchange) and compares them bit for bit with the same rows of a single-GPU run of the same kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from phyloligo_b200 import engine, hostsink, multigpu

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
for exchange in ("peer", "nccl"):
    for metric, n, dim in (("JSD", 3000, 256), ("Eucl", 1111, 64), ("EuclGram", 1500, 512), ("BC", 777, 256),
                           ("SC", 900, 256), ("KT", 500, 64)):
        rng = np.random.default_rng(11)
        X = torch.from_numpy(rng.dirichlet(np.ones(dim), size=n).astype(np.float32)).cuda()
        full = engine.distance_matrix_device(X, metric, torch.float32, symmetric=True)
        P, aux, d = engine.prepare(X, metric)
        job = multigpu.BlockRows(n, torch.float32, rank, world, exchange=exchange)
        job.matrix.fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        host = torch.full(tuple(job.matrix.shape), -1.0, dtype=torch.float32).pin_memory()
        for it in range(2):  # twice: the second pass overwrites live rows while peers may still read them
            job.compute(metric, P, aux, d, host_rows=host if it else None)  # second pass also ships the rows to the host
        torch.cuda.synchronize()
        ok = all(torch.equal(job.out_rows[i], full[job.ranges[i][0]:job.ranges[i][1]]) for i in job.out_rows)
        ok = ok and (job.rows_owned == 0 or torch.equal(host[:job.rows_owned], job.matrix[:job.rows_owned].cpu()))
        # third pass: block rows launched and shipped in row panels, only the upper triangle crosses PCIe into ONE
        # matrix in shared memory (own rows page-locked), the rest is mirrored there by the rank that shipped it
        path = "/dev/shm/po_check_%s_%s_%s.mat" % (os.environ.get("MASTER_PORT", "0"), exchange, metric)
        if rank == 0:
            fm = hostsink.FileMatrix(path, n, n, np.float32, create=True)
            fm.array[:] = np.nan
        dist.barrier()
        if rank != 0:
            fm = hostsink.FileMatrix(path, n, n, np.float32, create=False)
        registered = fm.register_rows([job.ranges[i] for i in job.my_ranges])
        shared = torch.from_numpy(fm.array)
        pool = engine.HostMirror(3)
        sink = multigpu.MirroredHostSink(shared, pool, sub_rows=128)
        job.matrix.fill_(float("nan"))
        torch.cuda.synchronize()
        dist.barrier()
        job.compute(metric, P, aux, d, ship=sink.ship, left_parts=False, panel_rows=256)
        sink.finish()
        torch.cuda.synchronize()
        dist.barrier()
        ok = ok and all(torch.equal(job.out_rows[i], full[job.ranges[i][0]:job.ranges[i][1]]) for i in job.out_rows)
        ok3 = torch.equal(shared, full.cpu())
        if not ok3:
            print("rank %d: shared host matrix differs (%s, %s; rows registered: %s)" % (rank, exchange, metric, registered), flush=True)
        ok = ok and ok3
        dist.barrier()
        pool.close()
        del shared, sink
        fm.close()
        if rank == 0:
            os.unlink(path)
        flag = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            used = job.exchange if job.exchange == exchange else "%s->%s" % (exchange, job.exchange)
            print("%-5s %-8s n=%d world=%d rows identical to the single-GPU matrix: %s"
                  % (used, metric, n, world, bool(flag.item())), flush=True)
        job.close()
        assert ok, "rank %d rows differ (%s, %s)" % (rank, exchange, metric)
dist.destroy_process_group()
