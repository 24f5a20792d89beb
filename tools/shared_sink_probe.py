#!/usr/bin/env python3
"""Where does the upper-triangle end-to-end sink spend its time (run on the GPU box, one GPU)?  The same panel
pipeline as engine.matrix_to_host / multigpu.MirroredHostSink without the distance kernels, into (A) a
cudaHostAlloc'ed matrix and (B) a page-locked matrix in a /dev/shm file (what the multi-GPU path uses):
DMA of the panels' right parts alone, the host mirroring alone, and both pipelined."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from phyloligo_b200 import _lib, engine, hostsink

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=60_000)
ap.add_argument("--panel", type=int, default=4096)
ap.add_argument("--sub", type=int, default=1024)
ap.add_argument("--threads", default="4,8,14")
ap.add_argument("--only-pinned", action="store_true")
ap.add_argument("--dma-rows", type=int, default=0)
args = ap.parse_args()
n, panel, sub = args.n, args.panel, args.sub
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()
matrix = torch.empty((n, n), dtype=torch.float32, device=dev)
matrix.copy_(torch.rand((1, n), device=dev).expand(n, n))
matrix += torch.arange(n, device=dev, dtype=torch.float32)[:, None]
torch.cuda.synchronize()
print("PO_HOST_MIRROR_JOB_ROWS=%s dma-rows=%d" % (os.environ.get("PO_HOST_MIRROR_JOB_ROWS", "(default)"), args.dma_rows))
print("n = %d: %.1f GB; cpus %d; THP anon: %s; shmem: %s" % (
    n, n * n * 4 / 1e9, os.cpu_count(),
    open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(),
    open("/sys/kernel/mm/transparent_hugepage/shmem_enabled").read().strip()), flush=True)


def panels():
    for r0 in range(0, n, panel):
        yield r0, min(n, r0 + panel)


def run(host, pool, do_dma, do_mirror, sub_rows):
    """one pass; returns seconds.  sub_rows = rows per mirror submit (one stream callback each);
    --dma-rows = rows per strided DMA (0: the same)"""
    cs = torch.cuda.Stream()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for r0, r1 in panels():
        for s0 in range(r0, r1, sub_rows):
            s1 = min(r1, s0 + sub_rows)
            if do_dma:
                step = args.dma_rows or (s1 - s0)
                for d0 in range(s0, s1, step):
                    d1 = min(s1, d0 + step)
                    engine.copy2d(host[d0:d1, r0:], matrix[d0:d1, r0:], cs)
            if do_mirror and r1 < n:
                pool.submit(host[r1:, s0:s1], host[s0:s1, r1:], cs, after_stream=do_dma)
    cs.synchronize()
    if do_mirror:
        pool.wait()
    return time.perf_counter() - t


def run_independent(host, pool, sub_rows):
    """DMA of every panel and the mirroring of every panel at the same time, with no ordering between them (the
    mirror reads what the previous pass left): pure interference, no callbacks.  Returns (dma_s, mirror_s, both_s)."""
    cs = torch.cuda.Stream()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record(cs)
    for r0, r1 in panels():
        for s0 in range(r0, r1, sub_rows):
            s1 = min(r1, s0 + sub_rows)
            engine.copy2d(host[s0:s1, r0:], matrix[s0:s1, r0:], cs)
    e1.record(cs)
    for r0, r1 in panels():
        for s0 in range(r0, r1, sub_rows):
            s1 = min(r1, s0 + sub_rows)
            if r1 < n:
                pool.submit(host[r1:, s0:s1], host[s0:s1, r1:], None, after_stream=False)
    pool.wait()
    tm = time.perf_counter() - t
    cs.synchronize()
    tb = time.perf_counter() - t
    return e0.elapsed_time(e1) * 1e-3, tm, tb


tri = sum((r1 - r0) * (n - r0) for r0, r1 in panels()) * 4
low = sum((r1 - r0) * (n - r1) for r0, r1 in panels()) * 4
for name in ("cudaHostAlloc",) if args.only_pinned else ("cudaHostAlloc", "/dev/shm file, page-locked"):
    t0 = time.perf_counter()
    fm = None
    if name == "cudaHostAlloc":
        host = torch.empty((n, n), dtype=torch.float32).pin_memory()
    else:
        path = "/dev/shm/po_sink_probe.mat"
        fm = hostsink.FileMatrix(path, n, n, np.float32, create=True)
        os.posix_fallocate(fm.fd, 0, fm.nbytes)
        assert fm.register_rows([(0, n)])
        host = torch.from_numpy(fm.array)
    print("---- %s (set up in %.1f s)" % (name, time.perf_counter() - t0), flush=True)
    for threads in [int(v) for v in args.threads.split(",")]:
        pool = engine.HostMirror(threads)
        for sub_rows in (panel, sub):
            run(host, pool, True, True, sub_rows)  # warm-up: page tables, pool
            d = min(run(host, pool, True, False, sub_rows) for _ in range(2))
            m = min(run(host, pool, False, True, sub_rows) for _ in range(2))
            b = min(run(host, pool, True, True, sub_rows) for _ in range(2))
            print("threads %2d sub-panels of %4d rows: DMA alone %6.1f ms (%.1f GB/s)   mirror alone %6.1f ms (%.1f GB/s)   "
                  "pipelined %6.1f ms" % (threads, sub_rows, d * 1e3, tri / d / 1e9, m * 1e3, low / m / 1e9, b * 1e3), flush=True)
            di, mi, bi = run_independent(host, pool, sub_rows)
            print("           unordered, at the same time:  DMA %6.1f ms (%.1f GB/s)   mirror %6.1f ms (%.1f GB/s)   both done after %6.1f ms"
                  % (di * 1e3, tri / di / 1e9, mi * 1e3, low / mi / 1e9, bi * 1e3), flush=True)
        pool.close()
    k = (n // 2) // panel * panel  # left of the panel that holds row n // 2: everything there was mirrored
    ok = torch.equal(host[n // 2, :k], matrix[:k, n // 2].cpu()) and torch.equal(host[n // 2, n // 2:], matrix[n // 2, n // 2:].cpu())
    print("mirrored rows correct: %s" % ok, flush=True)
    del host
    if fm is not None:
        fm.close()
        os.unlink(path)
