#!/usr/bin/env python3
"""Sweep of the end-to-end sink split (run on the GPU box): how much of the part of the matrix left of the
diagonal the host builds by transposition (engine.matrix_to_host host_mirror=share) against how much crosses
PCIe, and with how many host threads.  C2-shaped profiles (k=4, both strands) of `n` synthetic contigs."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from phyloligo_b200 import _lib, engine, synth

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=100_000)
ap.add_argument("--length", type=int, default=20_000)
ap.add_argument("--panel", type=int, default=4096)
ap.add_argument("--shares", default="0,0.5,0.75,0.9,1.0")
ap.add_argument("--threads", default="8,12,14,16")
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
lib = _lib.load()
n = args.n
t0 = time.perf_counter()
text, b, e, bases = synth.device_fasta(n, args.length, 2, dev)
X = engine.profile_device(text, b, e, "1111", "both", want=("freq32",))["freq32"]
del text
torch.cuda.synchronize()
print("profiles %s in %.1f s" % (tuple(X.shape), time.perf_counter() - t0), flush=True)
t0 = time.perf_counter()
host = torch.empty((n, n), dtype=torch.float32).pin_memory()
print("pinned %.1f GB in %.1f s; cpus %d" % (host.numel() * 4 / 1e9, time.perf_counter() - t0, os.cpu_count()), flush=True)
matrix = torch.empty((n, n), dtype=torch.float32, device=dev)
prepared = engine.prepare(X, "JSD")


def step(share, threads):
    stats = {}
    engine.matrix_to_host(None, "JSD", host, torch.float32, args.panel, prepared=prepared, device_matrix=matrix,
                          host_mirror=share, mirror_threads=threads, stats=stats)
    return stats


def timed(share, threads):
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(args.reps):
        t = time.perf_counter()
        stats = step(share, threads)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, stats


# kernel only, for reference
torch.cuda.synchronize()
t = time.perf_counter()
engine.distance_block("JSD", *prepared, 0, n, 0, n, matrix, 0, 0, _lib.FLAG_SKIP_LOWER | _lib.FLAG_MIRROR)
torch.cuda.synchronize()
print("kernel only: %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
# the transposition alone on this host (synchronous entry point), 8 GB sample
m = min(n, 44672)
for th in (8, 16):
    t = time.perf_counter()
    lib.po_host_transpose_f32(host[m // 2:, : m // 2].data_ptr(), n, host[: m // 2, m // 2:].data_ptr(), n, m // 2, n - m // 2, th)
    dt = time.perf_counter() - t
    print("host transpose alone, %d threads: %.2f GB/s" % (th, (m // 2) * (n - m // 2) * 4 / dt / 1e9), flush=True)
step(1.0, None)  # warm-up (pool start, first touch)
torch.cuda.synchronize()
pairs = n * (n + 1) // 2
for share in [float(v) for v in args.shares.split(",")]:
    for threads in ([int(v) for v in args.threads.split(",")] if share > 0 else [0]):
        dt, stats = timed(share, threads or None)
        print("share %.2f threads %2d: %7.1f ms  %.3e pairs/s  dma %.1f GB  host-mirrored %.1f GB"
              % (share, threads, dt * 1e3, pairs / dt, stats["dma_bytes"] / 1e9, stats["host_mirrored_bytes"] / 1e9), flush=True)
# correctness of the last configuration: sampled rows against the device matrix
for r in sorted(set(int(v) for v in np.linspace(0, n - 1, 23))):
    assert torch.equal(host[r], matrix[r].cpu()), r
print("host matrix == device matrix on 23 sampled rows")
