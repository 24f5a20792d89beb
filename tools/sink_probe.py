#!/usr/bin/env python3
"""What the host end of the output path costs on this box (one GPU): instantiating fresh pages of
the output file, page-locking the mapping, DMA straight into it, and the alternatives (pinned
ring + host threads copying or pwrite-ing, host-side transposition of mirrored blocks).

    python tools/sink_probe.py [--gb 8] [--dirs /dev/shm,/tmp]
"""
import argparse
import ctypes as C
import mmap
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from phyloligo_b200 import _lib, engine

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=8.0)
ap.add_argument("--dirs", default="/dev/shm,/tmp")
args = ap.parse_args()

lib = _lib.load()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
n = int((args.gb * 1e9 / 4) ** 0.5) // 128 * 128
nbytes = n * n * 4
print("matrix %d x %d float32 = %.2f GB; cpus %d" % (n, n, nbytes / 1e9, os.cpu_count()), flush=True)
print(subprocess.run("uname -r; cat /sys/kernel/mm/transparent_hugepage/shmem_enabled; mount | grep -E ' /dev/shm | /tmp | / '",
                     shell=True, capture_output=True, text=True).stdout, flush=True)

src = torch.empty((n, n), dtype=torch.float32, device=dev)
src.uniform_()
torch.cuda.synchronize()


def rate(tag, nb, dt):
    print("%-66s %7.3f s  %7.2f GB/s" % (tag, dt, nb / dt / 1e9), flush=True)


def fresh(path):
    if os.path.exists(path):
        os.unlink(path)
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    os.ftruncate(fd, nbytes)
    mm = mmap.mmap(fd, nbytes, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
    arr = np.frombuffer(mm, dtype=np.float32).reshape(n, n)
    return fd, mm, arr


def drop(fd, mm, arr, path):
    del arr
    try:
        mm.close()
    except BufferError:
        pass
    os.close(fd)
    os.unlink(path)


def d2h_into(ptr, tag, reps=1):
    """whole matrix by 1 GB row panels, cudaMemcpy2DAsync into host address ptr"""
    st = torch.cuda.current_stream().cuda_stream
    rows = max(128, (1 << 30) // (n * 4) // 128 * 128)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for r0 in range(0, n, rows):
            m = min(rows, n - r0)
            rc = lib.po_copy2d_async(C.c_void_p(ptr + r0 * n * 4), n * 4, C.c_void_p(src.data_ptr() + r0 * n * 4), n * 4, n * 4, m, C.c_void_p(st))
            _lib.check(rc, "copy2d")
    torch.cuda.synchronize()
    rate(tag, reps * nbytes, time.perf_counter() - t0)


# ---- baseline: pinned allocation and D2H into it ----
t0 = time.perf_counter()
ring = [torch.empty((256 << 20,), dtype=torch.uint8).pin_memory() for _ in range(2)]
rate("pin_memory 2 x 256 MB", 2 * (256 << 20), time.perf_counter() - t0)
t0 = time.perf_counter()
big_pinned = torch.empty((n, n), dtype=torch.float32).pin_memory()
rate("pin_memory whole matrix (cudaHostAlloc)", nbytes, time.perf_counter() - t0)
d2h_into(big_pinned.data_ptr(), "D2H into cudaHostAlloc'ed matrix", reps=2)

threads_list = [1, 4, 8, 16, 32]
threads_list = [t for t in threads_list if t <= 2 * (os.cpu_count() or 1)]

# ---- host transposition ----
tr_dst = np.empty((n, n), dtype=np.float32)
tr_dst.fill(0)
for th in threads_list:
    t0 = time.perf_counter()
    lib.po_host_transpose_f32(tr_dst.ctypes.data, n, big_pinned.data_ptr(), n, n, n, th)
    rate("host transpose, %d threads" % th, nbytes, time.perf_counter() - t0)
for th in threads_list:
    t0 = time.perf_counter()
    lib.po_host_copy2d(tr_dst.ctypes.data, n * 4, big_pinned.data_ptr(), n * 4, n * 4, n, th)
    rate("host copy (warm pages), %d threads" % th, nbytes, time.perf_counter() - t0)
del tr_dst

for d in args.dirs.split(","):
    if not os.path.isdir(d):
        continue
    path = os.path.join(d, "po_sink_probe.bin")
    print("---- %s ----" % d, flush=True)
    # prefault scaling on fresh files
    for th in threads_list:
        fd, mm, arr = fresh(path)
        t0 = time.perf_counter()
        rc = lib.po_host_prefault(arr.ctypes.data, nbytes, th)
        rate("prefault fresh file, %d threads (rc %d)" % (th, rc), nbytes, time.perf_counter() - t0)
        if th == threads_list[-1]:
            # register the populated mapping, DMA into it
            t0 = time.perf_counter()
            rc = lib.po_host_register(arr.ctypes.data, nbytes)
            rate("cudaHostRegister populated mapping (rc %d)" % rc, nbytes, time.perf_counter() - t0)
            if rc == 0:
                d2h_into(arr.ctypes.data, "D2H straight into the registered file mapping", reps=2)
                assert np.array_equal(arr[n // 2, :64], src[n // 2, :64].cpu().numpy())
                t0 = time.perf_counter()
                lib.po_host_unregister(arr.ctypes.data)
                rate("cudaHostUnregister", nbytes, time.perf_counter() - t0)
            else:
                print("   register failed:", lib.po_last_error().decode(), flush=True)
        drop(fd, mm, arr, path)
    # register a fresh (unpopulated) mapping
    fd, mm, arr = fresh(path)
    t0 = time.perf_counter()
    rc = lib.po_host_register(arr.ctypes.data, nbytes)
    rate("cudaHostRegister FRESH mapping (rc %d)" % rc, nbytes, time.perf_counter() - t0)
    if rc == 0:
        d2h_into(arr.ctypes.data, "D2H into it", reps=1)
        lib.po_host_unregister(arr.ctypes.data)
    drop(fd, mm, arr, path)
    # chunked: prefault + register 512 MB chunks
    fd, mm, arr = fresh(path)
    chunk = 512 << 20
    t0 = time.perf_counter()
    ok = True
    for off in range(0, nbytes, chunk):
        sz = min(chunk, nbytes - off)
        lib.po_host_prefault(arr.ctypes.data + off, sz, 16)
        if lib.po_host_register(arr.ctypes.data + off, sz) != 0:
            ok = False
            break
    rate("prefault(16 thr) + register in 512 MB chunks (ok %s)" % ok, nbytes, time.perf_counter() - t0)
    if ok:
        for off in range(0, nbytes, chunk):
            lib.po_host_unregister(arr.ctypes.data + off)
    drop(fd, mm, arr, path)
    # pinned matrix -> mapping with host threads (fresh pages), and pwrite
    for th in [t for t in threads_list if t >= 4]:
        fd, mm, arr = fresh(path)
        t0 = time.perf_counter()
        lib.po_host_copy2d(arr.ctypes.data, n * 4, big_pinned.data_ptr(), n * 4, n * 4, n, th)
        rate("host copy pinned -> FRESH mapping, %d threads" % th, nbytes, time.perf_counter() - t0)
        t0 = time.perf_counter()
        lib.po_host_copy2d(arr.ctypes.data, n * 4, big_pinned.data_ptr(), n * 4, n * 4, n, th)
        rate("host copy pinned -> populated mapping, %d threads" % th, nbytes, time.perf_counter() - t0)
        drop(fd, mm, arr, path)
        fd, mm, arr = fresh(path)
        t0 = time.perf_counter()
        rc = lib.po_host_pwrite2d(fd, 0, n * 4, big_pinned.data_ptr(), n * 4, n * 4, n, th)
        rate("pwrite pinned -> FRESH file, %d threads (rc %d)" % (th, rc), nbytes, time.perf_counter() - t0)
        t0 = time.perf_counter()
        rc = lib.po_host_pwrite2d(fd, 0, n * 4, big_pinned.data_ptr(), n * 4, n * 4, n, th)
        rate("pwrite pinned -> existing file, %d threads" % th, nbytes, time.perf_counter() - t0)
        drop(fd, mm, arr, path)
    # fallocate then prefault
    fd, mm, arr = fresh(path)
    t0 = time.perf_counter()
    try:
        os.posix_fallocate(fd, 0, nbytes)
        rate("posix_fallocate", nbytes, time.perf_counter() - t0)
        t0 = time.perf_counter()
        lib.po_host_prefault(arr.ctypes.data, nbytes, 16)
        rate("prefault after fallocate, 16 threads", nbytes, time.perf_counter() - t0)
    except OSError as exc:
        print("fallocate failed", exc)
    drop(fd, mm, arr, path)

# anonymous memory for comparison (what cudaHostAlloc pays)
t0 = time.perf_counter()
mm = mmap.mmap(-1, nbytes, mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
arr = np.frombuffer(mm, dtype=np.uint8)
lib.po_host_prefault(arr.ctypes.data, nbytes, 16)
rate("anonymous mmap + prefault 16 threads", nbytes, time.perf_counter() - t0)
t0 = time.perf_counter()
rc = lib.po_host_register(arr.ctypes.data, nbytes)
rate("cudaHostRegister anonymous populated (rc %d)" % rc, nbytes, time.perf_counter() - t0)
if rc == 0:
    d2h_into(arr.ctypes.data, "D2H into registered anonymous memory", reps=2)
    lib.po_host_unregister(arr.ctypes.data)
