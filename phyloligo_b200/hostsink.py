"""The host end of the output path: finished rows of the distance matrix -> the caller's file.

The reference's --large workers assign their block row into a mapping of the output file
(``output[s] = ...`` into the np.memmap of bin/phyloligo.py:413-425, or the 'distances' dataset
of the HDF5 file, :471-478).  Here the rows come off the device by DMA into a small ring of
pinned buffers and a pool of host threads moves each finished slot into the mapping
(po_host_copy2d) while the next slot is in flight and the next panel computes.

What bounds this on a fresh output file is neither PCIe (52 GB/s) nor the host copy (90 GB/s into
resident, mapped pages) but the kernel instantiating and mapping the file's 4 KB pages: allocation
(fallocate: 14 GB/s), zeroing + mapping (populate: 8-13 GB/s) whatever the thread count on the boxes
of this pool, 5.7 GB/s for the whole chain and 12-14 GB/s when the file's pages already exist
(profiles/r02_sink_probe.log, r02_cli_sink_bench.log; tmpfs huge pages are disabled there).
``PageWarmer`` starts that work (fallocate + populate, in row order) the moment the file exists
-- while the CUDA context comes up, the FASTA is profiled and the first panels compute -- so that
as much of it as possible is off the critical path.  Page-locking the mapping itself (po_host_register) so that the DMA lands
in the file directly costs more than it saves here (6 GB/s on top of the instantiation) and is
refused by file systems with dirty tracking; the entry point stays in the C ABI
(``FileMatrix.register``) for hosts where it pays.
"""
from __future__ import annotations

import ctypes as C
import mmap
import os
import queue
import threading

import numpy as np
import torch

from . import _lib, engine
from ._lib import PhyloligoError


def host_threads(world=1):
    return max(1, (os.cpu_count() or 1) // max(1, world))


class PageWarmer(threading.Thread):
    """Instantiate the pages of the byte ranges of a mapped file, in order, in the background."""

    def __init__(self, fd, base_addr, ranges, threads=4, chunk=256 << 20, mode=None):
        super().__init__(daemon=True)
        self.fd, self.base = fd, base_addr
        self.ranges = [(int(a), int(b)) for a, b in ranges if b > a]
        self.threads = max(1, int(threads))
        self.chunk = int(chunk)
        self.mode = mode or os.environ.get("PO_SINK_WARM", "fallocate_only")
        self._halt = threading.Event()
        self.error = None

    def run(self):
        lib = _lib.load()
        try:
            for lo, hi in self.ranges:
                for a in range(lo, hi, self.chunk):
                    if self._halt.is_set():
                        return
                    b = min(hi, a + self.chunk)
                    if self.mode in ("fallocate", "fallocate_only"):
                        try:
                            os.posix_fallocate(self.fd, a, b - a)
                        except OSError:
                            self.mode = "populate"  # file system without fallocate
                    if self.mode in ("fallocate", "populate"):
                        lib.po_host_prefault(C.c_void_p(self.base + a), b - a, self.threads)
                    elif self.mode == "premap":
                        lib.po_host_premap(self.fd, C.c_void_p(self.base + a), b - a, self.threads)
        except Exception as exc:  # a warmer must never take the run down: the copies fault the pages in themselves
            self.error = exc

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join()


class FileMatrix:
    """A row-major (rows x cols) matrix stored at byte `offset` of a file, mapped for writing.

    The raw --large memmap output (offset 0) and the data region of the HDF5 'distances' dataset
    are both this.  `create` sizes the file; every rank of a multi-GPU run then attaches."""

    def __init__(self, path, rows, cols, dtype=np.float32, offset=0, create=False):
        self.path, self.rows, self.cols = path, int(rows), int(cols)
        self.dtype = np.dtype(dtype)
        self.offset = int(offset)
        self.nbytes = self.rows * self.cols * self.dtype.itemsize
        flags = os.O_RDWR | (os.O_CREAT if create else 0)
        self.fd = os.open(path, flags, 0o644)
        total = self.offset + self.nbytes
        self.fresh = False  # True: the file's pages do not exist yet
        if create and os.fstat(self.fd).st_size != total:
            os.ftruncate(self.fd, total)
            self.fresh = True
        elif os.fstat(self.fd).st_size < total:
            os.close(self.fd)
            raise PhyloligoError("%s is smaller than the %d x %d matrix it should hold" % (path, self.rows, self.cols))
        self.mm = mmap.mmap(self.fd, total, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE) if total else None
        self.array = (np.frombuffer(self.mm, dtype=self.dtype, count=self.rows * self.cols, offset=self.offset)
                      .reshape(self.rows, self.cols)) if self.nbytes else np.zeros((self.rows, self.cols), self.dtype)
        self.base = self.array.ctypes.data - self.offset if self.nbytes else 0  # address of file byte 0
        self.warmer = None
        self.registered = False
        if self.mm is not None:
            try:  # huge pages where the administrator allows them for shared memory (shmem_enabled=advise)
                self.mm.madvise(mmap.MADV_HUGEPAGE)
            except (OSError, ValueError, AttributeError):
                pass

    def row_bytes(self, r0, r1):
        es = self.dtype.itemsize
        return self.offset + int(r0) * self.cols * es, self.offset + int(r1) * self.cols * es

    def warm(self, row_ranges, threads=4, fresh=None):
        """Start instantiating the pages of the given row ranges (in that order) in the background.
        A fresh file: fallocate only (allocation runs at 14 GB/s on one thread; mapping the pages is left to
        the copies -- populating them as well contends for the same locks and measured slower, 4.7 against
        6.6 GB/s for the whole chain).  A file whose pages exist: premap -- on tmpfs populate for READING, which
        maps 16 up-to-date pages per fault with writable entries (27-43 GB/s per the probe against 2-3 GB/s per
        thread for write-populating, the "populate" mode of the first half of the round: 12.7 against 9 GB/s
        for the whole chain); other file systems: populate for writing.
        PO_SINK_WARM = none | premap | populate | fallocate | fallocate_only overrides."""
        if self.mm is None or os.environ.get("PO_SINK_WARM", "") == "none":
            return
        fresh = self.fresh if fresh is None else fresh
        page = mmap.PAGESIZE
        ranges = []
        for r0, r1 in row_ranges:
            lo, hi = self.row_bytes(r0, r1)
            ranges.append((lo // page * page, min(self.offset + self.nbytes, -(-hi // page) * page)))
        threads = int(os.environ.get("PO_SINK_WARM_THREADS", "0")) or threads
        self.warmer = PageWarmer(self.fd, self.base, ranges, threads,
                                 mode=os.environ.get("PO_SINK_WARM") or ("fallocate_only" if fresh else "premap"))
        self.warmer.start()

    def register(self):
        """Page-lock the whole mapping so that DMA lands in it (po_host_register); False when refused."""
        if self.mm is None:
            return False
        lib = _lib.load()
        if lib.po_host_register(C.c_void_p(self.array.ctypes.data), self.nbytes) == 0:
            self.registered = True
        return self.registered

    def page_spans(self, row_ranges):
        """The byte ranges [lo, hi) of the file that hold the given row ranges, widened to page boundaries,
        sorted, adjacent or overlapping ones merged (two neighbouring block rows share a page)."""
        page = mmap.PAGESIZE
        spans = []
        for r0, r1 in sorted((int(a), int(b)) for a, b in row_ranges if b > a):
            lo, hi = self.row_bytes(r0, r1)
            lo, hi = lo // page * page, min(self.offset + self.nbytes, -(-hi // page) * page)
            if spans and lo <= spans[-1][1]:
                spans[-1][1] = max(spans[-1][1], hi)
            else:
                spans.append([lo, hi])
        return spans

    def register_rows(self, row_ranges):
        """Page-lock the pages that hold the given row ranges only (a rank's own rows of a matrix every rank
        maps): adjacent or overlapping ranges are merged, every range is widened to page boundaries.
        False when the kernel refuses (nothing stays registered then)."""
        if self.mm is None:
            return False
        spans = self.page_spans(row_ranges)
        lib = _lib.load()
        done = []
        for lo, hi in spans:
            if lib.po_host_register(C.c_void_p(self.base + lo), hi - lo) != 0:
                for a in done:
                    lib.po_host_unregister(C.c_void_p(self.base + a))
                return False
            done.append(lo)
        self._registered_spans = getattr(self, "_registered_spans", []) + done
        return True

    def close(self):
        if self.warmer is not None:
            self.warmer.stop()
            self.warmer = None
        for a in getattr(self, "_registered_spans", []):
            _lib.load().po_host_unregister(C.c_void_p(self.base + a))
        self._registered_spans = []
        if self.registered:
            _lib.load().po_host_unregister(C.c_void_p(self.array.ctypes.data))
            self.registered = False
        self.array = None
        if self.mm is not None:
            try:
                self.mm.close()
            except BufferError:  # a view is still alive somewhere: the mapping goes with it
                pass
            self.mm = None
        if self.fd is not None:
            os.close(self.fd)
            self.fd = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class RowShipper:
    """Device blocks -> pinned ring -> host destination (a FileMatrix mapping or any host array).

    ``ship(src, row0, col0)`` enqueues the device-to-host DMA of the 2-D device view `src` (rows
    contiguous) on the copy stream, after everything enqueued so far on the current stream, in
    pieces of at most one ring slot; a host thread waits for each piece and copies it to
    ``dest[row0 + ..., col0 : col0 + width]`` with `copy_threads` threads.  ``finish()`` drains.
    """

    def __init__(self, dest, slot_bytes=32 << 20, slots=4, copy_threads=4):
        if dest.ndim != 2 or dest.strides[1] != dest.itemsize:
            raise PhyloligoError("RowShipper: destination rows must be contiguous")
        self.dest = dest
        self.dst_pitch = int(dest.strides[0])
        self.esize = int(dest.itemsize)
        self.tdtype = {4: torch.float32, 8: torch.float64}[self.esize]
        slot_bytes = int(os.environ.get("PO_SINK_SLOT_MB", "0")) << 20 or int(slot_bytes)
        self.slot_elems = max(int(slot_bytes) // self.esize, int(dest.shape[1]))
        self.pinned = [None] * slots
        self.free = queue.Queue()
        self.work = queue.Queue()
        self.copy_stream = torch.cuda.Stream()
        self.copy_threads = int(os.environ.get("PO_SINK_COPY_THREADS", "0")) or max(1, int(copy_threads))
        self.bytes_shipped = 0
        self.error = None
        self.lib = _lib.load()
        self.device = torch.cuda.current_device()
        # page-locking host memory is slow (~2 GB/s here): the slots are allocated by a helper thread and
        # join the ring one by one, so the first panel's DMA starts after one slot, not after all of them
        self.alloc_thread = threading.Thread(target=self._allocate, daemon=True)
        self.alloc_thread.start()
        self.thread = threading.Thread(target=self._copier, daemon=True)
        self.thread.start()

    def _allocate(self):
        try:
            torch.cuda.set_device(self.device)
            for s in range(len(self.pinned)):
                self.pinned[s] = torch.empty(self.slot_elems, dtype=self.tdtype).pin_memory()
                self.free.put(s)
        except Exception as exc:
            self.error = exc
            self.free.put(-1)

    def _copier(self):
        while True:
            item = self.work.get()
            if item is None:
                return
            slot, ev, row0, col0, rows, width = item
            try:
                ev.synchronize()
                if self.error is None:
                    dst = self.dest.ctypes.data + row0 * self.dst_pitch + col0 * self.esize
                    rc = self.lib.po_host_copy2d(C.c_void_p(dst), self.dst_pitch, C.c_void_p(self.pinned[slot].data_ptr()),
                                                 width * self.esize, width * self.esize, rows, self.copy_threads)
                    _lib.check(rc, "po_host_copy2d")
            except Exception as exc:
                self.error = exc
            finally:
                self.free.put(slot)

    def ship(self, src, row0, col0=0):
        """Returns the event that marks the end of the last DMA out of `src` (the device buffer may be
        reused once it has completed)."""
        if src.dim() != 2 or src.dtype != self.tdtype or (src.shape[1] > 1 and src.stride(1) != 1):
            raise PhyloligoError("RowShipper.ship: 2-D device view with contiguous rows of the destination dtype expected")
        rows, width = int(src.shape[0]), int(src.shape[1])
        if rows == 0 or width == 0:
            return None
        if width > self.slot_elems:
            raise PhyloligoError("RowShipper: one row (%d elements) does not fit a ring slot" % width)
        ready = torch.cuda.Event()
        ready.record()
        self.copy_stream.wait_event(ready)
        per = max(1, self.slot_elems // width)
        last = None
        for a in range(0, rows, per):
            m = min(per, rows - a)
            slot = self.free.get()
            if self.error is not None:
                self.free.put(slot)
                raise self.error
            if slot < 0:
                raise PhyloligoError("RowShipper: no pinned slot could be allocated")
            host = self.pinned[slot][: m * width].view(m, width)
            engine.copy2d(host, src[a:a + m], self.copy_stream)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
            self.work.put((slot, ev, row0 + a, col0, m, width))
            self.bytes_shipped += m * width * self.esize
            last = ev
        return last

    def finish(self):
        self.work.put(None)
        self.thread.join()
        self.alloc_thread.join()
        if self.error is not None:
            raise self.error
