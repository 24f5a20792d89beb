"""Host-side driver of the CUDA hot path: device buffers (torch), launches (C ABI).

torch is used for allocation, streams and host<->device copies only; every
computation happens in libphyloligo_b200.so.  Nothing here falls back to the
CPU: without a CUDA device the functions raise PhyloligoError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import PhyloligoError, METRICS, STRANDS, PO_F32, PO_F64, FLAG_MIRROR, FLAG_SKIP_LOWER, TILE


def require_cuda():
    if not torch.cuda.is_available():
        raise PhyloligoError("no CUDA device visible: phyloligo_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    _lib.load()
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


# ----------------------------------------------------------------------------
# FASTA
# ----------------------------------------------------------------------------
def as_u8(text) -> np.ndarray:
    if isinstance(text, np.ndarray):
        if text.dtype != np.uint8:
            raise TypeError("text array must be uint8")
        return np.ascontiguousarray(text)
    if isinstance(text, torch.Tensor):
        return text.numpy()
    if isinstance(text, str):
        text = text.encode("latin-1")
    return np.frombuffer(text, dtype=np.uint8)


def fasta_index(text, threads=0):
    """Record index of FASTA text held in host memory: (begin, end) int64 arrays
    of the byte range of every record's sequence lines (Biopython record rules,
    reference call sites bin/phyloligo.py:87,114,154,869,914,959)."""
    lib = _lib.load()
    buf = as_u8(text)
    n_bytes = buf.shape[0]
    cap = max(1024, n_bytes // 2000)
    while True:
        begin = np.empty(cap, dtype=np.int64)
        end = np.empty(cap, dtype=np.int64)
        n = lib.po_fasta_index_host(buf.ctypes.data, n_bytes, begin.ctypes.data, end.ctypes.data, cap, threads)
        _lib.check(n, "po_fasta_index_host")
        if n <= cap:
            return begin[:n].copy(), end[:n].copy()
        cap = int(n)


def fasta_header_starts(text, begin, end):
    """Position of the '>' of every record of `text` (a bytes-like object), given the sequence ranges of
    fasta_index: a record's header starts where the previous record's range ends; the first one starts
    after the last line break before its own line."""
    n = len(begin)
    starts = np.zeros(n, dtype=np.int64)
    if n == 0:
        return starts
    starts[1:] = np.asarray(end[:-1], dtype=np.int64)
    raw = bytes(memoryview(text)[: int(begin[0])])
    j = len(raw)
    while j > 0 and raw[j - 1] in (10, 13):  # the line break(s) that end the first header line
        j -= 1
    starts[0] = max(raw.rfind(b"\n", 0, j), raw.rfind(b"\r", 0, j)) + 1
    return starts


def sequences_to_text(seqs):
    """Pack bare sequences (str/bytes) into one buffer, newline separated, with
    their (begin, end) ranges -- the batch form of the per-sequence worker API."""
    parts, begin, end = [], [], []
    pos = 0
    for s in seqs:
        b = s.encode("latin-1") if isinstance(s, str) else bytes(s)
        begin.append(pos)
        end.append(pos + len(b))
        parts.append(b)
        parts.append(b"\n")
        pos += len(b) + 1
    text = np.frombuffer(b"".join(parts), dtype=np.uint8) if parts else np.zeros(0, dtype=np.uint8)
    return text, np.asarray(begin, dtype=np.int64), np.asarray(end, dtype=np.int64)


# ----------------------------------------------------------------------------
# Profiling
# ----------------------------------------------------------------------------
def text_to_device(text, device=None, non_blocking=False):
    """Copy FASTA bytes to the device, padded so 16-byte loads never leave the buffer."""
    device = device or require_cuda()
    if isinstance(text, torch.Tensor):
        src = text
    else:
        arr = as_u8(text)
        if not arr.flags.writeable:  # read-only views (bytes, read-only mappings) are only read from
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                src = torch.from_numpy(arr)
        else:
            src = torch.from_numpy(arr)
    n = src.shape[0]
    dst = torch.empty(n + 64, dtype=torch.uint8, device=device)
    dst[:n].copy_(src, non_blocking=non_blocking)
    dst[n:].fill_(10)
    return dst


def file_to_device(path, byte_lo=0, byte_hi=None, device=None, slot_bytes=32 << 20, threads=4):
    """Bytes [byte_lo, byte_hi) of a file straight to the device, padded like text_to_device: parallel
    pread (po_host_pread) into a two-slot pinned ring, each slot leaving by asynchronous H2D while the
    other is being read.  Nothing the size of the file is page-locked or copied on the host (a
    pageable np.fromfile of a 2 GB assembly plus its staged copy costs more than profiling it)."""
    import os
    device = device or require_cuda()
    lib = _lib.load()
    size = os.path.getsize(path)
    byte_hi = size if byte_hi is None else min(int(byte_hi), size)
    byte_lo = int(byte_lo)
    n = max(0, byte_hi - byte_lo)
    dst = torch.empty(n + 64, dtype=torch.uint8, device=device)
    dst[n:].fill_(10)
    if n == 0:
        return dst
    slot_bytes = int(min(slot_bytes, max(1 << 20, n)))
    ring = [torch.empty(slot_bytes, dtype=torch.uint8).pin_memory() for _ in range(2)]
    events = [None, None]
    fd = os.open(path, os.O_RDONLY)
    try:
        for k, off in enumerate(range(0, n, slot_bytes)):
            m = min(slot_bytes, n - off)
            slot = k & 1
            if events[slot] is not None:
                events[slot].synchronize()
            _lib.check(lib.po_host_pread(fd, byte_lo + off, C.c_void_p(ring[slot].data_ptr()), m, threads), "po_host_pread")
            dst[off:off + m].copy_(ring[slot][:m], non_blocking=True)
            events[slot] = torch.cuda.Event()
            events[slot].record()
    finally:
        os.close(fd)
    for ev in events:
        if ev is not None:
            ev.synchronize()  # the ring is released on return
    return dst


def profile_device(d_text, d_begin, d_end, pattern, strand, want=("counts", "totals", "freq64")):
    """Launch po_profile_batch on device tensors; returns a dict of device tensors."""
    device = require_cuda()
    lib = _lib.load()
    pattern = str(pattern)
    if strand not in STRANDS:
        raise PhyloligoError("Error, strand parameter should be chosen from {'both', 'minus', 'plus'}")
    _, _, dim = _lib.pattern_info(pattern)
    n = int(d_begin.shape[0])
    out = {}
    counts = totals = f64 = f32 = None
    if "counts" in want or dim * 4 > 160 * 1024:
        counts = torch.empty((n, dim), dtype=torch.int32, device=device)
    if "totals" in want:
        totals = torch.empty((n,), dtype=torch.int64, device=device)
    if "freq64" in want:
        f64 = torch.empty((n, dim), dtype=torch.float64, device=device)
    if "freq32" in want:
        f32 = torch.empty((n, dim), dtype=torch.float32, device=device)
    rc = lib.po_profile_batch(_ptr(d_text), _ptr(d_begin), _ptr(d_end), n, pattern.encode(), STRANDS[strand],
                              _ptr(counts), _ptr(totals), _ptr(f64), _ptr(f32), _stream())
    _lib.check(rc, "po_profile_batch")
    if counts is not None and "counts" in want:
        out["counts"] = counts
    if totals is not None:
        out["totals"] = totals
    if f64 is not None:
        out["freq64"] = f64
    if f32 is not None:
        out["freq32"] = f32
    return out


def profile_text(text, pattern, strand="both", want=("freq64",), begin=None, end=None):
    """Profile every record of FASTA text in host memory; returns device tensors."""
    device = require_cuda()
    if begin is None:
        begin, end = fasta_index(text)
    d_text = text_to_device(text, device)
    d_begin = torch.from_numpy(np.ascontiguousarray(begin)).to(device)
    d_end = torch.from_numpy(np.ascontiguousarray(end)).to(device)
    return profile_device(d_text, d_begin, d_end, pattern, strand, want)


def profile_sequences(seqs, pattern, strand="both", want=("freq64",)):
    text, begin, end = sequences_to_text(seqs)
    return profile_text(text, pattern, strand, want, begin, end)


# ----------------------------------------------------------------------------
# Distances
# ----------------------------------------------------------------------------
def prepare(X, metric):
    """po_prepare_profiles: X is a device tensor (n, dim) float32/float64.
    Returns (P int32 tensor (n, words) holding po_prepared_bytes of operands,
    aux float64 tensor (n,), dim)."""
    device = require_cuda()
    lib = _lib.load()
    if metric not in METRICS:
        raise PhyloligoError("Error, unknown method {}".format(metric))
    if X.dim() != 2:
        raise PhyloligoError("profiles must be a 2-D matrix")
    if X.dtype not in (torch.float32, torch.float64):
        X = X.to(torch.float64)
    X = X.contiguous()
    n, dim = int(X.shape[0]), int(X.shape[1])
    total = lib.po_prepared_bytes(METRICS[metric], n, dim)
    _lib.check(total, "po_prepared_bytes")
    free = torch.cuda.mem_get_info()[0] if total >= (4 << 30) else total  # only worth a driver call for large operands
    if total > free:
        hint = (" (KT keeps two bit planes over the dim(dim-1)/2 element pairs of every profile: %.1f MB per "
                "profile at dim = %d)" % (total / max(1, n) / 1e6, dim)) if metric == "KT" else ""
        raise PhyloligoError("%s: the prepared operands of %d profiles of dimension %d need %.1f GB, %.1f GB of "
                             "device memory are free%s" % (metric, n, dim, total / 1e9, free / 1e9, hint))
    # a 2-D view with one row per profile keeps n visible to the callers; the buffer
    # itself is the opaque operand layout of the library (JSD pads it to 64-profile groups)
    words = total // 4
    per_row = -(-words // max(1, n))
    P = torch.empty((max(1, n) * per_row,), dtype=torch.int32, device=device)[: n * per_row].view(n, per_row)
    aux = torch.zeros((n,), dtype=torch.float64, device=device)
    rc = lib.po_prepare_profiles(METRICS[metric], _ptr(X), PO_F32 if X.dtype == torch.float32 else PO_F64,
                                 n, dim, dim, _ptr(P), _ptr(aux), _stream())
    _lib.check(rc, "po_prepare_profiles")
    return P, aux, dim


def rank_transform(X):
    """Average ranks (1..dim) of every row of the device tensor X (n, dim): float64 (n, dim)."""
    device = require_cuda()
    lib = _lib.load()
    if X.dim() != 2 or X.dtype not in (torch.float32, torch.float64):
        raise PhyloligoError("rank_transform expects a 2-D float32 / float64 device tensor")
    X = X.contiguous()
    n, dim = int(X.shape[0]), int(X.shape[1])
    R = torch.empty((n, dim), dtype=torch.float64, device=device)
    rc = lib.po_rank_transform(_ptr(X), PO_F32 if X.dtype == torch.float32 else PO_F64, n, dim, dim, _ptr(R), dim, _stream())
    _lib.check(rc, "po_rank_transform")
    return R


def distance_block(metric, P, aux, dim, row0, row1, col0, col1, out, out_row0, out_col0, flags=0,
                   mirror=None, mirror_row0=0, mirror_col0=0, mirror_ld=None):
    """po_distance_block into the device tensor `out` (2-D, float32 or float64).  With
    `mirror` the mirrored tiles go to mirror[c - mirror_row0, r - mirror_col0]: a tensor of the
    output's dtype, or a raw device address (int, e.g. a peer GPU's rows opened with
    PeerRows) together with its row pitch `mirror_ld` in elements."""
    lib = _lib.load()
    n = int(P.shape[0])
    dt = PO_F32 if out.dtype == torch.float32 else PO_F64
    if mirror is None:
        rc = lib.po_distance_block(METRICS[metric], _ptr(P), _ptr(aux), n, dim, row0, row1, col0, col1,
                                   _ptr(out), int(out.stride(0)), out_row0, out_col0, dt, flags, _stream())
    else:
        if isinstance(mirror, torch.Tensor):
            if mirror.dtype != out.dtype:
                raise PhyloligoError("mirror buffer must have the dtype of the output")
            mptr, mld = _ptr(mirror), int(mirror.stride(0))
        else:
            mptr, mld = C.c_void_p(int(mirror)), int(mirror_ld)
        rc = lib.po_distance_block_ex(METRICS[metric], _ptr(P), _ptr(aux), n, dim, row0, row1, col0, col1,
                                      _ptr(out), int(out.stride(0)), out_row0, out_col0,
                                      mptr, mld, mirror_row0, mirror_col0, dt, flags, _stream())
    _lib.check(rc, "po_distance_block")


class PeerRows:
    """The block-row buffers of all ranks of one node, mapped into this process (CUDA IPC).

    Every rank passes the tensor that holds its rows; after construction `address(rank)` is the
    device address of that rank's buffer as seen from here, usable as the `mirror` of
    distance_block.  Collective over the default process group (all_gather_object)."""

    def __init__(self, local_rows, rank, world):
        import torch.distributed as dist
        lib = _lib.load()
        handle = (C.c_ubyte * 64)()
        off = C.c_int64()
        mine = None
        if lib.po_ipc_export(_ptr(local_rows), handle, C.byref(off)) == 0:
            mine = (bytes(handle), int(off.value), torch.cuda.current_device())
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)  # every rank reaches this, whether its export worked or not
        if any(e is None for e in everyone):
            raise PhyloligoError("peer memory: po_ipc_export failed on rank(s) %s"
                                 % [r for r, e in enumerate(everyone) if e is None])
        self.local = local_rows
        self.rank = rank
        self._bases = {}
        self._addr = {}
        try:
            for r, (h, offset, dev) in enumerate(everyone):
                if r == rank:
                    self._addr[r] = local_rows.data_ptr()
                    continue
                base = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(lib.po_ipc_open(buf, C.byref(base)), "po_ipc_open")
                self._bases[r] = base
                self._addr[r] = int(base.value) + offset
        except PhyloligoError:
            self.close()
            raise

    def address(self, rank):
        return self._addr[rank]

    def close(self):
        lib = _lib.load()
        for base in self._bases.values():
            lib.po_ipc_close(base)
        self._bases = {}


def distance_matrix_device(X, metric, out_dtype=torch.float64, symmetric=True):
    """Full n x n matrix resident on the device."""
    device = require_cuda()
    P, aux, dim = prepare(X, metric)
    n = int(P.shape[0])
    out = torch.empty((n, n), dtype=out_dtype, device=device)
    flags = (FLAG_SKIP_LOWER | FLAG_MIRROR) if symmetric else 0
    # row panels keep every grid dimension inside the launch limits
    step = 65535 * 32
    for r0 in range(0, n, step):
        distance_block(metric, P, aux, dim, r0, min(n, r0 + step), 0, n, out, 0, 0, flags)
    if dim < 2 and metric == "KT":
        out.fill_(1.0)  # kendall() with no element pair: distance 0 -> KT = 1
    return out


def pair_distance(a, b, metric):
    """One pair through the tile kernel (the phylodist.<metric>(a, b) worker)."""
    device = require_cuda()
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    if a.shape != b.shape:
        raise PhyloligoError("profiles must have the same length")
    X = torch.from_numpy(np.stack([a, b])).to(device)
    out = distance_matrix_device(X, metric, torch.float64, symmetric=False)
    return float(out[0, 1].item())


def auto_panel_rows(n, element_size=4, target_bytes=256 << 20):
    """Rows per panel so that one pinned panel buffer is about 256 MB (page-locking host memory costs
    ~0.5 s per GB; the copies are long enough to run at link speed well below that)."""
    rows = target_bytes // max(1, int(n) * int(element_size))
    return int(max(TILE, min(4096, (rows // TILE) * TILE)))


class PanelStreamer:
    """Compute the matrix in row panels and stream each finished panel to the host.

    Symmetric mode keeps the whole n x n float32 matrix on the device (needs
    4 n^2 bytes of HBM): panel p computes only the tiles right of the diagonal and
    mirrors them, so by the time panel p is done its rows are complete and can
    leave over PCIe while panel p+1 computes.  Otherwise each panel computes its
    full rows into one of two panel buffers.
    """

    def __init__(self, X, metric, out_dtype=torch.float32, panel_rows=None, symmetric=None, rows=None):
        self.device = require_cuda()
        self.metric = metric
        self.P, self.aux, self.dim = prepare(X, metric)
        self.n = int(self.P.shape[0])
        self.out_dtype = out_dtype
        esize = 4 if out_dtype == torch.float32 else 8
        if panel_rows is None:
            panel_rows = auto_panel_rows(self.n, esize)
        self.panel_rows = max(TILE, (int(panel_rows) // TILE) * TILE)
        if symmetric is None:
            free, _ = torch.cuda.mem_get_info()
            symmetric = rows is None and self.n * self.n * esize < 0.6 * free
        self.symmetric = bool(symmetric)
        self.rows = rows  # optional (start, stop) row range owned by this rank
        self.compute_stream = torch.cuda.current_stream()
        self.copy_stream = torch.cuda.Stream()
        if self.symmetric:
            self.full = torch.empty((self.n, self.n), dtype=out_dtype, device=self.device)
            self.bufs = None
        else:
            self.full = None
            self.bufs = [torch.empty((self.panel_rows, self.n), dtype=out_dtype, device=self.device) for _ in range(2)]
        self.pinned = [torch.empty((self.panel_rows, self.n), dtype=out_dtype).pin_memory() for _ in range(2)]
        self.pairs_computed = 0

    def panels(self):
        lo, hi = (0, self.n) if self.rows is None else self.rows
        r = lo
        while r < hi:
            yield r, min(hi, r + self.panel_rows)
            r += self.panel_rows

    def run(self, sink):
        """sink(row0, row1, host_array) is called for every finished panel, in order.
        host_array is a view of a pinned buffer valid only during the call."""
        done_events = [None, None]   # copy finished -> pinned buffer may be consumed
        free_events = [None, None]   # device panel buffer free again
        pending = []                 # (slot, r0, r1)
        k = 0
        for r0, r1 in self.panels():
            slot = k & 1
            m = r1 - r0
            if self.symmetric:
                distance_block(self.metric, self.P, self.aux, self.dim, r0, r1, 0, self.n, self.full, 0, 0,
                               FLAG_SKIP_LOWER | FLAG_MIRROR)
                src = self.full[r0:r1]
                for t0 in range(r0, r1, TILE):  # tiles left of the diagonal tile are skipped
                    self.pairs_computed += (min(r1, t0 + TILE) - t0) * (self.n - t0)
            else:
                if free_events[slot] is not None:
                    self.compute_stream.wait_event(free_events[slot])
                distance_block(self.metric, self.P, self.aux, self.dim, r0, r1, 0, self.n, self.bufs[slot], r0, 0, 0)
                src = self.bufs[slot][:m]
                self.pairs_computed += m * self.n
            ready = torch.cuda.Event()
            ready.record(self.compute_stream)
            # the pinned slot must have been consumed by the sink before it is overwritten
            while pending and pending[0][0] == slot:
                s, p0, p1 = pending.pop(0)
                done_events[s].synchronize()
                sink(p0, p1, self.pinned[s][: p1 - p0].numpy())
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                self.pinned[slot][:m].copy_(src, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                done_events[slot] = ev
                free_events[slot] = ev
            pending.append((slot, r0, r1))
            k += 1
        for s, p0, p1 in pending:
            done_events[s].synchronize()
            sink(p0, p1, self.pinned[s][: p1 - p0].numpy())
        return self.pairs_computed


def copy2d(dst, src, stream=None):
    """dst[...] = src for two 2-D views of equal shape and dtype whose rows are contiguous
    (stride(1) == 1): one strided DMA (po_copy2d_async), no temporaries.  Either side may be a
    device tensor or a pinned host tensor."""
    lib = _lib.load()
    if dst.shape != src.shape or dst.dtype != src.dtype or dst.dim() != 2:
        raise PhyloligoError("copy2d: views must be 2-D with equal shape and dtype")
    rows, cols = int(dst.shape[0]), int(dst.shape[1])
    if rows == 0 or cols == 0:
        return
    if (cols > 1 and (dst.stride(1) != 1 or src.stride(1) != 1)):
        raise PhyloligoError("copy2d: rows must be contiguous")
    es = dst.element_size()
    st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
    rc = lib.po_copy2d_async(_ptr(dst), int(dst.stride(0)) * es, _ptr(src), int(src.stride(0)) * es, cols * es, rows, st)
    _lib.check(rc, "po_copy2d_async")


class HostMirror:
    """A pool of host threads that builds mirrored blocks of a symmetric float32 matrix in host memory, in
    stream order (po_host_mirror_*): ``submit(dst, src, stream)`` queues ``dst[...] = src.T`` and the pool
    starts on it once everything enqueued on `stream` before the call (the DMA that brings `src`) has
    completed; ``wait()`` blocks until every submitted block is written."""

    def __init__(self, threads=0):
        self.lib = _lib.load()
        self.threads = int(threads) if int(threads) > 0 else (os.cpu_count() or 1)
        self.handle = self.lib.po_host_mirror_open(self.threads)
        if not self.handle:
            raise PhyloligoError("po_host_mirror_open: " + self.lib.po_last_error().decode(errors="replace"))

    def submit(self, dst, src, stream=None, after_stream=True):
        if (dst.dtype != torch.float32 or src.dtype != torch.float32 or dst.dim() != 2 or src.dim() != 2
                or dst.shape[0] != src.shape[1] or dst.shape[1] != src.shape[0] or dst.is_cuda or src.is_cuda):
            raise PhyloligoError("HostMirror.submit: dst and src must be 2-D float32 host views, dst.shape == src.T.shape")
        rows, cols = int(src.shape[0]), int(src.shape[1])
        if rows == 0 or cols == 0:
            return
        if (cols > 1 and src.stride(1) != 1) or (rows > 1 and dst.stride(1) != 1):
            raise PhyloligoError("HostMirror.submit: rows must be contiguous")
        st = C.c_void_p((stream or torch.cuda.current_stream()).cuda_stream) if after_stream else None
        rc = self.lib.po_host_mirror_submit(self.handle, st, 1 if after_stream else 0, _ptr(dst), int(dst.stride(0)),
                                            _ptr(src), int(src.stride(0)), rows, cols)
        _lib.check(rc, "po_host_mirror_submit")

    def wait(self):
        _lib.check(self.lib.po_host_mirror_wait(self.handle), "po_host_mirror_wait")

    def close(self):
        if self.handle:
            self.lib.po_host_mirror_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


_MIRRORS = {}


def host_mirror_pool(threads=None):
    """The process-wide HostMirror with `threads` threads (default: PO_HOST_MIRROR_THREADS, else all cores but
    two -- the CUDA callback thread and the thread that drives the device want one each)."""
    if threads is None:
        threads = int(os.environ.get("PO_HOST_MIRROR_THREADS", "0")) or max(1, len(os.sched_getaffinity(0)) - 2)
    pool = _MIRRORS.get(threads)
    if pool is None or not pool.handle:
        pool = _MIRRORS[threads] = HostMirror(threads)
    return pool


def default_host_mirror_share():
    """Share of every panel's mirrored column block that the host builds (the rest crosses PCIe):
    PO_HOST_MIRROR in [0, 1]."""
    return min(1.0, max(0.0, float(os.environ.get("PO_HOST_MIRROR", "%r" % HOST_MIRROR_DEFAULT))))


# Measured on the B200 hosts of this pool (16 cores, PCIe 52 GB/s, tools/e2e_mirror_sweep.py): see DESIGN.md 6.1
HOST_MIRROR_DEFAULT = 1.0
# rows per strided DMA of a panel's right part (tools/shared_sink_probe.py, profiles/r02e_sink_job_rows_probe.log)
DMA_ROWS = 256


def matrix_to_host(X, metric, host, out_dtype=torch.float32, panel_rows=4096, prepared=None, device_matrix=None,
                   host_mirror=None, mirror_threads=None, stats=None):
    """The whole symmetric n x n matrix of profiles X into the host tensor `host` (n x n, pinned).

    The device keeps the matrix resident (4 n^2 bytes); row panels are computed top to bottom, upper
    triangle + mirror.  As soon as panel p = rows [r0, r1) is done, two blocks are final and leave
    on the copy stream while panel p+1 computes: the panel's rows from column r0 on, and the
    mirrored column block [r1, n) x [r0, r1) below it.  What has become final is always
    proportional to what has been computed, so the PCIe link never waits for the expensive
    top panels (copying whole rows only, the first third of the matrix runs at kernel speed
    and the link idles).

    The mirrored column block is the transpose of the part of the panel right of its diagonal block,
    which has just arrived in `host`: a share `host_mirror` of it (its bottom rows; float32 only;
    default ``default_host_mirror_share()``) is not copied but built there by the threads of a
    HostMirror pool, released in stream order by a callback behind the panel's DMA -- 40 GB over a
    52 GB/s link is what bounds the step otherwise, the kernels need 0.5 s.  The call returns when the
    DMA is enqueued and the host's share is written (the pool is waited for); the caller synchronises
    the device as before.  Returns the number of bytes copied to the host over PCIe; `stats`, when
    given, receives {"dma_bytes", "host_mirrored_bytes", "mirror_threads"}."""
    device = require_cuda()
    P, aux, dim = prepared if prepared is not None else prepare(X, metric)
    n = int(P.shape[0])
    if tuple(host.shape) != (n, n) or host.dtype != out_dtype or not host.is_pinned():
        raise PhyloligoError("matrix_to_host: host must be a pinned (n, n) tensor of the output dtype")
    full = device_matrix if device_matrix is not None else torch.empty((n, n), dtype=out_dtype, device=device)
    if tuple(full.shape) != (n, n) or full.dtype != out_dtype:
        raise PhyloligoError("matrix_to_host: device_matrix must be (n, n) of the output dtype")
    share = default_host_mirror_share() if host_mirror is None else min(1.0, max(0.0, float(host_mirror)))
    if out_dtype != torch.float32 or host.stride(1) != 1:
        share = 0.0
    pool = host_mirror_pool(mirror_threads) if share > 0.0 else None
    step = max(TILE, (int(panel_rows) // TILE) * TILE)
    compute = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    copied = 0
    mirrored = 0
    try:
        for r0 in range(0, n, step):
            r1 = min(n, r0 + step)
            distance_block(metric, P, aux, dim, r0, r1, 0, n, full, 0, 0, FLAG_SKIP_LOWER | FLAG_MIRROR)
            ready = torch.cuda.Event()
            ready.record(compute)
            copy_stream.wait_event(ready)
            # in DMA_ROWS-row pieces: beside the host's mirroring short copies keep a little more of the link's rate
            for d0 in range(r0, r1, DMA_ROWS):
                d1 = min(r1, d0 + DMA_ROWS)
                copy2d(host[d0:d1, r0:], full[d0:d1, r0:], copy_stream)
            # rows [r1, rs) of the mirrored column block by DMA, rows [rs, n) by the host from host[r0:r1, rs:]
            rs = n - int(round(share * (n - r1))) if pool is not None else n
            if rs < n:
                pool.submit(host[rs:, r0:r1], host[r0:r1, rs:], copy_stream)
                mirrored += (n - rs) * (r1 - r0) * full.element_size()
            copy2d(host[r1:rs, r0:r1], full[r1:rs, r0:r1], copy_stream)
            copied += ((r1 - r0) * (n - r0) + (rs - r1) * (r1 - r0)) * full.element_size()
        compute.wait_stream(copy_stream)
        full.record_stream(copy_stream)
    finally:
        # also on an error: the pool must not be left writing into `host` behind the caller's back
        if pool is not None and mirrored:
            pool.wait()
    if stats is not None:
        stats.update(dma_bytes=copied, host_mirrored_bytes=mirrored, mirror_threads=pool.threads if pool else 0)
    return copied


# ----------------------------------------------------------------------------
# Host placement
# ----------------------------------------------------------------------------
def bind_host_to_gpu_node(device_index=None):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned buffers it
    allocates afterwards (first touch) sit in the memory the GPU's PCIe root complex writes to.  One
    process per GPU under torchrun otherwise floats over both sockets and half of the device-to-host
    traffic crosses the socket interconnect.  Returns the node number, or None when the topology is
    not visible (no sysfs entry, single node): then nothing is changed."""
    import os
    try:
        if device_index is None:
            device_index = torch.cuda.current_device()
        props = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
            cpus = set()
            for part in fh.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError, AssertionError):
        return None
