"""The distance stage on the GPUs of one node: one process per GPU, paired block rows.

The reference splits the matrix in block rows over its workers and computes every
entry of a block row (``gen_even_slices``, bin/phyloligo.py:424, 516; workers
``distances_loc`` :195 / ``distances_h5py`` :233).  Here rank s owns block rows s and
2W-1-s (``sharding.paired_row_ranges``), computes only the tiles on or right of the
diagonal, and every off-diagonal tile is stored twice by the CTA that computed it:
into this rank's rows, and transposed into the rows of the rank that owns the
mirrored entries.

exchange = "peer" (default on NVLink boxes): the owners' row buffers are mapped into
every process with CUDA IPC (``engine.PeerRows``) and the tile kernel's mirror stores
go straight to peer memory over NVLink / NVSwitch while the tile math runs -- the
exchange is fused into the compute kernel, there is no staging buffer and no
transfer step.  One tiny all-reduce at the end is the device-side barrier that makes
every rank's rows complete.

exchange = "nccl": the transposed tiles go to a local staging buffer and one
``batch_isend_irecv`` step moves them (``sharding.exchange_transposed``); this is the
path the gloo CPU tests exercise and the fallback when peer mapping is unavailable.

Both give bit for bit the rows of a single-GPU run: same kernels, same tile grid.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from . import engine, sharding
from ._lib import FLAG_MIRROR, FLAG_SKIP_LOWER, TILE


def default_exchange():
    return os.environ.get("PO_EXCHANGE", "peer")


class BlockRows:
    """This rank's paired block rows of the symmetric n x n matrix, resident on the device.

    matrix       [rows_owned x n] tensor: the owned ranges stacked in ascending order
    out_rows[i]  view of the rows of owned range i
    """

    def __init__(self, n, out_dtype, rank, world, exchange=None, device=None):
        self.n, self.rank, self.world = int(n), rank, world
        self.out_dtype = out_dtype
        self.device = device or engine.require_cuda()
        self.exchange = exchange or default_exchange()
        self.ranges = sharding.paired_row_ranges(self.n, world)
        self.offsets = sharding.range_offsets(self.ranges, world)
        self.my_ranges = [i for i in sharding.owned_ranges(self.ranges, rank, world)
                          if self.ranges[i][1] > self.ranges[i][0]]
        self.rows_owned = sum(self.ranges[i][1] - self.ranges[i][0] for i in self.my_ranges)
        self.matrix = torch.empty((max(1, self.rows_owned), self.n), dtype=out_dtype, device=self.device)
        self.out_rows = {}
        for i in self.my_ranges:
            a, b = self.ranges[i]
            self.out_rows[i] = self.matrix[self.offsets[i]:self.offsets[i] + (b - a)]
        self.esize = self.matrix.element_size()
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.peers = None
        self.staging = None
        self._copy_stream = None
        if world > 1 and self.exchange == "peer":
            # mapping can be refused (no peer access between two GPUs, an allocator that hands out
            # unexportable memory): then every rank falls back to the NCCL exchange together
            try:
                self.peers = engine.PeerRows(self.matrix, rank, world)
                ok = 1
            except engine.PhyloligoError as exc:
                self.peers, ok = None, 0
                self._peer_error = str(exc)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                if self.peers is not None:
                    self.peers.close()
                    self.peers = None
                self.exchange = "nccl"
        if world > 1 and self.peers is None:
            self.staging = {}
            for i in self.my_ranges:
                a, b = self.ranges[i]
                self.staging[i] = torch.empty((max(1, self.n - b), b - a), dtype=out_dtype, device=self.device)

    def close(self):
        if self.peers is not None:
            torch.cuda.synchronize()
            if self.world > 1:
                dist.barrier()  # nobody unmaps while a peer may still be storing
            self.peers.close()
            self.peers = None

    def _device_barrier(self):
        """Stream-ordered barrier over the ranks: the all-reduce on a rank completes only after every
        rank has reached it on its stream, i.e. after every rank's tile kernels (and their peer stores,
        which are performed by kernel completion) are done."""
        if self.world > 1:
            dist.all_reduce(self._flag)

    def upper_area(self):
        return sharding.upper_area(self.ranges, self.rank, self.world, self.n)

    def compute(self, metric, P, aux, dim, host_rows=None, ship=None, left_parts=True, panel_rows=None):
        """Launch this rank's tiles; on return (in stream order) `matrix` holds its complete rows.

        With `host_rows` (a pinned [rows_owned x n] tensor) the rows also go to the host: the part of
        a block row from its diagonal block rightwards is written by this rank only, so it leaves on
        the copy stream as soon as that block row's launches are done, overlapping the next block
        row's tiles; the part left of the diagonal block is written by the other ranks and leaves
        after the closing barrier.  `ship(block, row0, col0)` instead hands every such finished
        device block (a view of `matrix`, with the matrix coordinates of its corner) to the caller in
        the same order -- the command line's file sink (hostsink.RowShipper).  With
        ``left_parts=False`` only the parts from the diagonal block rightwards are handed over: the
        caller builds the rest from them (``MirroredHostSink``: the matrix is symmetric, and the part of
        a block row left of its diagonal block is the transpose of right parts that other ranks ship).
        With `panel_rows` a block row is launched and shipped in row panels of that many rows (a multiple
        of the largest tile): a panel's finished part leaves while the next panel computes, instead of a
        whole block row's -- with two block rows per rank, the first of them most of the rank's work,
        shipping by whole block rows serialises compute and copy.  The tiles and their values do not
        depend on the split.  Returns `matrix`."""
        n = self.n
        if host_rows is not None:
            if tuple(host_rows.shape) != tuple(self.matrix.shape) or not host_rows.is_pinned():
                raise RuntimeError("host_rows must be a pinned tensor of the shape of BlockRows.matrix")
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            compute_stream = torch.cuda.current_stream()

            def ship(block, row0, col0, _host=host_rows, _cs=compute_stream):  # noqa: F811
                i = next(k for k in self.my_ranges if self.ranges[k][0] <= row0 < self.ranges[k][1])
                off = self.offsets[i] + (row0 - self.ranges[i][0])
                ready = torch.cuda.Event()
                ready.record(_cs)
                self._copy_stream.wait_event(ready)
                engine.copy2d(_host[off:off + block.shape[0], col0:col0 + block.shape[1]], block, self._copy_stream)

        def ship_panel(i, r0, r1):
            """what has just become final of rows [r0, r1) of block row i"""
            a, b = self.ranges[i]
            rows = self.out_rows[i][r0 - a:r1 - a]
            if left_parts:
                ship(rows[:, a:], r0, a)    # earlier panels of the block row mirrored into columns [a, r0)
            else:
                ship(rows[:, r0:], r0, r0)  # the caller mirrors everything right of the panel's own square

        def ship_left(i):
            a, b = self.ranges[i]
            if a > 0:
                ship(self.out_rows[i][:, :a], a, 0)

        step = None
        if panel_rows:
            step = max(TILE, (int(panel_rows) // TILE) * TILE)
        if self.peers is not None:
            self._device_barrier()  # the consumers of the previous result are done with the rows
        for i in self.my_ranges:
            a, b = self.ranges[i]
            rows = self.out_rows[i]
            for r0 in range(a, b, step or (b - a)):
                r1 = min(b, r0 + (step or (b - a)))
                # the panel's part of the diagonal block: tiles on or right of the diagonal, mirrored in place
                engine.distance_block(metric, P, aux, dim, r0, r1, r0, b, rows, a, 0, FLAG_SKIP_LOWER | FLAG_MIRROR)
                if b < n and self.staging is not None:
                    engine.distance_block(metric, P, aux, dim, r0, r1, b, n, rows, a, 0, FLAG_MIRROR,
                                          mirror=self.staging[i], mirror_row0=b, mirror_col0=a)
                elif b < n:
                    # one launch per block right of the diagonal; its transposed tiles are stored into the
                    # rows of the rank that owns that range (local or peer address, row pitch n)
                    for q in range(i + 1, len(self.ranges)):
                        aq, bq = self.ranges[q]
                        if bq <= aq:
                            continue
                        owner = sharding.range_owner(q, self.world)
                        base = self.matrix.data_ptr() if owner == self.rank else self.peers.address(owner)
                        addr = base + self.offsets[q] * n * self.esize
                        engine.distance_block(metric, P, aux, dim, r0, r1, aq, bq, rows, a, 0, FLAG_MIRROR,
                                              mirror=addr, mirror_row0=aq, mirror_col0=0, mirror_ld=n)
                if ship is not None:
                    ship_panel(i, r0, r1)
        if self.peers is not None:
            self._device_barrier()
        elif self.world > 1:
            sharding.exchange_transposed(self.staging, self.ranges, self.rank, self.world, self.out_rows)
        if ship is not None and left_parts:
            for i in self.my_ranges:
                ship_left(i)
        if host_rows is not None:
            torch.cuda.current_stream().wait_stream(self._copy_stream)
        return self.matrix


class MirroredHostSink:
    """The host end of a multi-GPU run that lets only the upper triangle cross PCIe.

    `host` is the whole n x n float32 matrix in host memory that every rank of the node maps (a file
    under /dev/shm: ``hostsink.FileMatrix``; a rank's own rows page-locked so that DMA lands in them).
    ``ship`` (the callback of ``BlockRows.compute(..., ship=sink.ship, left_parts=False)``) takes the
    part of a block row (or of a row panel of it) from its diagonal square rightwards, sends it to the rank's
    rows of `host` by strided DMA, and queues behind it (or behind every `sub_rows` rows of it) the transposition
    of its columns right of that square into the rows below (``engine.HostMirror``, released in
    stream order) -- rows that other ranks own.  Every entry left of the diagonal is therefore written
    by the rank that computed its mirror image, from host memory, and no rank ships the left part of
    its rows: half of the bytes cross PCIe.  The regions the ranks write are disjoint; the matrix is
    complete when every rank has called ``finish`` (a barrier is the caller's).  The reference's
    workers assign whole block rows (output[s] = ..., bin/phyloligo.py:202-222)."""

    def __init__(self, host, pool, sub_rows=None):
        if host.dim() != 2 or host.shape[0] != host.shape[1] or host.dtype != torch.float32 or host.is_cuda:
            raise RuntimeError("MirroredHostSink: host must be a square float32 host tensor")
        self.host, self.pool = host, pool
        self.n = int(host.shape[0])
        # rows per mirror submit = per stream callback (None: a whole shipped block; a callback stalls the copy
        # stream for ~0.2 ms, profiles/r02e_sink_callback_granularity_probe.log)
        self.sub = max(1, int(sub_rows)) if sub_rows else None
        self._copy_stream = None
        self.reset()

    def reset(self):
        """Zero the byte counters (a sink serves many steps)."""
        self.dma_bytes = 0
        self.mirrored_bytes = 0

    def ship(self, block, row0, col0):
        if col0 != row0 or block.shape[1] != self.n - col0:
            raise RuntimeError("MirroredHostSink.ship: expected the part of a block row from its diagonal block rightwards")
        h = int(block.shape[0])
        b = row0 + h
        on_device = block.is_cuda
        if on_device:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream()
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            self._copy_stream.wait_event(ready)
        sub = self.sub or h
        for r0 in range(0, h, sub):
            r1 = min(h, r0 + sub)
            dst = self.host[row0 + r0:row0 + r1, col0:]
            if on_device:
                for d0 in range(r0, r1, engine.DMA_ROWS):
                    d1 = min(r1, d0 + engine.DMA_ROWS)
                    engine.copy2d(self.host[row0 + d0:row0 + d1, col0:], block[d0:d1], self._copy_stream)
            else:  # the CPU stand-in of the gloo tests
                dst.copy_(block[r0:r1])
            self.dma_bytes += (r1 - r0) * (self.n - col0) * 4
            if b < self.n:
                self.pool.submit(self.host[b:, row0 + r0:row0 + r1], self.host[row0 + r0:row0 + r1, b:],
                                 self._copy_stream, after_stream=on_device)
                self.mirrored_bytes += (r1 - r0) * (self.n - b) * 4

    def finish(self):
        """Stream-order the copies before whatever follows on the current stream, and wait for this rank's
        share of the mirroring."""
        if self._copy_stream is not None:
            torch.cuda.current_stream().wait_stream(self._copy_stream)
        self.pool.wait()
