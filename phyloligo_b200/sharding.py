"""Host-side partitioning of the two stages over the ranks of one node.

The reference splits profiling per sequence / per chunk of sequences
(bin/phyloligo.py:868, 913, 967) and the distance matrix in block rows
(``gen_even_slices(N, n_jobs)``, bin/phyloligo.py:424, 516).  Here:

* records go to ranks in contiguous, byte-balanced ranges (``record_cuts``);
* the matrix is split in 2 x world contiguous block rows; rank s owns block rows s
  and 2 x world - 1 - s (``paired_row_ranges``), which balances both the rows per rank
  and the upper-triangle area per rank.  A rank computes only the part of its block
  rows on and right of the diagonal, mirrors its diagonal blocks locally, and hands
  the transposed off-diagonal blocks to the ranks that own those columns as rows
  (``exchange_transposed``) -- the one exchange step of the distance stage.

Nothing here touches a GPU; the functions work on any torch device and any
``torch.distributed`` backend (the CPU tests run them under gloo).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.distributed as dist


def record_cuts(lengths, world):
    """Cut indices c[0..world] so that records [c[r], c[r+1]) hold ~1/world of the bytes."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = int(lengths.shape[0])
    if n == 0:
        return [0] * (world + 1)
    cum = np.cumsum(lengths)
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(cum, total * r / world)))
    cuts.append(n)
    for i in range(1, len(cuts)):  # monotone
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def paired_row_ranges(n, world, align=128):
    """Block rows of the symmetric matrix for `world` ranks: 2*world contiguous ranges of (about) equal
    height, boundaries multiples of `align` (= the largest kernel tile edge, so that no tile straddles
    two ranges).  Rank s owns range s and range 2*world-1-s: a range near the top has many columns right
    of the diagonal, its partner near the bottom few, so every rank gets the same number of rows (equal
    device-to-host volume) AND the same upper-triangle area (equal compute).
    Returns the list of (start, stop) of the 2*world ranges."""
    parts = 2 * world
    bounds = [0]
    for i in range(1, parts):
        x = int(round(n * i / parts / align)) * align
        bounds.append(min(n, max(bounds[-1], x)))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(parts)]


def range_owner(i, world):
    return i if i < world else 2 * world - 1 - i


def owned_ranges(ranges, rank, world):
    """Indices of the ranges a rank owns, ascending."""
    return [i for i in range(len(ranges)) if range_owner(i, world) == rank]


def upper_area(ranges, rank, world, n):
    """Entries on or right of the diagonal in the block rows of `rank`."""
    total = 0
    for i in owned_ranges(ranges, rank, world):
        a, b = ranges[i]
        total += (b - a) * (n - a) - (b - a) * (b - a - 1) // 2
    return total


def exchange_transposed(T, ranges, rank, world, out_rows, group=None):
    """Fill the columns left of each owned diagonal block from the ranges that computed them.

    For an owned range R = [a, b):
      T[R]         [(n - b) x (b - a)] buffer the rank computed with the mirror output:
                   T[R][c - b, r - a] = D[r, c] for r in R and every column c right of R
      out_rows[R]  [(b - a) x n] the block row itself; its columns [a', b') for every range
                   Q' = [a', b') left of R are written here
    The owner of R sends, for every range Q right of R, the contiguous slab
    T[R][a_Q - b : b_Q - b] (= D[Q, R]) to the owner of Q, which drops it into out_rows[Q][:, a:b].
    T and out_rows are dicts keyed by range index.  Messages between two ranks are issued in
    (R, Q) lexicographic order on both sides, which is how NCCL / gloo match them.
    """
    mine = set(owned_ranges(ranges, rank, world))
    ops, recvs, local = [], [], []
    pairs = [(r, q) for r in range(len(ranges)) for q in range(r + 1, len(ranges))
             if ranges[r][1] > ranges[r][0] and ranges[q][1] > ranges[q][0]]
    for r, q in pairs:
        a, b = ranges[r]
        aq, bq = ranges[q]
        src, dst = range_owner(r, world), range_owner(q, world)
        if src == rank and dst == rank:
            local.append((r, q))
        elif src == rank:
            ops.append(dist.P2POp(dist.isend, T[r][aq - b:bq - b], dst, group=group))
        elif dst == rank:
            buf = torch.empty((bq - aq, b - a), dtype=out_rows[q].dtype, device=out_rows[q].device)
            recvs.append((r, q, buf))
            ops.append(dist.P2POp(dist.irecv, buf, src, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for r, q in local:
        a, b = ranges[r]
        aq, bq = ranges[q]
        out_rows[q][:, a:b].copy_(T[r][aq - b:bq - b])
    for r, q, buf in recvs:
        a, b = ranges[r]
        out_rows[q][:, a:b].copy_(buf)
    assert all(i in mine for i in out_rows)
    return out_rows


def range_offsets(ranges, world):
    """Row offset of every range inside its owner's stacked row buffer (owned ranges in ascending order)."""
    offsets = [0] * len(ranges)
    for rank in range(world):
        off = 0
        for i in owned_ranges(ranges, rank, world):
            offsets[i] = off
            off += ranges[i][1] - ranges[i][0]
    return offsets
