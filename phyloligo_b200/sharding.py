"""Host-side partitioning of the two stages over the ranks of one node.

The reference splits profiling per sequence / per chunk of sequences
(bin/phyloligo.py:868, 913, 967) and the distance matrix in block rows
(``gen_even_slices(N, n_jobs)``, bin/phyloligo.py:424, 516).  Here:

* records go to ranks in contiguous, byte-balanced ranges (``record_cuts``);
* the matrix is split in contiguous block rows whose *upper-triangle* areas are
  balanced (``triangle_row_ranges``): a rank computes only the part of its block
  row on and right of the diagonal, mirrors its diagonal block locally, and hands
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


def triangle_row_ranges(n, world, align=128):
    """Row boundaries R[0..world] (multiples of `align` = the largest kernel tile edge, so that no
    tile straddles two ranks' rows; R[0]=0, R[world]=n) such that the
    number of matrix entries on or right of the diagonal is about equal in every block row
    [R[s], R[s+1]).  The area above row x is x*n - x^2/2, so R[s] = n (1 - sqrt(1 - s/world))."""
    bounds = [0]
    for s in range(1, world):
        x = n * (1.0 - math.sqrt(1.0 - s / world))
        x = int(round(x / align)) * align
        bounds.append(min(n, max(bounds[-1], x)))
    bounds.append(n)
    return bounds


def upper_area(bounds, s, n):
    """Entries on or right of the diagonal in block row s."""
    a, b = bounds[s], bounds[s + 1]
    return (b - a) * (n - a) - (b - a) * (b - a - 1) // 2


def exchange_transposed(T, bounds, rank, world, out_rows, group=None):
    """Fill the columns left of this rank's diagonal block from its peers.

    T         [(n - R[rank+1]) x rows] buffer this rank computed with the mirror output:
              T[c - R[rank+1], r - R[rank]] = D[r, c] for its rows r and every column c of a later rank
    out_rows  [rows x n] this rank's block row; columns [R[s], R[s+1]) for s < rank are written here
    Rank s sends rank d > s the contiguous slab T[R[d] - R[s+1] : R[d+1] - R[s+1]], which is exactly
    D[rows of d, rows of s]; rank d drops it into its column range of s.
    """
    a, b = bounds[rank], bounds[rank + 1]
    rows = b - a
    ops, recvs = [], []
    if rows > 0:
        for d in range(rank + 1, world):
            lo, hi = bounds[d] - b, bounds[d + 1] - b
            if hi > lo:
                ops.append(dist.P2POp(dist.isend, T[lo:hi], d, group=group))
    for s in range(rank):
        cols = bounds[s + 1] - bounds[s]
        if cols > 0 and rows > 0:
            buf = torch.empty((rows, cols), dtype=out_rows.dtype, device=out_rows.device)
            recvs.append((s, buf))
            ops.append(dist.P2POp(dist.irecv, buf, s, group=group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for s, buf in recvs:
        out_rows[:, bounds[s]:bounds[s + 1]].copy_(buf)
    return out_rows
