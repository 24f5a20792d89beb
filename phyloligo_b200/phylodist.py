"""Pair-level distance functions with the reference's names and argument meaning
(reference phylopackage/core/phylodist.py:36-85), executed by the CUDA tile kernel.

    KL(a, b)    sum a ln(a/b), NaN / Inf terms zeroed (1-D)     core/phylodist.py:18-34
    Eucl(a, b)  sqrt(sum (a-b)^2)                              core/phylodist.py:36-41
    JSD(a, b)   Jensen-Shannon divergence in nats               core/phylodist.py:43-68
                (1-D x 1-D -> scalar; 2-D x 2-D -> matrix whose rows index `b`)
    KT(a, b)    1 - Bio.Cluster 'k' distance = Kendall tau_b    core/phylodist.py:71-74
    BC(a, b)    Bray-Curtis sum|a-b| / sum|a+b|                 core/phylodist.py:76-79
    SC(a, b)    1 - Spearman rho (the reference body raises NameError; this is the
                intended meaning)                               core/phylodist.py:82-85

Inputs are host arrays; the work happens on the GPU.  No CPU fallback exists.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine


def _cross(a, b, metric):
    """Rows of `a` against rows of `b` -> (len(b), len(a)) like the reference's 2-D JSD."""
    device = engine.require_cuda()
    a = np.atleast_2d(np.asarray(a))
    b = np.atleast_2d(np.asarray(b))
    out_dtype = torch.float32 if (a.dtype == np.float32 and b.dtype == np.float32) else torch.float64
    X = torch.from_numpy(np.ascontiguousarray(np.vstack([b, a]).astype(np.float64 if out_dtype == torch.float64 else np.float32))).to(device)
    P, aux, dim = engine.prepare(X, metric)
    nb, na = b.shape[0], a.shape[0]
    out = torch.empty((nb, na), dtype=out_dtype, device=device)
    engine.distance_block(metric, P, aux, dim, 0, nb, nb, nb + na, out, 0, nb, 0)
    return out.cpu().numpy()


def _pair_or_cross(a, b, metric):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.ndim == 1 and b.ndim == 1:
        return engine.pair_distance(a, b, metric)
    return _cross(a, b, metric)


def KL(a, b):
    """Kullback-Leibler divergence of two profiles in nats, terms that are NaN or infinite (a zero
    on either side) dropped -- the 1-D branch of the reference (core/phylodist.py:20-24).  Its 2-D
    branch calls an un-imported helper (NameError) and is never reached from phyloligo.py."""
    from . import kount
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.ndim != 1 or b.ndim != 1:
        raise engine.PhyloligoError("KL takes two 1-D profiles (the reference's 2-D branch raises NameError)")
    return kount.KL(a, b)


def Eucl(a, b):
    return _pair_or_cross(a, b, "Eucl")


def JSD(a, b):
    return _pair_or_cross(a, b, "JSD")


def KT(a, b):
    return _pair_or_cross(a, b, "KT")


def BC(a, b):
    return _pair_or_cross(a, b, "BC")


def SC(a, b):
    return _pair_or_cross(a, b, "SC")


def weighted_rank(a, b):
    """Stub in the reference as well (core/phylodist.py:87-88)."""
    return 0
