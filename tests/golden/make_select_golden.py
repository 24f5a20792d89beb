#!/usr/bin/env python3
"""Generate tests/golden/select_golden.npz by running the REFERENCE's own KMedoids class
(AST-extracted from /root/reference/phylopackage/bin/phyloselect.py:37-309; nothing is copied) on
seeded distance matrices.

    python tests/golden/make_select_golden.py

The module's top-level imports (matplotlib, Bio, hdbscan, h5py) are absent here and are not needed by
the class: only its ClassDef is executed, in a namespace that provides the scikit-learn names it uses.
"""
import ast
import os
import sys
import warnings

import numpy as np
from sklearn.base import BaseEstimator, ClusterMixin, TransformerMixin
from sklearn.metrics.pairwise import PAIRWISE_DISTANCE_FUNCTIONS
from sklearn.utils import check_array, check_random_state
from sklearn.utils.validation import check_is_fitted

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import phylo_oracle as po  # noqa: E402
from phyloligo_b200 import synth  # noqa: E402

SRC = os.path.join(os.environ.get("PHYLOLIGO_REFERENCE", "/root/reference"), "phylopackage", "bin", "phyloselect.py")


def load_kmedoids():
    tree = ast.parse(open(SRC).read(), filename=SRC)
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "KMedoids"]
    ns = dict(np=np, warnings=warnings, BaseEstimator=BaseEstimator, ClusterMixin=ClusterMixin, TransformerMixin=TransformerMixin,
              PAIRWISE_DISTANCE_FUNCTIONS=PAIRWISE_DISTANCE_FUNCTIONS, check_array=check_array,
              check_random_state=check_random_state, check_is_fitted=check_is_fitted)
    exec(compile(ast.Module(body=cls, type_ignores=[]), SRC, "exec"), ns)
    return ns["KMedoids"]


def matrices():
    rng = np.random.default_rng(12)
    # (1) JSD matrix of synthetic contigs (two GC populations): the use case of the tool
    seqs = synth.make_sequences(140, 4000, seed=31)
    X = np.vstack([po.frequency_np(s, "1111", "both") for s in seqs])
    yield "jsd_contigs", po.pairwise_np(X, "JSD"), dict(n_clusters=2)
    yield "jsd_contigs_k5", po.pairwise_np(X, "JSD"), dict(n_clusters=5)
    # (2) Euclidean matrix of three Gaussian blobs, float32 as the --large memmap files are
    P = np.vstack([rng.normal(c, 0.6, size=(50, 4)) for c in (0.0, 3.0, 6.0)])
    E = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    yield "blobs_f32", E, dict(n_clusters=3)
    yield "blobs_random_init", E.astype(np.float64), dict(n_clusters=4, init="random", random_state=0)
    # (3) duplicated points: ties in the assignment, an empty cluster
    Q = np.repeat(rng.random((12, 3)), 5, axis=0)
    T = np.sqrt(((Q[:, None, :] - Q[None, :, :]) ** 2).sum(-1))
    yield "duplicates", T, dict(n_clusters=6)
    yield "max_iter_1", E.astype(np.float64), dict(n_clusters=3, max_iter=1)


def main():
    KMedoids = load_kmedoids()
    out = {}
    names = []
    for name, D, kw in matrices():
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            km = KMedoids(distance_metric="precomputed", **kw).fit(D)
        medoids = np.array([int(np.flatnonzero((D == c).all(axis=1))[0]) for c in km.cluster_centers_])
        out[name + "_D"] = D
        out[name + "_labels"] = np.asarray(km.labels_, dtype=np.int64)
        out[name + "_medoids"] = medoids
        out[name + "_n_iter"] = np.int64(km.n_iter_)
        out[name + "_empty_warnings"] = np.int64(sum("is empty" in str(w.message) for w in caught))
        out[name + "_kwargs"] = np.array(repr(sorted(kw.items())))
        names.append(name)
        print(name, D.shape, D.dtype, kw, "n_iter", km.n_iter_, "medoids", medoids.tolist(), "sizes", np.bincount(km.labels_).tolist())
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "select_golden.npz"), **out)


if __name__ == "__main__":
    main()
