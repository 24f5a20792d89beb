#!/usr/bin/env python3
"""The front half of phyloselect.py on B200: K-medoids on the contig distance
matrix and the nearest-neighbour graph its t-SNE / HDBSCAN consumers start from, computed on the
device while the matrix is resident -- straight after the distance stage if wanted, without the
round trip through a 40 GB file and 800 GB of host RAM the reference's documentation asks for.

Mirrors reference phylopackage/bin/phyloselect.py: ``KMedoids`` (:37-309; same constructor
arguments and fitted attributes), ``clusterize`` (:573-590), ``find_clusters`` (:403-428), the
matrix readers of ``main`` (:597-622), ``write_fastafile`` (:547-571) and the
``data_cluster_indexes.dat`` writer (:742-751).  Plotting, t-SNE itself, the interactive loop and
HDBSCAN (third-party, not installed) are out of scope: ``knn_graph`` hands those consumers the
sparse precomputed neighbour graph scikit-learn's TSNE accepts in place of the dense matrix.

There is no CPU fallback: the sums, assignments and selections run in libphyloligo_b200.so.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys
import warnings

import numpy as np
import torch

from . import _lib, engine, io_formats
from ._lib import PhyloligoError, PO_F32, PO_F64


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dt(D):
    if D.dtype == torch.float32:
        return PO_F32
    if D.dtype == torch.float64:
        return PO_F64
    raise PhyloligoError("the distance matrix must be float32 or float64")


def as_device_matrix(D):
    """A 2-D float32 / float64 device tensor with contiguous rows from an ndarray / memmap / tensor."""
    device = engine.require_cuda()
    if isinstance(D, torch.Tensor):
        t = D.to(device)
    else:
        a = np.asarray(D)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    if t.dim() != 2 or t.shape[0] != t.shape[1]:
        raise PhyloligoError("a square distance matrix is expected, got shape %s" % (tuple(t.shape),))
    if t.stride(1) != 1:
        t = t.contiguous()
    _dt(t)
    return t


def row_sums(D, rows=None, labels=None, row_labels=None):
    """float64 sums of rows of the device matrix D (po_matrix_rowsums): all columns, or only the columns j
    with labels[j] == row_labels[i]."""
    lib = _lib.load()
    n_cols = int(D.shape[1])
    n_rows = int(rows.shape[0]) if rows is not None else int(D.shape[0])
    out = torch.empty(n_rows, dtype=torch.float64, device=D.device)
    rc = lib.po_matrix_rowsums(_ptr(D), int(D.stride(0)), _dt(D), _ptr(rows), n_rows, n_cols, _ptr(labels), _ptr(row_labels),
                               _ptr(out), _stream())
    _lib.check(rc, "po_matrix_rowsums")
    return out


def assign_to_medoids(D, medoids):
    """labels[j] = argmin_c D[medoids[c], j] (po_matrix_argmin_rows), int32."""
    lib = _lib.load()
    n = int(D.shape[1])
    out = torch.empty(n, dtype=torch.int32, device=D.device)
    rc = lib.po_matrix_argmin_rows(_ptr(D), int(D.stride(0)), _dt(D), _ptr(medoids), int(medoids.shape[0]), n, _ptr(out), _stream())
    _lib.check(rc, "po_matrix_argmin_rows")
    return out


def cluster_argmin(cost, labels, k):
    """(best_idx int64[k], best_cost float64[k], count int64[k]) per cluster (po_cluster_argmin)."""
    lib = _lib.load()
    dev = cost.device
    work = torch.empty(2 * k, dtype=torch.int64, device=dev)
    best_idx = torch.empty(k, dtype=torch.int64, device=dev)
    best_cost = torch.empty(k, dtype=torch.float64, device=dev)
    count = torch.empty(k, dtype=torch.int64, device=dev)
    rc = lib.po_cluster_argmin(_ptr(cost), _ptr(labels), int(cost.shape[0]), int(k), _ptr(work), _ptr(best_idx), _ptr(best_cost),
                               _ptr(count), _stream())
    _lib.check(rc, "po_cluster_argmin")
    return best_idx, best_cost, count


def knn_graph(D, k, row0=0, as_sparse=False):
    """The k nearest neighbours of every row of the device matrix D (a block row when row0 > 0: row i is
    profile row0 + i and that column is excluded), ascending, ties by column (po_matrix_knn).
    Returns (indices int32 [n, k], distances float32 [n, k]) device tensors, or with `as_sparse` the
    scipy CSR matrix that sklearn's TSNE(metric="precomputed") / kneighbors_graph(mode="distance") use."""
    lib = _lib.load()
    D = D if isinstance(D, torch.Tensor) and D.is_cuda else as_device_matrix(D)
    n_rows, n_cols = int(D.shape[0]), int(D.shape[1])
    idx = torch.empty((n_rows, k), dtype=torch.int32, device=D.device)
    dist = torch.empty((n_rows, k), dtype=torch.float32, device=D.device)
    rc = lib.po_matrix_knn(_ptr(D), int(D.stride(0)), _dt(D), n_rows, n_cols, int(row0), int(k), _ptr(idx), _ptr(dist), _stream())
    _lib.check(rc, "po_matrix_knn")
    if not as_sparse:
        return idx, dist
    from scipy.sparse import csr_matrix
    indptr = np.arange(0, n_rows * k + 1, k)
    return csr_matrix((dist.cpu().numpy().ravel().astype(np.float64), idx.cpu().numpy().ravel(), indptr), shape=(n_rows, n_cols))


class KMedoids:
    """k-medoids (PAM) on the device.  Same parameters and fitted attributes as the reference class
    (bin/phyloselect.py:37-309): ``labels_``, ``cluster_centers_`` (the medoids' rows of X), ``n_iter_``;
    ``medoid_indices_`` in addition.  ``distance_metric`` is "precomputed" (X is the N x N matrix: ndarray,
    memmap or a device tensor, which is used in place) or one of the package's metrics
    (Eucl, JSD, KT, BC, SC: X is the N x D profile matrix and the distances never leave the device)."""

    CLUSTERING_METHODS = ["pam"]
    INIT_METHODS = ["random", "heuristic"]

    def __init__(self, n_clusters=8, distance_metric="precomputed", clustering_method="pam", init="heuristic", max_iter=300,
                 random_state=None):
        self.n_clusters = n_clusters
        self.distance_metric = distance_metric
        self.init = init
        self.max_iter = max_iter
        self.clustering_method = clustering_method
        self.random_state = random_state

    def _check_init_args(self):
        if self.n_clusters is None or not isinstance(self.n_clusters, int) or self.n_clusters <= 0:
            raise ValueError("n_clusters has to be nonnegative integer")
        if self.distance_metric != "precomputed" and self.distance_metric not in _lib.METRICS:
            raise ValueError("distance_metric needs to be 'precomputed' or one of {}. Instead, '{}' was given.".format(
                sorted(_lib.METRICS), self.distance_metric))
        if self.clustering_method not in self.CLUSTERING_METHODS:
            raise ValueError("clustering must be one of the following: {}".format(self.CLUSTERING_METHODS))
        if self.init not in self.INIT_METHODS:
            raise ValueError("init needs to be one of the following: {}".format(self.INIT_METHODS))

    def fit(self, X, y=None):
        self._check_init_args()
        if self.distance_metric == "precomputed":
            D = as_device_matrix(X)
        else:
            Xd = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(X)))
            D = engine.distance_matrix_device(Xd.to(engine.require_cuda()), self.distance_metric, torch.float32, symmetric=True)
        n, k = int(D.shape[0]), self.n_clusters
        if k > n:
            raise ValueError("The number of medoids ({}) must be larger than the number of samples ({})".format(k, n))
        dev = D.device
        # initial medoids (:291-309)
        if self.init == "random":
            rs = self.random_state if isinstance(self.random_state, np.random.RandomState) else np.random.RandomState(self.random_state)
            medoids = torch.from_numpy(np.asarray(rs.permutation(n)[:k], dtype=np.int64)).to(dev)
        else:
            medoids = torch.argsort(row_sums(D), stable=True)[:k].contiguous()
        cluster_ids = torch.arange(k, dtype=torch.int32, device=dev)
        old = None
        labels = None
        self.n_iter_ = 0
        while (old is None or not torch.equal(old, medoids)) and self.n_iter_ < self.max_iter:
            self.n_iter_ += 1
            old = medoids.clone()
            labels = assign_to_medoids(D, medoids)                                   # :187-195
            cost = row_sums(D, labels=labels, row_labels=labels)                     # every point's cost inside its own cluster
            curr = row_sums(D, rows=medoids, labels=labels, row_labels=cluster_ids)  # the current medoids' costs (:207-211)
            best_idx, best_cost, count = cluster_argmin(cost, labels, k)             # :213-227
            empty = torch.nonzero(count == 0).flatten().tolist()
            for c in empty:
                warnings.warn("Cluster {} is empty!".format(c))
            better = (count > 0) & (best_cost < curr)                                # :229-240
            medoids = torch.where(better, best_idx, medoids)
        self.medoid_indices_ = medoids.cpu().numpy()
        self.labels_ = labels.cpu().numpy().astype(np.int64) if labels is not None else np.zeros(n, dtype=np.int64)
        if isinstance(X, torch.Tensor):
            self.cluster_centers_ = X[medoids.to(X.device)].cpu().numpy()
        else:
            self.cluster_centers_ = np.asarray(X).take(self.medoid_indices_, axis=0)
        return self


# ---------------------------------------------------------------------------
# the script's functions
# ---------------------------------------------------------------------------
def read_distmat(path):
    """reference :357-371"""
    return np.loadtxt(path)


def read_matrix(path, large=False):
    """The three on-disk conventions of phyloligo.py (reference main :601-622)."""
    if large == "memmap":
        matrix = np.memmap(path, dtype=np.float32, mode="r")
        n = int(round(np.sqrt(matrix.shape[0])))
        if n * n != matrix.shape[0]:
            print("Error, weird shape for matrix {}".format(path), file=sys.stderr)
            sys.exit(1)
        return matrix.reshape((n, n))
    if large == "h5py":
        return io_formats.read_hdf5(path, "distances")
    return read_distmat(path)


def find_clusters(data, method, kwargs):
    """reference :403-428"""
    if method == "kmedoids":
        return KMedoids(**kwargs).fit(data).labels_
    if method == "hdbscan":
        try:
            import hdbscan
        except ImportError:
            print("Error, method hdbscan needs the hdbscan package, which is not installed; "
                  "knn_graph() provides its neighbour graph", file=sys.stderr)
            sys.exit(1)
        host = data.cpu().numpy() if isinstance(data, torch.Tensor) else np.asarray(data)
        clusterer = hdbscan.HDBSCAN(**kwargs)
        clusterer.fit(host.astype(np.float64))
        return clusterer.labels_
    print("Error, unknown method {}".format(method), file=sys.stderr)
    sys.exit(1)


def clusterize(data, method, min_cluster_size=None, min_samples=None, nbk=None):
    """reference :573-590"""
    kwargs = dict()
    if method == "hdbscan":
        if min_cluster_size is not None:
            kwargs["min_cluster_size"] = min_cluster_size
        if min_samples is not None:
            kwargs["min_samples"] = min_samples
        kwargs["metric"] = "precomputed"
    if method == "kmedoids":
        if nbk is not None:
            kwargs["n_clusters"] = nbk
        kwargs["distance_metric"] = "precomputed"
    return find_clusters(data, method, kwargs)


def write_cluster_indexes(labels_pred, pathout):
    """``data_cluster_indexes.dat``: one "class index" line per contig, grouped by class (reference :742-751)."""
    labels_pred = np.asarray(labels_pred)
    with open(pathout, "w") as outf:
        for cl in np.unique(labels_pred):
            for idx in np.where(labels_pred == cl)[0]:
                outf.write("{} {}\n".format(cl, idx))


def write_fastafile(labels_pred, fastafile, outputdir):
    """One FASTA file per class, ``data_fasta_cl{c}.fa`` / ``data_fasta_unclust.fa`` (reference :547-571).
    Records are written as Biopython's SeqIO.write does: the title line unchanged, the sequence wrapped at
    60 columns."""
    labels_pred = np.asarray(labels_pred)
    text = np.fromfile(fastafile, dtype=np.uint8)
    begin, end = engine.fasta_index(text)
    if len(begin) != len(labels_pred):
        raise PhyloligoError("{} holds {} records, the matrix {} rows".format(fastafile, len(begin), len(labels_pred)))
    raw = text.tobytes()
    starts = engine.fasta_header_starts(raw, begin, end)
    for cl in np.unique(labels_pred):
        name = "data_fasta_unclust.fa" if cl == -1 else "data_fasta_cl{}.fa".format(cl)
        with open(os.path.join(outputdir, name), "wb") as outf:
            for idx in np.where(labels_pred == cl)[0]:
                b, e = int(begin[idx]), int(end[idx])
                title = raw[int(starts[idx]):b].rstrip(b"\r\n")
                seq = b"".join(raw[b:e].split())
                outf.write(title + b"\n")
                for p in range(0, len(seq), 60):
                    outf.write(seq[p:p + 60] + b"\n")


def get_cmd(argv=None):
    parser = argparse.ArgumentParser(description="Cluster contigs from their oligonucleotide-profile distances on B200 GPUs")
    parser.add_argument("-i", action="store", dest="distmat", help="The input matrix file")
    parser.add_argument("-m", action="store", dest="method", required=True, choices=["hdbscan", "kmedoids"],
                        help="Method to use to compute cluster on the distance matrix")
    parser.add_argument("--minclustersize", action="store", dest="min_cluster_size", type=int,
                        help="Set the minimal cluster size of an HDBSCAN cluster")
    parser.add_argument("--minsamples", action="store", dest="min_samples", type=int,
                        help="Set the minimal sample size of an HDBSCAN cluster")
    parser.add_argument("-k", action="store", dest="nbk", type=int, help="Number of cluster")
    parser.add_argument("-f", action="store", dest="fastafile",
                        help="Path of the original fasta file used for the computation of the distance matrix")
    parser.add_argument("--large", action="store", choices=["memmap", "h5py"], dest="large", default=False,
                        help="Format of a matrix written with phyloligo.py --large")
    parser.add_argument("-o", action="store", dest="outputdir", required=True)
    parser.add_argument("-t", action="store_true", dest="performtsne", default=False, help="(not built: plotting is out of scope)")
    parser.add_argument("--interactive", action="store_true", dest="interactive", default=False,
                        help="(not built: plotting is out of scope)")
    parser.add_argument("--noX", action="store_true", dest="noX", help="accepted for compatibility")
    parser.add_argument("-p", action="store", dest="perplexity", default=100, type=int, help="accepted for compatibility")
    parser.add_argument("-q", "--infreq", action="store", dest="in_freq_file", help="accepted for compatibility")
    # without a matrix file: profile the assembly and keep the matrix on the device
    parser.add_argument("--assembly", action="store", dest="assembly",
                        help="instead of -i: compute the matrix from this multi-FASTA on the device (no matrix file at all)")
    parser.add_argument("--pattern", action="store", dest="pattern", default="1111", help="spaced-word pattern for --assembly")
    parser.add_argument("--strand", action="store", dest="strand", default="both", choices=["both", "plus", "minus"])
    parser.add_argument("-d", "--distance", action="store", dest="dist", default="JSD", choices=["Eucl", "JSD", "KT", "BC", "SC"])
    params = parser.parse_args(argv)
    if params.performtsne or params.interactive:
        print("Error, t-SNE projection and the interactive mode draw pictures: not part of the GPU front half "
              "(use knn_graph() to feed sklearn.manifold.TSNE(metric='precomputed'))", file=sys.stderr)
        sys.exit(1)
    if not params.distmat and not params.assembly:
        print("Error, give a matrix (-i) or an assembly (--assembly)", file=sys.stderr)
        sys.exit(1)
    return params


def main(argv=None):
    params = get_cmd(argv)
    if not os.path.isdir(params.outputdir):
        os.makedirs(params.outputdir)
    if params.assembly:
        print("Compute matrix")
        from . import phyloligo
        res, n, _ = phyloligo._profile_file(params.assembly, params.pattern, params.strand, ("freq32",))
        if res is None:
            print("Error, no FASTA record in {}".format(params.assembly), file=sys.stderr)
            sys.exit(1)
        matrix = engine.distance_matrix_device(res["freq32"], phyloligo._large_metric(params.dist), torch.float32, symmetric=True)
        if params.fastafile is None:
            params.fastafile = params.assembly
    else:
        print("Read matrix")
        matrix = read_matrix(params.distmat, params.large)
    print("Clusterize")
    labels_pred = clusterize(matrix, params.method, min_cluster_size=params.min_cluster_size, min_samples=params.min_samples,
                             nbk=params.nbk)
    pathout = os.path.join(params.outputdir, "data_cluster_indexes.dat")
    print("Store cluster indexes in {}".format(pathout))
    write_cluster_indexes(labels_pred, pathout)
    if params.fastafile:
        print("Write fasta per classes in {}/data_fasta_*.fa".format(params.outputdir))
        write_fastafile(labels_pred, params.fastafile, params.outputdir)
    return 0


if __name__ == "__main__":
    sys.exit(main())
