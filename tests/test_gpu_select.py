"""GPU parity of the phyloselect front half (phyloligo_b200/phyloselect.py, csrc/po_select.cu) against the
goldens of the reference's own KMedoids class, the numpy oracle and scikit-learn's neighbour graph."""
import os
import warnings

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import phylo_oracle as po
from oracle import select_oracle as so
from phyloligo_b200 import engine, io_formats, phyloselect, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "select_golden.npz")))


def test_kmedoids_matches_the_reference_goldens(golden):
    for name in golden["names"]:
        D = golden[name + "_D"]
        kw = dict(eval(str(golden[name + "_kwargs"])))
        with warnings.catch_warnings(record=True) as caught:
            warnings.simplefilter("always")
            km = phyloselect.KMedoids(distance_metric="precomputed", **kw).fit(D)
        assert np.array_equal(km.labels_, golden[name + "_labels"]), name
        assert km.n_iter_ == int(golden[name + "_n_iter"]), name
        assert np.array_equal(km.cluster_centers_, D[golden[name + "_medoids"]]), name
        assert np.array_equal(D[km.medoid_indices_], D[golden[name + "_medoids"]]), name
        assert sum("is empty" in str(w.message) for w in caught) == int(golden[name + "_empty_warnings"]), name


@pytest.mark.parametrize("n,k,dtype", [(1000, 4, np.float32), (2500, 7, np.float64), (333, 1, np.float32), (64, 64, np.float64)])
def test_kmedoids_matches_the_oracle_on_larger_matrices(n, k, dtype):
    rng = np.random.default_rng(n + k)
    centres = rng.normal(0, 4, size=(max(2, k), 6))
    P = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 1, size=(n, 6))
    D = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)).astype(dtype)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        labels, medoids, n_iter = so.kmedoids_fit(D.astype(np.float64), n_clusters=k)
        km = phyloselect.KMedoids(n_clusters=k).fit(D)
    assert km.n_iter_ == n_iter and np.array_equal(km.medoid_indices_, medoids) and np.array_equal(km.labels_, labels)
    # a device tensor is used in place, and the row sums are float64 sums of the stored entries
    Dd = torch.from_numpy(D).cuda()
    assert np.allclose(phyloselect.row_sums(Dd).cpu().numpy(), D.astype(np.float64).sum(axis=1), rtol=1e-12)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        assert np.array_equal(phyloselect.KMedoids(n_clusters=k).fit(Dd).labels_, labels)


def test_kmedoids_from_profiles_never_leaves_the_device():
    seqs = synth.make_sequences(400, 5000, seed=44)
    X = np.vstack([po.frequency_np(s, "1111", "both") for s in seqs]).astype(np.float32)
    km = phyloselect.KMedoids(n_clusters=2, distance_metric="JSD").fit(X)
    D = po.pairwise_np(X.astype(np.float64), "JSD")
    labels, medoids, _ = so.kmedoids_fit(D, n_clusters=2)
    assert np.array_equal(km.medoid_indices_, medoids) and np.array_equal(km.labels_, labels)
    assert np.array_equal(km.cluster_centers_, X[medoids])
    with pytest.raises(ValueError):
        phyloselect.KMedoids(n_clusters=2, distance_metric="cosine").fit(X)
    with pytest.raises(ValueError):
        phyloselect.KMedoids(n_clusters=0).fit(D)


@pytest.mark.parametrize("n,k", [(50, 1), (700, 15), (3000, 90), (1200, 300), (40, 39)])
def test_knn_graph_matches_sklearn(n, k):
    rng = np.random.default_rng(n * 7 + k)
    P = rng.random((n, 4))
    D = np.sqrt(((P[:, None, :] - P[None, :, :]) ** 2).sum(-1)).astype(np.float32)
    idx, dist = phyloselect.knn_graph(torch.from_numpy(D).cuda(), k)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    ref_idx, ref_dist = so.knn_graph(D.astype(np.float64), k)
    assert np.array_equal(dist, ref_dist.astype(np.float32))
    assert np.array_equal(dist, np.take_along_axis(D, idx.astype(np.int64), axis=1))
    assert (idx != np.arange(n)[:, None]).all() and all(len(set(r)) == k for r in idx.tolist())
    same = idx == ref_idx  # equal distances may be ordered differently by sklearn; ours: ties by column
    assert same.mean() > 0.99
    ties = ~same
    assert np.array_equal(dist[ties], ref_dist.astype(np.float32)[ties])
    # sparse form for TSNE(metric="precomputed"); block rows with the row's own column excluded
    G = phyloselect.knn_graph(D, k, as_sparse=True)
    assert G.shape == (n, n) and G.nnz == n * k
    r0 = n // 3
    bidx, bdist = phyloselect.knn_graph(torch.from_numpy(D[r0:r0 + 17]).cuda(), k, row0=r0)
    assert np.array_equal(bidx.cpu().numpy(), idx[r0:r0 + 17]) and np.array_equal(bdist.cpu().numpy(), dist[r0:r0 + 17])


def test_knn_ties_go_by_column_and_nan_sorts_last():
    D = np.ones((9, 9), dtype=np.float32)
    np.fill_diagonal(D, 0.0)
    D[0, 5] = np.nan
    D[0, 7] = 0.5
    idx, dist = phyloselect.knn_graph(D, 4)
    idx = idx.cpu().numpy()
    assert idx[0].tolist() == [7, 1, 2, 3] and idx[4].tolist() == [0, 1, 2, 3] and idx[8].tolist() == [0, 1, 2, 3]
    with pytest.raises(Exception):
        phyloselect.knn_graph(D, 9)


@pytest.mark.parametrize("kind", ["smooth", "quantised", "runs", "nan_rows"])
def test_knn_wide_rows_sampled_threshold_and_fallback(kind):
    """Rows longer than the 8192-entry sample: the one-pass threshold path, and the rows that must fall back
    to the exact five-pass select (thousands of ties at the threshold, small values hidden between the sampled
    runs, rows that are mostly NaN).  Expected: stable argsort of the row with its own column removed."""
    rng = np.random.default_rng(11)
    rows, n, k, r0 = 24, 30000, 301, 100
    D = rng.random((rows, n), dtype=np.float32)
    if kind == "quantised":
        D = np.floor(D * 4.0).astype(np.float32)  # 7500 ties per level
    elif kind == "runs":
        D += 1.0
        D[:, 200:200 + 2 * k] = rng.random((rows, 2 * k), dtype=np.float32)  # all k smallest in one stretch
        D[5, :] = 2.0
        D[5, 29000:29000 + k + 3] = np.arange(k + 3, dtype=np.float32)[::-1] / 1000.0
    elif kind == "nan_rows":
        D[::2, 400:] = np.nan
        D[3, :] = np.nan
    idx, dist = phyloselect.knn_graph(torch.from_numpy(D).cuda(), k, row0=r0)
    idx, dist = idx.cpu().numpy(), dist.cpu().numpy()
    for i in range(rows):
        row = D[i].astype(np.float64)
        key = np.where(np.isnan(row), np.inf, row)
        order = np.lexsort((np.arange(n), np.isnan(row), key))
        order = order[order != r0 + i][:k]
        assert np.array_equal(idx[i], order), (kind, i)
        assert np.array_equal(dist[i], D[i][order], equal_nan=True), (kind, i)


def test_command_line_front_half(tmp_path, capsys):
    seqs = synth.make_sequences(150, 3000, seed=9)
    fasta = os.path.join(tmp_path, "asm.fa")
    synth.write_fasta(fasta, seqs, line=70)
    X = np.vstack([po.frequency_np(s, "1111", "both") for s in seqs])
    D = po.pairwise_np(X, "JSD")
    want, _, _ = so.kmedoids_fit(D, n_clusters=3)
    outs = {}
    np.savetxt(os.path.join(tmp_path, "m.txt"), D, delimiter="\t")
    D.astype(np.float32).tofile(os.path.join(tmp_path, "m.bin"))
    io_formats.write_hdf5(os.path.join(tmp_path, "m.h5"), "distances", D.astype(np.float32))
    for tag, args in (("txt", ["-i", os.path.join(tmp_path, "m.txt")]),
                      ("memmap", ["-i", os.path.join(tmp_path, "m.bin"), "--large", "memmap"]),
                      ("h5py", ["-i", os.path.join(tmp_path, "m.h5"), "--large", "h5py"]),
                      ("device", ["--assembly", fasta, "-d", "JSD", "--pattern", "1111"])):
        out = os.path.join(tmp_path, "out_" + tag)
        assert phyloselect.main(args + ["-m", "kmedoids", "-k", "3", "-o", out, "-f", fasta]) == 0
        rows = [tuple(int(v) for v in line.split()) for line in open(os.path.join(out, "data_cluster_indexes.dat"))]
        labels = np.empty(len(seqs), dtype=np.int64)
        for cl, idx in rows:
            labels[idx] = cl
        outs[tag] = labels
        assert rows == sorted(rows)  # grouped by class, indices ascending (reference :742-751)
        # one FASTA per class, records in file order, sequences intact
        for cl in np.unique(labels):
            recs = open(os.path.join(out, "data_fasta_cl%d.fa" % cl)).read().split(">")[1:]
            members = np.where(labels == cl)[0]
            assert len(recs) == len(members)
            for rec, idx in zip(recs, members):
                lines = rec.split("\n")
                assert lines[0] == "c%d" % idx and "".join(lines[1:]) == seqs[idx].decode()
                assert max(len(l) for l in lines[1:]) <= 60
    assert np.array_equal(outs["txt"], want)
    # float32 files: the same partition unless a cost comparison is within float32 rounding (not here)
    for tag in ("memmap", "h5py", "device"):
        assert np.array_equal(outs[tag], want), tag
    assert "Clusterize" in capsys.readouterr().out
    with pytest.raises(SystemExit):
        phyloselect.main(["-i", "x", "-m", "kmedoids", "-o", str(tmp_path), "-t"])


def test_command_line_against_the_unmodified_reference_scripts(tmp_path):
    """The matrices written by the reference's unmodified bin/phyloligo.py, clustered by its unmodified
    bin/phyloselect.py (tests/golden/make_select_cli_golden.py): this command line on the same matrix files gives
    the same data_cluster_indexes.dat and the same per-cluster FASTA files, byte for byte."""
    import sys
    sys.path.insert(0, GOLDEN)
    import make_cli_golden as mk
    golden = dict(np.load(os.path.join(GOLDEN, "select_cli_golden.npz")))
    fasta, n = mk.assembly()
    path = os.path.join(tmp_path, "asm.fasta")
    open(path, "wb").write(fasta)
    assert len(golden["names"]) == 4
    for name in golden["names"]:
        mat = os.path.join(tmp_path, name + ".mat")
        open(mat, "wb").write(golden[name + "_matrix_bytes"].tobytes())
        out = os.path.join(tmp_path, name + "_sel")
        assert phyloselect.main(["-i", mat, "-o", out, "-f", path] + str(golden[name + "_select_args"]).split()) == 0
        assert open(os.path.join(out, "data_cluster_indexes.dat"), "rb").read() == golden[name + "_indexes"].tobytes(), name
        files = sorted(f for f in os.listdir(out) if f.startswith("data_fasta_"))
        assert files == [str(f) for f in golden[name + "_fasta_names"]], name
        for f in files:
            assert open(os.path.join(out, f), "rb").read() == golden[name + "_" + f].tobytes(), (name, f)
