"""GPU parity: po_prepare_profiles + po_distance_block (through the C ABI) against
the reference golden matrices and the oracle.

Tolerances (BASELINE.json north_star): Eucl / BC / JSD <= 1e-6 relative against the
float64-accumulated oracle on the same inputs (the tensor-core Eucl path, used from 256
dimensions up, has a stated tolerance of 1e-4; PO_EUCL_EXACT=1 selects the exact kernel);
KT and SC exact up to float64 rounding of the final division (1e-12)."""
import numpy as np
import pytest
import torch

from oracle import phylo_oracle as po
from phyloligo_b200 import engine, phylodist, synth

pytestmark = pytest.mark.gpu

RTOL = 1e-6
RTOL_TC = 1e-4  # stated tolerance of the tensor-core Euclidean path


@pytest.fixture(params=["tensor", "exact"])
def eucl_path(request):
    """Run a test once per Euclidean kernel: metric EuclGram (tensor cores from 256 dimensions up,
    the form of the reference's --large workers) and metric Eucl (the exact direct sum)."""
    return request.param


def _eucl_name(metric, eucl_path):
    return "EuclGram" if (metric == "Eucl" and eucl_path == "tensor") else metric


def _gpu_matrix(X, metric, out_dtype=torch.float64, symmetric=True):
    Xd = torch.from_numpy(np.ascontiguousarray(X)).cuda()
    return engine.distance_matrix_device(Xd, metric, out_dtype, symmetric).cpu().numpy()


def _profiles(n, mean_len, pattern, seed, dtype=np.float64):
    seqs = synth.make_sequences(n, mean_len, seed=seed)
    return np.vstack([po.frequency_np(s, pattern, "both", dtype) for s in seqs])


def _assert_close(got, want, rtol=RTOL, atol=0.0):
    err = np.abs(got - want)
    tol = rtol * np.abs(want) + atol
    bad = err > tol
    assert not bad.any(), "max rel err %.3e at %s" % (
        (err / np.maximum(np.abs(want), 1e-300)).max(), np.argwhere(bad)[:5].tolist())


def test_golden_matrices(distance_golden, eucl_path):
    for name in ("real_k4", "sparse_64", "onehot_16"):
        X = distance_golden[name + "_X"]
        # oracle on the float32-rounded inputs the kernel consumes, float64 accumulation
        X32 = X.astype(np.float32).astype(np.float64)
        for metric in ("Eucl", "JSD"):
            got = _gpu_matrix(X, _eucl_name(metric, eucl_path))
            tc = metric == "Eucl" and eucl_path == "tensor" and X.shape[1] >= 256
            _assert_close(got, po.pairwise_np(X32, metric), rtol=RTOL_TC if tc else RTOL, atol=1e-12)
            # and against the reference's own float64 output on the float64 inputs
            ref = distance_golden[name + "_" + metric]
            assert np.allclose(got, ref, rtol=RTOL_TC if tc else 5e-6, atol=1e-9)
            assert np.array_equal(got, got.T)
            assert (np.diag(got) == 0).all()
    e = np.eye(16)[:6]
    got = _gpu_matrix(e, "JSD")
    off = got[~np.eye(6, dtype=bool)]
    assert np.allclose(off, np.log(2), rtol=1e-6)


@pytest.mark.parametrize("metric", ["Eucl", "JSD", "BC"])
@pytest.mark.parametrize("pattern,n,length", [("1111", 150, 4000), ("11111", 70, 2000), ("111010011", 66, 3000)])
def test_float_metrics_vs_fp64_oracle(metric, pattern, n, length, eucl_path):
    if metric != "Eucl" and eucl_path == "exact":
        pytest.skip("only Eucl has two kernels")
    X = _profiles(n, length, pattern, seed=31, dtype=np.float32)
    X[3] = 0.0  # an empty contig: all-zero profile
    want = po.pairwise_np(X.astype(np.float64), metric)
    tol = RTOL_TC if (metric == "Eucl" and eucl_path == "tensor" and X.shape[1] >= 256) else RTOL
    for out_dtype, rt in ((torch.float64, tol), (torch.float32, tol)):
        got = _gpu_matrix(X, _eucl_name(metric, eucl_path), out_dtype).astype(np.float64)
        mask = np.isfinite(want)
        assert np.array_equal(np.isnan(got), np.isnan(want))  # BC of two zero rows: 0/0
        _assert_close(got[mask], want[mask], rtol=rt, atol=1e-12)
    full = _gpu_matrix(X, _eucl_name(metric, eucl_path), torch.float64, symmetric=False)
    sym = _gpu_matrix(X, _eucl_name(metric, eucl_path), torch.float64, symmetric=True)
    assert np.array_equal(full, sym, equal_nan=True)  # mirrored tiles are bitwise the computed ones


def test_jsd_edge_values():
    z = np.zeros(256)
    a = np.full(256, 1.0 / 256)
    X = np.vstack([z, a, a, np.eye(256)[0], np.eye(256)[1]])
    got = _gpu_matrix(X, "JSD")
    assert got[0, 0] == 0.0 and got[1, 2] == 0.0
    assert got[0, 1] == pytest.approx(0.5 * np.log(2), rel=1e-6)   # against an all-zero row
    assert got[3, 4] == pytest.approx(np.log(2), rel=1e-6)
    assert (got >= 0).all() and (got <= np.log(2) * (1 + 1e-6)).all()


def test_jsd_near_identical_profiles_keep_relative_accuracy():
    rng = np.random.default_rng(3)
    base = rng.random(1024)
    base /= base.sum()
    rows = [base]
    for eps in (1e-1, 1e-2, 1e-3):
        r = base * (1 + eps * rng.standard_normal(1024))
        rows.append(np.abs(r) / np.abs(r).sum())
    X = np.vstack(rows).astype(np.float32)
    got = _gpu_matrix(X, "JSD")
    want = po.pairwise_np(X.astype(np.float64), "JSD")
    off = ~np.eye(len(rows), dtype=bool)
    assert (np.abs(got[off] / want[off] - 1) < RTOL).all()


@pytest.mark.parametrize("metric", ["KT", "SC"])
def test_rank_metrics_exact(metric):
    # tie-heavy short-read profiles (config C4 shape) plus constant / empty rows
    seqs = synth.make_sequences(60, 375, seed=4, model="short")
    X = np.vstack([po.frequency_np(s, "1111", "both") for s in seqs])
    X[7] = 0.0
    X[9] = 1.0 / 256
    got = _gpu_matrix(X, metric)
    fn = po.KT if metric == "KT" else po.SC
    n = X.shape[0]
    want = np.array([[fn(X[i], X[j]) for j in range(n)] for i in range(n)])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert np.abs(got[m] - want[m]).max() < 1e-12
    if metric == "KT":
        assert (got[7] == 0).all() and (got[9] == 0).all()
        assert np.array_equal(got, want)  # integer counts, same float64 expression: bit exact


def test_spearman_k7_does_not_overflow():
    """SC above 4096 dimensions runs on the CUDA cores; at dim = 16384 (k = 7) the rank products of
    correlated rows exceed 2^31 within one 32-element chunk (and a single product does above
    dim = 46341): the accumulation is 64-bit."""
    dim = 4 ** 7
    rng = np.random.default_rng(77)
    X = rng.random((9, dim))
    X[0] = np.arange(dim)            # extreme centred ranks +/-(dim-1), perfectly correlated with row 1
    X[1] = np.arange(dim) * 2.0 + 5
    X[2] = -np.arange(dim)           # and perfectly anti-correlated
    X[3] = np.round(X[3] * 6)        # tie-heavy
    X[4] = 1.0                       # constant row -> NaN
    got = _gpu_matrix(X, "SC")
    want = np.array([[po.SC(a, b) for b in X] for a in X])
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    assert np.abs(got[m] - want[m]).max() < 1e-12
    assert got[0, 1] == 0.0 and got[0, 2] == 2.0


def test_jsd_sparse_k5_short_contigs_accuracy():
    """C5-like profiles (k = 5, 5 kb contigs: about ten windows per bin, many exact zeros, so nearly
    every term takes the log-based branch): 2 000 rows x 1024 dimensions against the float64 C oracle."""
    from oracle import coracle
    n = 2000
    seqs = synth.make_sequences(n, 5000, seed=55)
    text, begin, end = engine.sequences_to_text(seqs)
    X = coracle.profile_batch(text, begin, end, "11111", "both").astype(np.float32)
    X[17] = 0.0
    assert (X == 0).mean() > 0.001
    got = _gpu_matrix(X, "JSD", torch.float64)
    want = coracle.pairwise_rows("JSD", X.astype(np.float64))
    off = ~np.eye(n, dtype=bool)
    rel = np.abs(got - want)[off] / want[off]
    print("sparse k=5 JSD: max rel err %.3e, mean %.3e" % (rel.max(), rel.mean()))
    assert rel.max() < RTOL
    assert (np.diag(got) == 0).all() and np.array_equal(got, got.T)
    got32 = _gpu_matrix(X, "JSD", torch.float32).astype(np.float64)
    assert (np.abs(got32 - want)[off] / want[off]).max() < RTOL


def test_rank_metrics_small_dims():
    rng = np.random.default_rng(2)
    for dim in (2, 3, 5, 16, 33, 64):
        X = rng.integers(0, 3, size=(9, dim)).astype(np.float64)
        for metric, fn in (("KT", po.KT), ("SC", po.SC)):
            got = _gpu_matrix(X, metric)
            want = np.array([[fn(a, b) for b in X] for a in X])
            assert np.array_equal(np.isnan(got), np.isnan(want)), (metric, dim)
            m = ~np.isnan(want)
            assert np.abs(got[m] - want[m]).max() < 1e-12, (metric, dim)


def test_pair_api_and_block_rows():
    X = _profiles(130, 2500, "1111", seed=12)
    assert phylodist.Eucl(X[0], X[1]) == pytest.approx(po.Eucl(X[0], X[1]), rel=RTOL)
    assert phylodist.JSD(X[0], X[1]) == pytest.approx(po.JSD(X[0], X[1]), rel=RTOL)
    assert phylodist.BC(X[0], X[1]) == pytest.approx(po.BC(X[0], X[1]), rel=RTOL)
    assert phylodist.KT(X[0], X[1]) == pytest.approx(po.KT(X[0], X[1]), abs=1e-12)
    assert phylodist.SC(X[0], X[1]) == pytest.approx(po.SC(X[0], X[1]), abs=1e-12)
    assert phylodist.KL(X[0], X[1]) == pytest.approx(po.KL(X[0], X[1]), rel=1e-12)
    z = X[2].copy()
    z[::3] = 0.0  # zeros on one side: those terms are dropped, as posdef_check_value does
    assert phylodist.KL(z, X[1]) == pytest.approx(po.KL(z, X[1]), rel=1e-12)
    assert phylodist.KL(X[1], z) == pytest.approx(po.KL(X[1], z), rel=1e-12)
    # 2-D x 2-D JSD: rows index the second argument (core/phylodist.py:58-66)
    X32 = X.astype(np.float32)
    got = phylodist.JSD(X32, X32[10:31])
    want = po.JSD(X32.astype(np.float64), X32[10:31].astype(np.float64))
    assert got.shape == (21, 130) and got.dtype == np.float32
    assert np.allclose(got, want, rtol=RTOL, atol=1e-9)
    # block rows written at an offset, ragged edges (130 is not a multiple of 64)
    Xd = torch.from_numpy(X32).cuda()
    P, aux, dim = engine.prepare(Xd, "Eucl")
    full = engine.distance_matrix_device(Xd, "Eucl", torch.float32, symmetric=False)
    blk = torch.full((40, 130), -1.0, dtype=torch.float32, device="cuda")
    engine.distance_block("Eucl", P, aux, dim, 70, 110, 0, 130, blk, 70, 0)
    assert torch.equal(blk, full[70:110])


def test_panel_streamer_matches_resident_matrix():
    X = _profiles(300, 1500, "1111", seed=13, dtype=np.float32)
    Xd = torch.from_numpy(X).cuda()
    want = engine.distance_matrix_device(Xd, "JSD", torch.float32, symmetric=True).cpu().numpy()
    for symmetric in (True, False):
        got = np.zeros((300, 300), dtype=np.float32)
        st = engine.PanelStreamer(Xd, "JSD", torch.float32, panel_rows=128, symmetric=symmetric)

        def sink(r0, r1, host):
            got[r0:r1] = host

        st.run(sink)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("n,dim", [(300, 256), (129, 1024), (200, 4096), (128, 320)])
def test_tensor_core_euclidean(n, dim):
    """tcgen05 Gram path: tolerance 1e-4 (stated), symmetric, exact zero diagonal, identical
    whichever block computes an entry; near-duplicate rows keep their small distances."""
    rng = np.random.default_rng(dim + n)
    X = rng.dirichlet(np.full(dim, 0.3), size=n).astype(np.float32)
    X[5] = X[4] * (1 + 1e-3 * rng.standard_normal(dim)).astype(np.float32)  # near-duplicate pair
    X[7] = 0.0
    want = po.pairwise_np(X.astype(np.float64), "Eucl")
    got = _gpu_matrix(X, "EuclGram", torch.float64, symmetric=True)
    off = ~np.eye(n, dtype=bool)
    rel = np.abs(got - want)[off] / want[off]
    print("tensor-core Eucl n=%d dim=%d: max rel err %.3e (near-duplicate pair %.3e)" % (
        n, dim, rel.max(), abs(got[4, 5] - want[4, 5]) / want[4, 5]))
    assert rel.max() < RTOL_TC
    assert (np.diag(got) == 0).all() and np.array_equal(got, got.T)
    full = _gpu_matrix(X, "EuclGram", torch.float64, symmetric=False)
    assert np.array_equal(full, got)
    # a block row through the worker API and float32 output
    Xd = torch.from_numpy(X).cuda()
    P, aux, d = engine.prepare(Xd, "EuclGram")
    blk = torch.empty((40, n), dtype=torch.float32, device="cuda")
    engine.distance_block("EuclGram", P, aux, d, 70, 110, 0, n, blk, 70, 0)
    # the float32 epilogue takes a float32 square root of the same float64 d^2
    assert np.allclose(blk.cpu().numpy(), got[70:110], rtol=2e-7, atol=0)


@pytest.mark.parametrize("n,dim", [(130, 64), (300, 256), (129, 1024), (140, 4096), (70, 200)])
def test_tensor_core_spearman_is_exact(n, dim, monkeypatch):
    """SC on tcgen05 (integer float16 digit planes, exact float32 accumulation): bit for bit the
    CUDA-core integer kernel, 1e-12 from scipy.stats.spearmanr, NaN for constant rows."""
    rng = np.random.default_rng(dim * 7 + n)
    X = rng.integers(0, 12, size=(n, dim)).astype(np.float64)  # tie-heavy
    X[: n // 2] = rng.random((n // 2, dim))                     # and tie-free rows (largest rank sums)
    X[3] = 0.0                                                  # constant row -> NaN
    X[5] = np.arange(dim)                                       # extreme ranks: +/-(dim-1)
    X[6] = -np.arange(dim)
    X /= np.maximum(X.sum(axis=1, keepdims=True), 1e-300)
    for dtype in (torch.float64, torch.float32):
        monkeypatch.setenv("PO_SC_CUDA_CORES", "1")
        ref = _gpu_matrix(X, "SC", dtype, symmetric=True)
        monkeypatch.delenv("PO_SC_CUDA_CORES")
        got = _gpu_matrix(X, "SC", dtype, symmetric=True)
        full = _gpu_matrix(X, "SC", dtype, symmetric=False)
        assert np.array_equal(got, ref, equal_nan=True)
        assert np.array_equal(full, ref, equal_nan=True)
    want = np.array([[po.SC(a, b) for b in X[:24]] for a in X[:24]])
    sub = _gpu_matrix(X, "SC", torch.float64, symmetric=True)[:24, :24]
    assert np.array_equal(np.isnan(sub), np.isnan(want))
    m = ~np.isnan(want)
    assert np.abs(sub[m] - want[m]).max() < 1e-12
    assert sub[5, 6] == 2.0 and sub[5, 5] == 0.0


@pytest.mark.parametrize("metric,n", [("JSD", 700), ("EuclGram", 520), ("SC", 300)])
@pytest.mark.parametrize("share", [0.0, 0.4, 1.0])
def test_matrix_to_host_ships_finished_blocks(metric, n, share):
    """engine.matrix_to_host: panels' right parts by strided DMA (po_copy2d_async), the mirrored column
    blocks partly by DMA and partly built on the host from what has arrived (po_host_mirror_*, released in
    stream order): exactly the resident matrix in pinned host memory, whatever the split."""
    rng = np.random.default_rng(n)
    X = torch.from_numpy(rng.dirichlet(np.ones(256), size=n).astype(np.float32)).cuda()
    want = engine.distance_matrix_device(X, metric, torch.float32, symmetric=True).cpu()
    host = torch.full((n, n), -1.0, dtype=torch.float32).pin_memory()
    stats = {}
    copied = engine.matrix_to_host(X, metric, host, torch.float32, panel_rows=128, host_mirror=share, mirror_threads=3,
                                   stats=stats)
    torch.cuda.synchronize()
    assert copied == stats["dma_bytes"] and copied + stats["host_mirrored_bytes"] == n * n * 4
    if share == 0.0:
        assert stats["host_mirrored_bytes"] == 0
    if share == 1.0:
        tri = sum((min(n, r0 + 128) - r0) * (n - r0) for r0 in range(0, n, 128)) * 4
        assert copied == tri
    assert torch.equal(host, want)
    with pytest.raises(Exception):
        engine.matrix_to_host(X, metric, torch.empty((n, n), dtype=torch.float32), torch.float32)  # not pinned


def test_matrix_to_host_float64_goes_by_dma_only():
    X = torch.from_numpy(np.random.default_rng(3).dirichlet(np.ones(64), size=300).astype(np.float32)).cuda()
    want = engine.distance_matrix_device(X, "JSD", torch.float64, symmetric=True).cpu()
    host = torch.zeros((300, 300), dtype=torch.float64).pin_memory()
    stats = {}
    engine.matrix_to_host(X, "JSD", host, torch.float64, panel_rows=128, host_mirror=1.0, stats=stats)
    torch.cuda.synchronize()
    assert stats["host_mirrored_bytes"] == 0 and torch.equal(host, want)


def test_host_mirror_repeated_steps_reuse_the_pool():
    """Back-to-back steps into the same host matrix (what bench.py's e2e loop does): every step's result is
    complete when the call returns and the device is synchronised."""
    n = 900
    host = torch.zeros((n, n), dtype=torch.float32).pin_memory()
    for seed in range(4):
        X = torch.from_numpy(np.random.default_rng(seed).dirichlet(np.ones(256), size=n).astype(np.float32)).cuda()
        engine.matrix_to_host(X, "JSD", host, torch.float32, panel_rows=256, host_mirror=1.0)
        torch.cuda.synchronize()
        want = engine.distance_matrix_device(X, "JSD", torch.float32, symmetric=True).cpu()
        assert torch.equal(host, want), seed


@pytest.mark.parametrize("metric,n", [("JSD", 1100), ("EuclGram", 900)])
def test_block_rows_row_panels_and_mirrored_host_sink_one_rank(metric, n):
    """multigpu.BlockRows on one rank (it owns both block rows): launched whole and in row panels the rows equal
    the symmetric matrix bit for bit; with MirroredHostSink only the part on and right of the diagonal goes to the
    host by DMA and the rest is mirrored there (what bench.py's multi-GPU end-to-end leg does on every rank)."""
    from phyloligo_b200 import multigpu

    rng = np.random.default_rng(n)
    X = torch.from_numpy(rng.dirichlet(np.ones(256), size=n).astype(np.float32)).cuda()
    want = engine.distance_matrix_device(X, metric, torch.float32, symmetric=True)
    P, aux, dim = engine.prepare(X, metric)
    job = multigpu.BlockRows(n, torch.float32, 0, 1)
    assert job.rows_owned == n
    for panel in (None, 256):
        job.matrix.fill_(float("nan"))
        job.compute(metric, P, aux, dim, panel_rows=panel)
        torch.cuda.synchronize()
        for i in job.my_ranges:
            a, b = job.ranges[i]
            assert torch.equal(job.out_rows[i], want[a:b]), (panel, i)
    host = torch.full((n, n), -1.0, dtype=torch.float32).pin_memory()
    pool = engine.HostMirror(3)
    try:
        sink = multigpu.MirroredHostSink(host, pool)
        for _ in range(2):  # a sink serves many steps
            sink.reset()
            job.compute(metric, P, aux, dim, ship=sink.ship, left_parts=False, panel_rows=256)
            sink.finish()
            torch.cuda.synchronize()
            assert torch.equal(host, want.cpu())
            assert sink.dma_bytes + sink.mirrored_bytes == n * n * 4 and sink.mirrored_bytes > 0
    finally:
        pool.close()
        job.close()


def test_copy2d_strided_views():
    a = torch.arange(40 * 50, dtype=torch.float32, device="cuda").reshape(40, 50)
    h = torch.zeros((40, 50), dtype=torch.float32).pin_memory()
    engine.copy2d(h[5:30, 7:41], a[5:30, 7:41])
    torch.cuda.synchronize()
    want = torch.zeros((40, 50))
    want[5:30, 7:41] = a[5:30, 7:41].cpu()
    assert torch.equal(h, want)
    engine.copy2d(h[0:0, :], a[0:0, :])  # empty block: nothing to do


@pytest.mark.parametrize("metric,n,dim", [("JSD", 100_000, 256), ("EuclGram", 50_000, 4096), ("SC", 20_000, 256)])
def test_full_size_matrix_properties(metric, n, dim):
    """BASELINE.json sizes (C2 JSD 100k x 256, C3 Eucl 50k x 4096, C4 SC 20k x 256) through
    size-independent properties: exact diagonal, bitwise symmetry, value range, sampled entries against
    the float64 oracle, and sampled row panels recomputed without the symmetry shortcut."""
    need = n * n * 4 + 6 * n * dim * 4 + (2 << 30)
    free, _ = torch.cuda.mem_get_info()
    if free < need:
        pytest.skip("needs %.0f GB of free HBM" % (need / 1e9))
    g = torch.Generator(device="cuda").manual_seed(n + dim)
    X = torch.rand((n, dim), device="cuda", generator=g) ** 3
    X[::97, ::5] = 0.0                      # exact zeros, as sparse profiles have
    if metric == "SC":
        X = torch.round(X * 40) / 40        # tie-heavy
    X /= X.sum(dim=1, keepdim=True)
    X[12345 % n] = 0.0                      # an empty record's all-zero row
    M = engine.distance_matrix_device(X, metric, torch.float32, symmetric=True)
    diag = torch.diagonal(M)
    zero_row = 12345 % n
    if metric == "SC":
        assert torch.isnan(M[zero_row]).all() and torch.isnan(M[:, zero_row]).all()
        keep = torch.ones(n, dtype=torch.bool, device="cuda")
        keep[zero_row] = False
        assert (diag[keep] == 0).all()
    else:
        assert (diag == 0).all()
    for r0 in range(0, n, 8192):            # bitwise symmetry, panel by panel
        r1 = min(n, r0 + 8192)
        a, b = M[r0:r1], M[:, r0:r1].T
        assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0))
    finite = M[~torch.isnan(M)] if metric == "SC" else M
    lo, hi = float(finite.min()), float(finite.max())
    if metric == "JSD":
        assert lo == 0.0 and hi <= np.log(2) * (1 + 1e-6)
        others = torch.ones(n, dtype=torch.bool, device="cuda")
        others[zero_row] = False
        assert torch.allclose(M[zero_row][others], torch.full((n - 1,), np.log(2) / 2, device="cuda"), rtol=RTOL)
    elif metric == "SC":
        assert lo > -1e-6 and hi < 2.0 + 1e-6
    else:
        assert lo == 0.0
    # sampled entries against the oracle in float64
    rng = np.random.default_rng(5)
    ii, jj = rng.integers(0, n, 600), rng.integers(0, n, 600)
    Xh = X[np.unique(np.concatenate([ii, jj]))].double().cpu().numpy()
    index = {v: k for k, v in enumerate(np.unique(np.concatenate([ii, jj])))}
    got = M[torch.from_numpy(ii).cuda(), torch.from_numpy(jj).cuda()].cpu().numpy().astype(np.float64)
    fn = {"JSD": po.JSD, "EuclGram": po.Eucl, "SC": po.SC}[metric]
    want = np.array([fn(Xh[index[i]], Xh[index[j]]) for i, j in zip(ii, jj)])
    ok = ~np.isnan(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    tol = {"JSD": RTOL, "EuclGram": RTOL_TC, "SC": RTOL}[metric]  # float32 output
    assert np.allclose(got[ok], want[ok], rtol=tol, atol=1e-7), np.abs(got[ok] - want[ok]).max()
    # row panels recomputed as plain block rows (every entry computed, no mirror)
    P, aux, d = engine.prepare(X, metric)
    for r0 in (0, (n // 2) // 128 * 128, n - 300):
        blk = torch.empty((300, n), dtype=torch.float32, device="cuda")
        engine.distance_block(metric, P, aux, d, r0, r0 + 300, 0, n, blk, r0, 0)
        assert torch.equal(torch.nan_to_num(blk, nan=-7.0), torch.nan_to_num(M[r0:r0 + 300], nan=-7.0))


@pytest.mark.parametrize("dim", [1, 2, 7, 256, 1000, 4096])
def test_rank_transform_matches_scipy_rankdata(dim):
    import scipy.stats as sst
    rng = np.random.default_rng(dim)
    X = rng.integers(0, 9, size=(37, dim)).astype(np.float64)
    X[:10] = rng.random((10, dim))
    X[11] = 3.0
    for dtype in (np.float64, np.float32):
        got = engine.rank_transform(torch.from_numpy(X.astype(dtype)).cuda()).cpu().numpy()
        want = np.vstack([sst.rankdata(r.astype(dtype), method="average") for r in X])
        assert np.array_equal(got, want)
