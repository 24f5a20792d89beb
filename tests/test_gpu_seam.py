"""The worker-level seam of the reference (bin/phyloligo.py:166-171, 195-301, 693-813),
function by function through the CUDA library; the host sink on the device side; and the ctypes
stub of INTEGRATION.md executed as written."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import phylo_oracle as po
from phyloligo_b200 import _lib, engine, hostsink, io_formats, phyloligo, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def seqs():
    return [s.decode() for s in synth.make_sequences(150, 1800, seed=21)] + ["", "NNNNACGTNNN"]


@pytest.fixture(scope="module")
def profiles(seqs):
    return np.vstack([po.frequency_np(s, "1111", "both") for s in seqs])


def _oracle_matrix(X, metric):
    X = np.asarray(X, dtype=np.float64)
    if metric in ("Eucl", "JSD", "BC"):
        return po.pairwise_np(X, metric)
    fn = po.KT if metric == "KT" else po.SC
    return np.array([[fn(a, b) for b in X] for a in X])


def test_compute_frequency_workers(seqs, tmp_path):
    for pattern, strand in (("1111", "both"), ("110101", "minus"), ("111", "plus")):
        for s in seqs[:6] + seqs[-2:]:
            want = po.frequency_np(s, pattern, strand)
            assert np.array_equal(phyloligo.compute_frequency(s, pattern, strand), want)
            assert np.array_equal(phyloligo.frequency_pack((s, pattern, strand)), want)
    # compute_frequency_memmap: row i of a caller-owned float32 memmap (reference :693-720)
    mm = np.memmap(os.path.join(tmp_path, "freq"), dtype=np.float32, mode="w+", shape=(len(seqs), 256))
    for i, s in enumerate(seqs[:9]):
        phyloligo.compute_frequency_memmap(mm, i, s, "1111", "both")
    want = np.vstack([po.frequency_np(s, "1111", "both") for s in seqs[:9]]).astype(np.float32)
    assert np.array_equal(np.asarray(mm[:9]), want) and not np.asarray(mm[9:]).any()
    # compute_frequency_h5py_chunk: file frequencies_{start}_{stop}, dataset 'frequencies' (reference :756-792)
    phyloligo.compute_frequency_h5py_chunk(str(tmp_path), seqs[10:31], "1111", "both", 10, 31)
    got = io_formats.read_hdf5(os.path.join(tmp_path, "frequencies_10_31"), "frequencies")
    assert got.dtype == np.float32
    assert np.array_equal(got, np.vstack([po.frequency_np(s, "1111", "both") for s in seqs[10:31]]).astype(np.float32))


@pytest.mark.parametrize("metric", ["Eucl", "JSD", "KT", "BC", "SC"])
def test_compute_unpack(profiles, metric):
    want = _oracle_matrix(profiles[[3, 40]], metric)[0, 1]
    i, j, d = phyloligo.compute_unpack((3, 40, profiles[3], profiles[40], metric))
    assert (i, j) == (3, 40)
    assert d == pytest.approx(want, rel=1e-6 if metric in ("Eucl", "JSD", "BC") else 1e-12, abs=1e-12)


@pytest.mark.parametrize("metric", ["Eucl", "JSD", "KT", "BC", "SC"])
def test_distances_loc_ragged_slices(profiles, metric):
    """output[s] = D(X[s], X) for the slices gen_even_slices hands the workers (reference :195-222, 424)."""
    from sklearn.utils import gen_even_slices
    X = profiles.astype(np.float32)
    n = X.shape[0]
    out = np.full((n, n), -5.0, dtype=np.float32)
    for s in gen_even_slices(n, 7):
        phyloligo.distances_loc(out, X, s, metric)
    want = _oracle_matrix(X, metric)
    assert np.array_equal(np.isnan(out), np.isnan(want))
    m = ~np.isnan(want)
    tol = 1e-4 if metric == "Eucl" else 1e-6  # the --large workers' Eucl is the Gram form
    assert np.allclose(out[m], want[m], rtol=tol, atol=1e-7)
    # a slice with open ends, float64 destination
    out64 = np.zeros((n, n))
    phyloligo.distances_loc(out64, X, slice(None, 40), metric)
    assert np.allclose(out64[:40][m[:40]], want[:40][m[:40]], rtol=tol, atol=1e-7) and not out64[40:].any()


@pytest.mark.parametrize("metric", ["Eucl", "JSD", "SC"])
def test_distances_h5py_worker(profiles, metric, tmp_path):
    """distance_{start}_{stop} files from an HDF5 frequency file (reference :233-301)."""
    X = profiles.astype(np.float32)
    n = X.shape[0]
    fpath = os.path.join(tmp_path, "frequencies_results")
    io_formats.write_hdf5(fpath, "frequencies", X)
    phyloligo.distances_h5py(str(tmp_path), fpath, slice(30, 77), metric)
    got = io_formats.read_hdf5(os.path.join(tmp_path, "distance_30_77"), "distances")
    assert got.shape == (47, n) and got.dtype == np.float32
    want = _oracle_matrix(X, metric)[30:77]
    m = ~np.isnan(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(got[m], want[m], rtol=1e-4 if metric == "Eucl" else 1e-6, atol=1e-7)


def test_integration_md_stub_runs_as_written(seqs, profiles):
    """The ctypes stub INTEGRATION.md tells a maintainer of the reference to paste: executed verbatim
    (only the library path is made absolute), then checked against the oracle."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "def compute_frequency" in b and "def distances_loc" in b)
    assert 'C.CDLL("libphyloligo_b200.so")' in stub
    ns = {}
    exec(compile(stub.replace('C.CDLL("libphyloligo_b200.so")', "C.CDLL(%r)" % _lib.LIB_PATH), "INTEGRATION.md", "exec"), ns)
    for s in seqs[:5]:
        assert np.array_equal(ns["compute_frequency"](s, "1111", "both"), po.frequency_np(s, "1111", "both"))
    assert np.array_equal(ns["compute_frequency"](seqs[2], "10101", "plus"), po.frequency_np(seqs[2], "10101", "plus"))
    X = profiles.astype(np.float32)
    n = X.shape[0]
    for metric in ("Eucl", "JSD", "KT", "BC", "SC"):
        out = np.zeros((n, n), dtype=np.float32)
        ns["distances_loc"](out, X, slice(20, 90), metric)
        want = _oracle_matrix(X, metric)[20:90]
        m = ~np.isnan(want)
        assert np.array_equal(np.isnan(out[20:90]), np.isnan(want))
        assert np.allclose(out[20:90][m], want[m], rtol=1e-4 if metric == "Eucl" else 1e-6, atol=1e-7)
        assert not out[:20].any() and not out[90:].any()


def test_row_shipper_into_file_mapping(tmp_path):
    """hostsink.RowShipper: strided device blocks -> pinned ring -> the mapping of a file, slots far
    smaller than the blocks, several blocks in flight."""
    n = 700
    src = torch.rand((n, n), dtype=torch.float32, device="cuda")
    path = os.path.join(tmp_path, "m.bin")
    with hostsink.FileMatrix(path, n, n, np.float32, create=True) as fm:
        fm.warm([(0, n)], threads=2)
        sh = hostsink.RowShipper(fm.array, slot_bytes=64 << 10, slots=3, copy_threads=2)
        sh.ship(src[0:300], 0, 0)
        sh.ship(src[300:, 300:], 300, 300)   # right part of a block row
        sh.ship(src[300:, :300], 300, 0)     # left part
        sh.finish()
        assert sh.bytes_shipped == n * n * 4
    assert np.array_equal(np.fromfile(path, np.float32).reshape(n, n), src.cpu().numpy())


def test_host_register_makes_a_mapping_a_dma_target(tmp_path):
    lib = _lib.load()
    shm = "/dev/shm" if os.path.isdir("/dev/shm") else str(tmp_path)
    path = os.path.join(shm, "po_test_register_%d.bin" % os.getpid())
    try:
        with hostsink.FileMatrix(path, 256, 512, np.float32, create=True) as fm:
            lib.po_host_prefault(fm.array.ctypes.data, fm.nbytes, 2)
            if not fm.register():
                pytest.skip("the kernel refuses to page-lock this mapping: " + lib.po_last_error().decode())
            src = torch.rand((256, 512), dtype=torch.float32, device="cuda")
            host = torch.from_numpy(fm.array)
            engine.copy2d(host[10:200, 16:400], src[10:200, 16:400])
            torch.cuda.synchronize()
            want = np.zeros((256, 512), np.float32)
            want[10:200, 16:400] = src[10:200, 16:400].cpu().numpy()
            assert np.array_equal(fm.array, want)
            del host
    finally:
        if os.path.exists(path):
            os.unlink(path)


def test_file_to_device_matches_file(tmp_path):
    data = np.random.default_rng(5).integers(0, 255, size=3_000_001, dtype=np.uint8)
    path = os.path.join(tmp_path, "blob")
    data.tofile(path)
    d = engine.file_to_device(path, slot_bytes=1 << 20)
    assert d.shape[0] == data.shape[0] + 64 and (d[-64:] == 10).all()
    assert np.array_equal(d[:-64].cpu().numpy(), data)
    d = engine.file_to_device(path, 1000, 2_000_123, slot_bytes=1 << 20)
    assert np.array_equal(d[:-64].cpu().numpy(), data[1000:2_000_123])


def test_in_ram_matrix_streams_in_panels_when_it_does_not_fit(profiles, monkeypatch):
    """--large None convention: float64 (N, N) in RAM.  Same values whether the device keeps the whole
    matrix (symmetric panels) or only two panel buffers."""
    X = profiles[:140]
    a = phyloligo.compute_distances_device(X, "JSD")
    monkeypatch.setattr(phyloligo, "RESIDENT_FRACTION", 0.0)
    monkeypatch.setattr(phyloligo, "PANEL_ROWS", 128)
    b = phyloligo.compute_distances_device(X, "JSD")
    assert a.dtype == np.float64 and np.array_equal(a, b)
    assert np.allclose(a, _oracle_matrix(X, "JSD"), rtol=1e-6, atol=1e-12)
