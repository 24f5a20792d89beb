"""Host end of the output path (no GPU needed): the po_host_* entry points and hostsink.FileMatrix."""
import ctypes as C
import os

import numpy as np
import pytest

from phyloligo_b200 import _lib, hostsink


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def test_copy2d_pwrite2d_pread_roundtrip(lib, tmp_path):
    rng = np.random.default_rng(1)
    a = rng.random((501, 333)).astype(np.float32)
    c = np.zeros((501, 400), np.float32)
    assert lib.po_host_copy2d(c.ctypes.data, 1600, a.ctypes.data, 333 * 4, 333 * 4, 501, 3) == 0
    assert np.array_equal(c[:, :333], a) and not c[:, 333:].any()
    path = os.path.join(tmp_path, "m.bin")
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    os.ftruncate(fd, 64 + 501 * 1600)
    assert lib.po_host_pwrite2d(fd, 64, 1600, a.ctypes.data, 333 * 4, 333 * 4, 501, 4) == 0
    back = np.zeros(501 * 1600, np.uint8)
    assert lib.po_host_pread(fd, 64, back.ctypes.data, back.nbytes, 3) == 0
    os.close(fd)
    assert np.array_equal(back.view(np.float32).reshape(501, 400)[:, :333], a)
    assert lib.po_host_pread(-1, 0, back.ctypes.data, 8, 1) < 0
    assert lib.po_host_copy2d(c.ctypes.data, 4, a.ctypes.data, 4, 8, 1, 1) < 0  # pitch < width
    assert b"geometry" in lib.po_last_error()


@pytest.mark.parametrize("rows,cols", [(1, 1), (4, 4), (63, 65), (130, 257), (1000, 777)])
def test_transpose(lib, rows, cols):
    a = np.random.default_rng(rows).random((rows, cols + 3)).astype(np.float32)
    out = np.full((cols, rows + 5), -1.0, np.float32)
    assert lib.po_host_transpose_f32(out.ctypes.data, rows + 5, a.ctypes.data, cols + 3, rows, cols, 3) == 0
    assert np.array_equal(out[:, :rows], a[:, :cols].T) and (out[:, rows:] == -1).all()


def test_file_matrix_create_attach_warm(tmp_path):
    path = os.path.join(tmp_path, "d.mat")
    with hostsink.FileMatrix(path, 300, 300, np.float32, create=True) as fm:
        assert os.path.getsize(path) == 300 * 300 * 4
        fm.warm([(0, 100), (200, 300)], threads=2)
        fm.warmer.join()
        assert fm.warmer.error is None
        fm.array[5, 7] = 3.5
        with hostsink.FileMatrix(path, 300, 300, np.float32, create=False) as other:  # a second rank attaching
            other.array[299, 299] = 1.25
    m = np.fromfile(path, np.float32).reshape(300, 300)
    assert m[5, 7] == 3.5 and m[299, 299] == 1.25 and m.sum() == 4.75
    # a region at an offset (the data region of an HDF5 file)
    with open(path, "r+b") as fh:
        fh.truncate(4096 + 300 * 300 * 4)
    with hostsink.FileMatrix(path, 300, 300, np.float32, offset=4096, create=False) as fm:
        fm.warm([(0, 300)], threads=1)
        fm.array[0, 0] = 9.0
    assert np.fromfile(path, np.float32)[1024] == 9.0
    with pytest.raises(_lib.PhyloligoError):
        hostsink.FileMatrix(path, 4000, 4000, np.float32, create=False)  # file too small for that shape


def test_prefault_is_harmless_on_written_pages(lib, tmp_path):
    path = os.path.join(tmp_path, "p.bin")
    m = np.memmap(path, dtype=np.float32, mode="w+", shape=(1 << 16,))
    m[:] = np.arange(1 << 16, dtype=np.float32)
    assert lib.po_host_prefault(m.ctypes.data + 100, m.nbytes - 200, 2) == 0
    assert np.array_equal(m, np.arange(1 << 16, dtype=np.float32))
    assert lib.po_host_prefault(None, 0, 1) == 0
