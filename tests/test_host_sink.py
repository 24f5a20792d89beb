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


def _aligned_f32(n, align=64, fill=None):
    raw = np.empty(n * 4 + align, np.uint8)
    off = (-raw.ctypes.data) % align
    a = raw[off:off + n * 4].view(np.float32)
    if fill is not None:
        a[:] = fill
    return a


@pytest.mark.parametrize("rows,cols,ld_dst,col_off,row_off", [
    (24, 8, 32, 0, 0),          # the smallest block the vector path takes
    (200, 130, 256, 0, 0),      # pitch % 16 == 0: 64-byte phase, ragged rows and columns
    (200, 130, 264, 0, 0),      # pitch % 8 == 0 only: 32-byte phase
    (333, 97, 512, 5, 3),       # destination and source start off the alignment grid
    (333, 97, 520, 7, 1),
    (129, 64, 257, 0, 0),       # odd pitch: the scalar / SSE form
    (1024, 4096 - 64, 1056, 16, 0),
])
def test_transpose_vector_path_alignments(lib, rows, cols, ld_dst, col_off, row_off):
    """The AVX2 form (streaming stores, aligned from the first row whose address allows it) against numpy,
    for every alignment case: nothing outside the destination block may change."""
    rng = np.random.default_rng(rows * 31 + cols)
    ld_src = cols + 11
    src = _aligned_f32(rows * ld_src + row_off)[row_off:]
    src[:] = rng.random(src.shape[0]).astype(np.float32)
    S = src[: rows * ld_src].reshape(rows, ld_src)
    assert ld_dst >= rows + col_off
    dst = _aligned_f32((cols + 2) * ld_dst, fill=-7.0)
    D = dst.reshape(cols + 2, ld_dst)
    target = D[1:, col_off:]
    rc = lib.po_host_transpose_f32(target.ctypes.data, ld_dst, S.ctypes.data, ld_src, rows, cols, 4)
    assert rc == 0
    assert np.array_equal(D[1:cols + 1, col_off:col_off + rows], S[:, :cols].T)
    D[1:cols + 1, col_off:col_off + rows] = -7.0
    assert (dst == -7.0).all()


def test_mirror_pool_builds_lower_triangle(lib):
    """po_host_mirror_*: blocks submitted for immediate release (no stream) are transposed by the pool;
    the lower triangle of a symmetric matrix is rebuilt panel by panel from its upper triangle."""
    n, panel = 1000, 128
    rng = np.random.default_rng(5)
    full = rng.random((n, n)).astype(np.float32)
    full = np.triu(full) + np.triu(full, 1).T
    host = _aligned_f32(n * n).reshape(n, n)
    host[:] = np.triu(full)
    pool = lib.po_host_mirror_open(3)
    assert pool
    for r0 in range(0, n, panel):
        r1 = min(n, r0 + panel)
        if r1 == n:
            break
        dst, src = host[r1:, r0:r1], host[r0:r1, r1:]
        rc = lib.po_host_mirror_submit(pool, None, 0, dst.ctypes.data, n, src.ctypes.data, n, r1 - r0, n - r1)
        assert rc == 0
    assert lib.po_host_mirror_wait(pool) == 0
    assert lib.po_host_mirror_wait(pool) == 0  # idempotent
    # the diagonal blocks' lower halves were not submitted: mirror them here
    for r0 in range(0, n, panel):
        r1 = min(n, r0 + panel)
        blk = host[r0:r1, r0:r1]
        blk[:] = np.triu(blk) + np.triu(blk, 1).T
    assert np.array_equal(host, full)
    assert lib.po_host_mirror_submit(pool, None, 0, host.ctypes.data, 4, host.ctypes.data, 4, 8, 8) < 0  # pitch < extent
    assert b"geometry" in lib.po_last_error()
    assert lib.po_host_mirror_submit(pool, None, 0, None, 8, None, 8, 0, 8) == 0  # empty block: nothing to do
    assert lib.po_host_mirror_close(pool) == 0
    assert lib.po_host_mirror_close(None) == 0


@pytest.mark.parametrize("where", ["tmp", "shm"])
def test_premap_existing_file(lib, tmp_path, where):
    """po_host_premap maps the pages of an existing file ahead of the writes (tmpfs: populate for reading, the
    entries are writable at once; elsewhere: populate for writing); what is written afterwards reaches the file."""
    folder = str(tmp_path) if where == "tmp" else "/dev/shm"
    if not os.path.isdir(folder):
        pytest.skip("no " + folder)
    path = os.path.join(folder, "po_premap_%d.mat" % os.getpid())
    try:
        with hostsink.FileMatrix(path, 257, 300, np.float32, create=True) as fm:
            fm.array[:] = 1.0  # the pages exist and are up to date
        with hostsink.FileMatrix(path, 257, 300, np.float32, create=False) as fm:
            assert not fm.fresh
            assert lib.po_host_premap(fm.fd, fm.array.ctypes.data, fm.nbytes, 3) == 0
            fm.array[5] = 7.0
            fm.warm([(0, 257)], threads=2)  # the warmer's mode for an existing file
            fm.warmer.join()
            assert fm.warmer.error is None and fm.warmer.mode == "premap"
            fm.array[200, 299] = -3.0
        m = np.fromfile(path, np.float32).reshape(257, 300)
        assert (m[5] == 7.0).all() and m[200, 299] == -3.0 and m[0, 0] == 1.0 and m.sum() == 257 * 300 + 300 * 6 - 4
        assert lib.po_host_premap(-1, None, 0, 1) == 0
        assert lib.po_host_premap(-1, None, 8, 1) < 0
    finally:
        if os.path.exists(path):
            os.unlink(path)


def test_file_matrix_page_spans_merge_neighbours(tmp_path):
    """FileMatrix.page_spans (what register_rows page-locks): page-aligned, inside the file, neighbours merged."""
    path = os.path.join(tmp_path, "s.mat")
    with hostsink.FileMatrix(path, 1000, 1000, np.float32, create=True) as fm:  # 4000-byte rows, 4096-byte pages
        spans = fm.page_spans([(640, 896), (128, 384), (384, 512), (0, 0)])
        # rows [128, 512) and [640, 896) are 512 000 bytes apart: two spans; [128, 384) + [384, 512) merge
        assert len(spans) == 2 and spans[0][0] == 512000 // 4096 * 4096 and spans[0][1] == -(-2048000 // 4096) * 4096
        assert spans[1] == [2560000 // 4096 * 4096, -(-3584000 // 4096) * 4096]
        for lo, hi in spans:
            assert lo % 4096 == 0 and (hi % 4096 == 0 or hi == fm.nbytes) and 0 <= lo < hi <= fm.nbytes
        assert fm.page_spans([(990, 1000)])[0][1] == fm.nbytes  # clipped to the end of the matrix
        # two block rows that meet inside one page become one span
        assert len(fm.page_spans([(0, 5), (5, 9)])) == 1
    with hostsink.FileMatrix(path, 10, 10, np.float32, offset=4096, create=True) as fm:  # a region at an offset (HDF5)
        assert fm.page_spans([(0, 10)]) == [[4096, 4096 + 400]]


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
