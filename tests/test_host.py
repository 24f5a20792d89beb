"""CPU tests of the host layer: the C-ABI library loads and exports every declared
symbol, pattern geometry, FASTA indexing, file formats, CLI surface, and the loud
failure without a GPU (no compute calls are made here)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import phylo_oracle as po
from phyloligo_b200 import _lib, engine, io_formats, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "phyloligo_b200.h")).read()
    declared = set(re.findall(r"\b(po_[a-z0-9_]+)\s*\(", header))
    declared -= {"po_stream_t"}
    assert declared == set(_lib.EXPORTED)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.po_version()


def test_pattern_info_and_limits():
    assert _lib.pattern_info("1111") == (4, 4, 256)
    assert _lib.pattern_info("111010011") == (9, 6, 4096)
    assert _lib.pattern_info("5") == (1, 0, 1)       # `-p 5` is the string "5": no '1' at all
    assert _lib.pattern_info("100a1") == (5, 2, 16)  # anything but '1' is a don't-care position
    with pytest.raises(_lib.PhyloligoError):
        _lib.pattern_info("1" * 33)
    with pytest.raises(_lib.PhyloligoError):
        _lib.pattern_info("")
    with pytest.raises(_lib.PhyloligoError):
        _lib.pattern_info("1" * 11)


def test_fasta_index_matches_oracle_reader(tmp_path):
    seqs = synth.make_sequences(50, 700, seed=2) + [b"", b"ACGT"]
    text = b"; stray text before the first header\n" + synth.to_fasta_bytes(seqs, line=70)
    text = text.replace(b">c3\n", b">c3 some description > with a bracket\n")
    path = tmp_path / "x.fa"
    path.write_bytes(text)
    begin, end = engine.fasta_index(text)
    assert len(begin) == len(seqs)
    want = list(po.read_fasta(str(path)))
    got = [bytes(text[b:e]).replace(b"\n", b"").decode() for b, e in zip(begin, end)]
    assert got == want == [s.decode() for s in seqs]
    # no trailing newline, CRLF, multi-threaded scan on a larger buffer
    big = synth.to_fasta_bytes(synth.make_sequences(3000, 600, seed=3))
    b1, e1 = engine.fasta_index(big, threads=1)
    b8, e8 = engine.fasta_index(big, threads=8)
    assert np.array_equal(b1, b8) and np.array_equal(e1, e8) and len(b1) == 3000
    b, e = engine.fasta_index(b">a\r\nAC\r\nGT\r\n>b\r\nTT")
    assert b.tolist() == [4, 16] and e.tolist() == [12, 18]
    assert engine.fasta_index(b"")[0].shape == (0,)


def test_hdf5_and_memmap_formats(tmp_path):
    a = np.random.default_rng(0).random((37, 19)).astype(np.float32)
    p = str(tmp_path / "d.h5")
    io_formats.write_hdf5(p, "distances", a)
    assert np.array_equal(io_formats.read_hdf5(p, "distances"), a)
    raw = open(p, "rb").read()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and b"TREE" in raw and b"HEAP" in raw and b"SNOD" in raw
    with pytest.raises(KeyError):
        io_formats.read_hdf5(p, "frequencies")
    with io_formats.Hdf5DatasetWriter(p, "frequencies", (10, 4), np.float64) as w:
        for r in range(0, 10, 3):
            w.write_rows(r, np.full((min(3, 10 - r), 4), float(r)))
    got = io_formats.read_hdf5(p, "frequencies")
    assert got.dtype == np.float64 and got[9, 0] == 9.0 and got[4, 3] == 3.0
    m = np.arange(49, dtype=np.float32)
    mp = str(tmp_path / "m.raw")
    m.tofile(mp)
    assert io_formats.read_memmap(mp).shape == (7, 7)
    t = str(tmp_path / "t.txt")
    io_formats.savetxt(t, a.astype(np.float64))
    first = open(t).readline().split("\t")
    assert len(first) == 19 and re.fullmatch(r"\d\.\d{18}e[+-]\d\d", first[0])
    assert np.array_equal(io_formats.read_numpy(t), a.astype(np.float64))


def test_cli_surface_matches_reference_flags():
    from phyloligo_b200 import phyloligo
    p = phyloligo.get_cmd(["-i", "x.fa", "--method", "joblib"])
    assert p.pattern == 4 and p.strand == "both" and p.dist == "Eucl" and p.large == "None"
    assert p.threads_max == 4 and p.out_file == "phyloligo.out" and p.freqchunksize == 250 and p.distchunksize == 250
    assert os.path.isabs(p.workdir)
    assert phyloligo.get_cmd(["-i", "x", "--method", "joblib", "-k", "5", "-p", "1101"]).pattern == "1101"
    assert phyloligo.get_cmd(["-i", "x", "--method", "scoop", "-p", "1101", "-k", "5"]).pattern == 5
    assert phyloligo.get_cmd(["-i", "x", "--method", "joblib", "-p", "5"]).pattern == "5"
    with pytest.raises(SystemExit):
        phyloligo.get_cmd(["-i", "x"])  # --method is required
    with pytest.raises(SystemExit):
        phyloligo.get_cmd(["-i", "x", "--method", "joblib", "-d", "XX"])


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(tmp_path):
    from phyloligo_b200 import phyloligo, phylodist
    with pytest.raises(_lib.PhyloligoError):
        phyloligo.compute_frequency("ACGTACGT", "11", "both")
    with pytest.raises(_lib.PhyloligoError):
        phylodist.JSD(np.ones(4) / 4, np.ones(4) / 4)
    fa = tmp_path / "a.fa"
    fa.write_bytes(b">a\nACGT\n")
    with pytest.raises(_lib.PhyloligoError):
        phyloligo.compute_frequencies("joblib", "None", str(fa), "11", "both", 250, 4, str(tmp_path))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phyloligo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("phylo_oracle", "oracle") or f == "_none_", (dirpath, f)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_savetxt_writes_the_bytes_numpy_writes(tmp_path, dtype):
    """po_savetxt_host against np.savetxt(delimiter='\\t') -- the reference's writer, bin/phyloligo.py:1059-1066."""
    rng = np.random.default_rng(0)
    span = 300 if dtype == np.float64 else 30
    A = (rng.standard_normal((211, 97)) * 10.0 ** rng.integers(-span, span, size=(211, 97))).astype(dtype)
    A[0, 0], A[1, 1], A[2, 2], A[3, 3], A[4, 4] = np.nan, np.inf, -np.inf, 0.0, -0.0
    A[5, 5] = np.finfo(dtype).tiny / 4  # subnormal
    A[6, :] = rng.random(97)            # the range frequencies and distances live in
    ours, ref = os.path.join(tmp_path, "a.txt"), os.path.join(tmp_path, "b.txt")
    for threads in (0, 1, 3):
        io_formats.savetxt(ours, A, threads=threads)
        np.savetxt(ref, A, delimiter="\t")
        assert open(ours, "rb").read() == open(ref, "rb").read()
    for arr in (np.zeros((0, 5)), np.arange(7), np.zeros((3, 4), dtype=np.int64)):  # empty, 1-D, the all-zero int rows of :660
        io_formats.savetxt(ours, arr)
        np.savetxt(ref, arr, delimiter="\t")
        assert open(ours, "rb").read() == open(ref, "rb").read()


def test_host_binding_is_a_no_op_without_topology():
    before = os.sched_getaffinity(0)
    assert engine.bind_host_to_gpu_node(0) is None or isinstance(engine.bind_host_to_gpu_node(0), int)
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before


def test_kount_window_table_matches_the_oracle_windows():
    """kount.window_table (vectorised) against the restated make_genome_chunk rules (reference
    bin/Kount.py:343-407, pinned by tests/golden/kount_golden.json) over many lengths and parameters."""
    from oracle import kount_oracle as ko
    from phyloligo_b200 import kount

    rng = np.random.default_rng(4)
    for w, t in ((5000, 500), (300, 50), (250, 100), (1000, 999), (64, 1), (7, 3), (100, 100), (101, 50)):
        lengths = sorted(set([0, 1, w - 1, w, w + 1, 20 * t - 1, 20 * t, 20 * t + 1, w + t, w + 20 * t]
                             + [int(v) for v in rng.integers(0, 30 * max(w, t), 25)]))
        lengths = [n for n in lengths if n >= 0 and n < 400_000]
        records = [("c%d" % i, "A" * n) for i, n in enumerate(lengths)]
        want = ko.make_windows(records, w, t)
        rec, start, size, dstart, dstop = kount.window_table(np.array(lengths), w, t)
        assert rec.shape[0] == len(want)
        got = [("c%d" % r, int(a), int(b), int(m)) for r, a, b, m in zip(rec, dstart, dstop, size)]
        assert got == [(sid, a, b, len(s)) for sid, a, b, s in want], (w, t)
        # window start offsets: every window string is the record's slice [start, start + size)
        assert all(int(s0) + int(m) <= lengths[int(r)] for r, s0, m in zip(rec, start, size))


def test_hdf5_attach_lets_other_processes_fill_the_dataset(tmp_path):
    """The multi-GPU command line: rank 0 creates the one-dataset file, every rank attaches to the
    data region and writes its rows; the file reads back as one matrix."""
    path = os.path.join(tmp_path, "d.h5")
    io_formats.Hdf5DatasetWriter(path, "distances", (7, 5), np.float32).close()
    shape, dtype, addr = io_formats.dataset_location(path, "distances")
    assert shape == (7, 5) and np.dtype(dtype) == np.float32 and addr % 8 == 0
    want = np.arange(35, dtype=np.float32).reshape(7, 5)
    for rows in ((0, 3), (3, 7)):  # two "ranks"
        with io_formats.Hdf5DatasetWriter.attach(path, "distances") as w:
            w.write_rows(rows[0], want[rows[0]:rows[1]])
    assert np.array_equal(io_formats.read_hdf5(path, "distances"), want)
    assert os.path.getsize(path) == addr + 35 * 4  # the data region ends the file (superblock EOF address)


def test_kount_assembly_layout_matches_the_fasta_reader(tmp_path):
    """kount.Assembly (host side only): record ids as Biopython's record.id, sequences without line
    breaks, one separator between records."""
    from oracle import kount_oracle as ko
    from phyloligo_b200 import kount

    path = os.path.join(tmp_path, "a.fasta")
    with open(path, "wb") as fh:
        fh.write(b"junk before the first record\n>c1 first contig\nACGT\nacgtNN\n\n>c2\n>c3|x desc\r\nAC GT\t\r\nTT \x0b\r\n>last\nGATTACA")
    asm = kount.Assembly(path)
    want = ko.read_records(path)
    assert asm.ids == [w[0] for w in want] == ["c1", "c2", "c3|x", "last"]
    assert [asm.sequence(i) for i in range(asm.n)] == [w[1] for w in want]
    assert list(asm.lengths) == [10, 0, 6, 7]
    for i in range(asm.n):  # every record is followed by exactly one separator
        assert asm.text[int(asm.offsets[i] + asm.lengths[i])] == 10


def test_savetxt_digits_are_correctly_rounded_everywhere(tmp_path):
    """The printf-free '%.18e' path of po_savetxt_host (128-bit scaled arithmetic, snprintf only when a
    rounding boundary cannot be decided) against Python's own formatting: random bit patterns over the
    whole double range (subnormals included), powers of two and ten, exact ties at the 19th digit."""
    rng = np.random.default_rng(1)
    vals = rng.integers(0, 2 ** 63, size=120_000, dtype=np.int64).astype(np.uint64).view(np.float64)
    vals = vals[np.isfinite(vals)]
    vals[::2] *= -1.0
    spec = [0.0, -0.0, 1.0, 0.5, 0.1, 1 / 3, 1e22, 1e23, 9.999999999999999e22, 5e-324, 2.2250738585072014e-308,
            2.225073858507201e-308, 1.7976931348623157e308, 1e-20, 1e19, 9.5, 99.5, 4.35, 3.0517578125e-05]
    spec += [float(2.0 ** i) for i in range(-1074, 1024, 7)] + [float(10 ** i) for i in range(23)]
    spec += [1.0 / float(10 ** i) for i in range(1, 23)]
    spec += [(2 * i + 1) / 2.0 ** 20 for i in range(0, 3000, 7)]          # 20 significant digits ending in 5: ties
    spec += [(2 * i + 1) / 2.0 ** 40 for i in range(10 ** 6, 10 ** 6 + 300)]
    spec += [float(i) / 1024 for i in range(1, 300)]
    allv = np.concatenate([vals, np.array(spec, dtype=np.float64)])
    path = os.path.join(tmp_path, "v.txt")
    io_formats.savetxt(path, allv.reshape(-1, 1), threads=3)
    got = open(path).read().split("\n")[:-1]
    assert len(got) == allv.shape[0]
    bad = [(float(x), g) for x, g in zip(allv, got) if g != "%.18e" % float(x)]
    assert not bad, bad[:5]
    f32 = rng.random(50_000).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 30, 50_000).astype(np.float32)
    io_formats.savetxt(path, f32.reshape(-1, 1))
    got = open(path).read().split("\n")[:-1]
    assert all(g == "%.18e" % float(x) for x, g in zip(f32, got))


def test_fasta_index_fuzz_against_the_reader(tmp_path):
    """Random line soups (headers anywhere, '>' inside lines, blank lines, CRLF, lone-CR line ends, blanks inside
    lines, no final newline, text before the first header) : po_fasta_index_host yields the records
    SeqIO.parse would."""
    from hypothesis import given, settings, strategies as st

    line = st.one_of(
        st.text(alphabet="ACGTNacgtn >\t\x0c", min_size=0, max_size=30),
        st.text(alphabet="ACGT", min_size=0, max_size=5).map(lambda s: ">" + s + " desc"),
        st.just(""), st.just(">"), st.just(">>x"))
    path = os.path.join(tmp_path, "fuzz.fa")

    @settings(max_examples=300, deadline=None)
    @given(st.lists(line, min_size=0, max_size=25), st.sampled_from(["\n", "\r\n", "\r"]), st.booleans())
    def run(lines, eol, final_eol):
        text = eol.join(lines) + (eol if final_eol and lines else "")
        raw = text.encode()
        with open(path, "wb") as fh:
            fh.write(raw)
        want = list(po.read_fasta(path))
        begin, end = engine.fasta_index(raw, threads=2)
        got = [bytes(raw[b:e]).decode().replace("\n", "").replace("\r", "").replace(" ", "").replace("\t", "").replace("\x0c", "")
               for b, e in zip(begin, end)]
        assert got == want

    run()


def test_header_starts_and_per_cluster_fasta(tmp_path):
    """engine.fasta_header_starts (the '>' of every record, also when a title contains '>' or the file uses CRLF /
    lone-CR line ends) and phyloselect.write_fastafile on top of it: titles unchanged, sequences wrapped at 60."""
    from phyloligo_b200 import phyloselect

    for eol in (b"\n", b"\r\n", b"\r"):
        raw = eol.join([b"junk > before", b">r1 a>b c", b"ACGT" * 20, b"AC", b">r2", b">r3 x", b"TT GG", b""])
        begin, end = engine.fasta_index(raw)
        starts = engine.fasta_header_starts(raw, begin, end)
        titles = [raw[int(s):int(b)].rstrip(b"\r\n") for s, b in zip(starts, begin)]
        assert titles == [b">r1 a>b c", b">r2", b">r3 x"], (eol, titles)
        path = os.path.join(tmp_path, "in.fa")
        with open(path, "wb") as fh:
            fh.write(raw)
        out = os.path.join(tmp_path, "out_%d" % len(eol + eol[:1]))
        os.makedirs(out, exist_ok=True)
        phyloselect.write_fastafile(np.array([0, 1, 0]), path, out)
        assert open(os.path.join(out, "data_fasta_cl0.fa"), "rb").read() == \
            b">r1 a>b c\n" + b"ACGT" * 15 + b"\n" + b"ACGT" * 5 + b"AC\n" + b">r3 x\nTTGG\n"
        assert open(os.path.join(out, "data_fasta_cl1.fa"), "rb").read() == b">r2\n"
