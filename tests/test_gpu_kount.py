"""Kount.py sliding-window mode on the GPU against the goldens produced by the reference's own
function bodies (tests/golden/kount_golden.json) and against the oracle on larger random assemblies."""
import json
import os

import numpy as np
import pytest

from oracle import kount_oracle as ko
from phyloligo_b200 import kount, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


class _Opt:
    def __init__(self, strand, n_max):
        self.strand, self.n_max_freq_in_windows = strand, n_max


def _rows(genome, mcp, dist, case):
    out = []
    for chunk in kount.sliding_windows_distances(genome, mcp, dist, case["pattern"], case["window"], case["step"],
                                                 _Opt(case["strand"], case["n_max"])):
        out.extend(chunk)
    return out


def test_golden_windows_and_distances(tmp_path):
    cases = json.load(open(os.path.join(HERE, "golden", "kount_golden.json")))
    for ci, case in enumerate(cases):
        path = os.path.join(tmp_path, "asm%d.fasta" % ci)
        with open(path, "w") as fh:
            fh.write(case["fasta"])
        mcp = kount.compute_whole_composition(path, case["pattern"], case["strand"])
        want_mcp = np.array([float.fromhex(v) for v in case["mcp"]])
        assert np.array_equal(mcp, want_mcp)  # integer counts, one float64 division: bit exact
        for dist, rows in case["rows"].items():
            got = _rows(path, mcp, dist, case)
            assert [r[:3] for r in got] == [r[:3] for r in rows]  # ids and displayed coordinates
            g = np.array([r[3] for r in got])
            w = np.array([float.fromhex(r[3]) for r in rows])
            assert np.allclose(g, w, rtol=1e-11, atol=1e-13), (ci, dist, np.abs(g - w).max())
        # the reference's chunk generator, window strings included
        info, seqs = [], []
        for a, b in kount.make_genome_chunk(path, case["window"], case["step"], None, 37):
            info.extend(a)
            seqs.extend(b)
        want = ko.make_windows(ko.read_records(path), case["window"], case["step"])
        assert [tuple(i) for i in info] == [w_[:3] for w_ in want] and seqs == [w_[3] for w_ in want]


@pytest.mark.parametrize("dist,pattern,strand", [("JSD", "11111", "both"), ("KL", "1101", "plus"), ("Eucl", "111", "minus")])
def test_random_assembly_against_the_oracle(tmp_path, dist, pattern, strand):
    seqs = synth.make_sequences(40, 9000, seed=21) + [b"ACGT" * 10, b"N" * 7000, b""]
    path = os.path.join(tmp_path, "asm.fasta")
    synth.write_fasta(path, seqs, line=73)
    records = ko.read_records(path)
    mcp = kount.compute_whole_composition(path, pattern, strand)
    assert np.array_equal(mcp, ko.compute_whole_composition(records, pattern, strand))
    case = {"pattern": pattern, "window": 2000, "step": 300, "strand": strand, "n_max": 0.02}
    got = _rows(path, mcp, dist, case)
    k = pattern.count("1")
    want = []
    for sid, a, b, window in ko.make_windows(records, 2000, 300):
        if len(window) == 0:
            want.append([sid, a, b, None])  # the reference divides by zero here
        elif (window.count("N") / len(window)) > 0.02:
            want.append([sid, a, b, 0.0])   # all-NaN profile, every term zeroed (also where ksize**4 != 4**ksize)
        else:
            want.append([sid, a, b, ko.compute_distance(dist, mcp, window, pattern, strand, 0.02)])
    assert [r[:3] for r in got] == [r[:3] for r in want]
    for g, w in zip(got, want):
        if w[3] is not None:
            assert g[3] == pytest.approx(w[3], rel=1e-11, abs=1e-13)
    assert any(w[3] == 0.0 for w in want) and len(want) > 500


def test_worker_functions_and_command_line(tmp_path, capsys):
    rng = np.random.default_rng(1)
    a = rng.dirichlet(np.ones(256))
    b = rng.dirichlet(np.ones(256))
    a[:7] = 0.0
    assert kount.JSD(a, b) == pytest.approx(ko.JSD(a, b), rel=1e-12)
    assert kount.KL(a, b) == pytest.approx(ko.KL(a, b), rel=1e-12)
    assert kount.Eucl(a, b) == pytest.approx(ko.Eucl(a, b), rel=1e-12)
    seq = "ACGTTGCANNACGTAGCTAGCTAGGATCCGATCGATTTAGC" * 9
    assert np.array_equal(kount.compute_frequency(seq, 0.4, "1111", "both"), ko.compute_frequency(seq, 0.4, "1111", "both"))
    assert np.isnan(kount.compute_frequency("NNNNACGT", 0.4, "11", "plus")).all()
    seqs = synth.make_sequences(12, 7000, seed=3)
    path = os.path.join(tmp_path, "asm.fasta")
    synth.write_fasta(path, seqs, line=60)
    kount.main(["-i", path, "-k", "4", "-w", "1000", "-t", "200", "-d", "JSD", "-W", str(tmp_path)])
    out = capsys.readouterr().out.splitlines()
    assert out[0] == "Genome : {}".format(path) and out[1] == "Contaminant : None"
    result = os.path.join(tmp_path, "asm.fasta.mcp_windows_vs_whole_JSD.dist")
    lines = [l.rstrip("\n").split("\t") for l in open(result)]
    records = ko.read_records(path)
    mcp = ko.compute_whole_composition(records, "1111", "both")
    want = ko.sliding_windows_distances(records, mcp, "JSD", "1111", 1000, 200, "both", 0.4)
    assert len(lines) == len(want)
    for l, w in zip(lines, want):
        assert l[0] == w[0] and int(l[1]) == w[1] and int(l[2]) == w[2]
        assert float(l[3]) == pytest.approx(w[3], rel=1e-11, abs=1e-13)


def test_command_line_against_the_unmodified_reference_script(tmp_path):
    """Output files of the reference's unmodified bin/Kount.py (tests/golden/make_kount_cli_golden.py) against
    this command line: same file names, same rows, ids and displayed coordinates identical, distances to 1e-11
    (float64 logs and summation order differ from numpy's in the last bits)."""
    import sys
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import make_kount_cli_golden as mk
    golden = dict(np.load(os.path.join(HERE, "golden", "kount_cli_golden.npz")))
    path = os.path.join(tmp_path, "asm.fasta")
    open(path, "wb").write(mk.assembly())
    assert len(golden["names"]) == 3
    for name in golden["names"]:
        outdir = os.path.join(tmp_path, str(name))
        kount.main(["-i", path, "-u", "2", "-W", outdir] + str(golden[str(name) + "_args"]).split())
        files = sorted(os.listdir(outdir))
        assert files == [str(f) for f in golden[str(name) + "_files"]], name
        for f in files:
            want = [l.split("\t") for l in golden[str(name) + "_" + f].tobytes().decode().splitlines()]
            got = [l.split("\t") for l in open(os.path.join(outdir, f)).read().splitlines()]
            assert [r[:3] for r in got] == [r[:3] for r in want], (name, f)
            g, w = np.array([float(r[3]) for r in got]), np.array([float(r[3]) for r in want])
            assert np.allclose(g, w, rtol=1e-11, atol=1e-13), (name, f, np.abs(g - w).max())
