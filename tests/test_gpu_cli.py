"""End-to-end through the reference-facing command line (phyloligo.py main) on the GPU:
every metric in every output mode, read back through the formats the reference's own
readers expect (bin/phyloligo_comparemat.py:7-24), against the oracle."""
import os

import numpy as np
import pytest

from oracle import phylo_oracle as po
from phyloligo_b200 import io_formats, phyloligo, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fasta(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    seqs = synth.make_sequences(70, 1500, seed=9) + [b"", b"NNNNNN"]
    path = os.path.join(d, "asm.fasta")
    synth.write_fasta(path, seqs, line=70)
    return path, [s.decode() for s in seqs]


def _oracle(seqs, pattern, strand, metric, dtype):
    X = np.vstack([po.frequency_np(s, pattern, strand) for s in seqs]).astype(dtype).astype(np.float64)
    n = len(seqs)
    if metric in ("Eucl", "JSD", "BC"):
        return po.pairwise_np(X, metric)
    fn = po.KT if metric == "KT" else po.SC
    return np.array([[fn(X[i], X[j]) for j in range(n)] for i in range(n)])


def _check(got, want, metric):
    assert got.shape == want.shape
    assert np.array_equal(np.isnan(got), np.isnan(want))
    m = ~np.isnan(want)
    tol = 1e-12 if metric in ("KT", "SC") else 1e-6
    assert np.allclose(got[m], want[m], rtol=tol, atol=1e-12), np.abs(got[m] - want[m]).max()


@pytest.mark.parametrize("metric", ["Eucl", "JSD", "KT", "BC", "SC"])
@pytest.mark.parametrize("method", ["joblib", "scoop"])
def test_text_output_matches_oracle(fasta, tmp_path, metric, method, capsys):
    path, seqs = fasta
    out = os.path.join(tmp_path, "m.txt")
    freq = os.path.join(tmp_path, "f.txt")
    phyloligo.main(["-i", path, "-k", "4", "-d", metric, "--method", method, "-o", out, "-q", freq,
                    "-w", str(tmp_path)])
    lines = capsys.readouterr().out.splitlines()
    assert lines[:3] == ["Using pattern 1111", "Computing frequencies", "Computing Pairwise distances"]
    assert "Writing frequency matrix" in lines and "Writing distance matrix" in lines
    F = np.loadtxt(freq)
    assert np.array_equal(F, np.vstack([po.frequency_np(s, "1111", "both") for s in seqs]))  # bit exact, %.18e
    _check(io_formats.read_numpy(out), _oracle(seqs, "1111", "both", metric, np.float64), metric)


@pytest.mark.parametrize("metric", ["Eucl", "JSD", "KT", "BC", "SC"])
@pytest.mark.parametrize("large", ["memmap", "h5py"])
def test_large_modes_match_oracle(fasta, tmp_path, metric, large):
    path, seqs = fasta
    out = os.path.join(tmp_path, "m.bin")
    freq = os.path.join(tmp_path, "f.txt")
    phyloligo.main(["-i", path, "-p", "110011", "-s", "plus", "-d", metric, "--method", "joblib", "--large", large,
                    "-o", out, "-q", freq, "-w", str(tmp_path)])
    got = io_formats.read_memmap(out) if large == "memmap" else io_formats.read_hdf5(out, "distances")
    assert got.dtype == np.float32
    want = _oracle(seqs, "110011", "plus", metric, np.float32)
    assert got.shape == want.shape
    m = ~np.isnan(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    # Eucl in the --large modes is the Gram form on the tensor cores (stated tolerance 1e-4)
    assert np.allclose(got[m], want[m], rtol=1e-4 if metric == "Eucl" else 1e-6, atol=1e-7)
    # the reference's own cross-mode acceptance (bin/phyloligo_comparemat.py:44)
    assert np.allclose(np.nan_to_num(got), np.nan_to_num(want), atol=1e-3)
    # float32 frequency file of the --large modes; the temp dir is gone afterwards
    F = np.loadtxt(freq)
    assert np.array_equal(F.astype(np.float32),
                          np.vstack([po.frequency_np(s, "110011", "plus") for s in seqs]).astype(np.float32))
    assert not [d for d in os.listdir(tmp_path) if d.startswith("tmp")]


def test_bad_strand_exits_like_the_reference(fasta, capsys):
    path, _ = fasta
    with pytest.raises(SystemExit) as e:
        phyloligo.compute_frequency("ACGT", "1111", "sideways")
    assert e.value.code == 1
    assert "strand parameter" in capsys.readouterr().err


def test_command_line_against_the_unmodified_reference_script(tmp_path):
    """Every mode of the reference's own bin/phyloligo.py that runs in the build container (eleven runs through
    oracle/run_reference_cli.py, committed as tests/golden/cli_golden.npz) against this command line on the
    same assembly with the same arguments: the -q frequency files bit for bit, the matrices within the
    tolerances of BASELINE.json (1e-6; 1e-4 for the tensor-core Eucl of the --large modes)."""
    import sys
    from conftest import GOLDEN
    sys.path.insert(0, GOLDEN)
    import make_cli_golden as mk
    golden = dict(np.load(os.path.join(GOLDEN, "cli_golden.npz")))
    fasta, n = mk.assembly()
    path = os.path.join(tmp_path, "asm.fasta")
    open(path, "wb").write(fasta)
    for name in golden["names"]:
        args = str(golden[name + "_args"]).split()
        large = args[args.index("--large") + 1] if "--large" in args else "None"
        metric = args[args.index("-d") + 1]
        out, freq = os.path.join(tmp_path, name + ".mat"), os.path.join(tmp_path, name + ".freq")
        phyloligo.main(["-i", path, "-o", out, "-q", freq, "-w", str(tmp_path)] + args)
        ref_M, ref_F = golden[name + "_matrix"], golden[name + "_freq"]
        F = np.loadtxt(freq)
        assert np.array_equal(F, ref_F), name  # the text the reference wrote, value for value
        got = np.loadtxt(out) if large == "None" else io_formats.read_memmap(out)
        assert got.shape == ref_M.shape and got.dtype == ref_M.dtype, name
        assert np.array_equal(np.isnan(got), np.isnan(ref_M)), name
        m = ~np.isnan(ref_M)
        if metric == "KT":
            tol, atol = 1e-12, 1e-12
        elif large != "None":
            # both sides compute in float32 here: the reference's own rounding (sklearn's float32 Gram form,
            # the broadcast float32 JSD) is part of the difference
            tol, atol = (1e-4, 1e-6) if metric == "Eucl" else (5e-6, 1e-7)
        else:
            tol, atol = 1e-6, 1e-12
        assert np.allclose(got[m], ref_M[m], rtol=tol, atol=atol), (name, np.abs(got[m] - ref_M[m]).max())
        # the reference's own acceptance threshold between its modes (bin/phyloligo_comparemat.py:44)
        assert np.allclose(np.nan_to_num(got), np.nan_to_num(ref_M), atol=1e-3), name
