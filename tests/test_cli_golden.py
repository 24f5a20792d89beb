"""The oracle against the end-to-end fixtures written by the reference's UNMODIFIED command-line script
(tests/golden/make_cli_golden.py -> cli_golden.npz: eleven runs of bin/phyloligo.py through
oracle/run_reference_cli.py).  CPU only; the GPU counterpart is tests/test_gpu_cli.py."""
import hashlib
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import phylo_oracle as po

sys.path.insert(0, GOLDEN)
import make_cli_golden as mk  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    return dict(np.load(os.path.join(GOLDEN, "cli_golden.npz")))


def parse_args(text):
    a = text.split()
    get = lambda f, d=None: a[a.index(f) + 1] if f in a else d  # noqa: E731
    pattern = get("-p") or "1" * int(get("-k", "4"))
    return pattern, get("-s", "both"), get("-d", "Eucl"), get("--method"), get("--large", "None")


def sequences():
    fasta, n = mk.assembly()
    recs = fasta.decode().split(">")[1:]
    return fasta, ["".join(r.split("\n")[1:]) for r in recs]


def test_assembly_is_the_one_the_goldens_were_made_from(golden):
    fasta, n = mk.assembly()
    assert hashlib.sha256(fasta).hexdigest() == str(golden["fasta_sha256"]) and n == int(golden["n_records"])
    assert len(golden["names"]) == 11


def test_oracle_reproduces_the_reference_command_line(golden):
    _, seqs = sequences()
    for name in golden["names"]:
        pattern, strand, metric, method, large = parse_args(str(golden[name + "_args"]))
        F = np.vstack([po.frequency_np(s, pattern, strand) for s in seqs])
        ref_F, ref_M = golden[name + "_freq"], golden[name + "_matrix"]
        if large == "None":
            assert np.array_equal(F, ref_F), name                       # '%.18e' text round-trips float64 exactly
            X = F
        else:
            assert np.array_equal(F.astype(np.float32), ref_F.astype(np.float32)), name  # float32 memmap of the quotient
            X = F.astype(np.float32).astype(np.float64)
        if metric in ("Eucl", "JSD", "BC"):
            want = po.pairwise_np(X, metric)
        else:
            want = np.array([[po.KT(a, b) for b in X] for a in X])
        assert ref_M.shape == want.shape
        assert np.array_equal(np.isnan(ref_M), np.isnan(want)), name
        m = ~np.isnan(want)
        # float64 modes: summation order only; the --large modes compute in float32 (sklearn's Gram form, the
        # broadcast JSD of core/phylodist.py:58-66)
        tol = 1e-10 if large == "None" else 2e-5
        assert np.allclose(ref_M[m], want[m], rtol=tol, atol=1e-7 if large != "None" else 1e-12), (name, np.abs(ref_M[m] - want[m]).max())
