"""CPU tests: the C restatement (oracle/oracle.c) against the Python oracle, which is
pinned to the reference's own outputs."""
import numpy as np

from conftest import golden_case_arrays
from oracle import coracle, phylo_oracle as po
from phyloligo_b200 import engine, synth


def test_c_counts_match_reference_golden(profile_golden):
    seqs = profile_golden["sequences"]
    for case in profile_golden["cases"][::3]:
        counts, total, _ = golden_case_arrays(case)
        c, t = coracle.count(seqs[case["seq"]].encode("latin-1"), case["pattern"], case["strand"])
        assert t == total and np.array_equal(c, counts)


def test_c_profile_batch_and_distances():
    seqs = synth.make_sequences(30, 1500, seed=14)
    text, begin, end = engine.sequences_to_text(seqs)
    F = coracle.profile_batch(text, begin, end, "1111", "both", threads=3)
    for r, s in enumerate(seqs):
        assert np.array_equal(F[r], po.frequency_np(s, "1111", "both"))
    F[4] = 0
    for metric in ("Eucl", "JSD", "BC", "KT", "SC"):
        M = coracle.pairwise_rows(metric, F[:12], threads=4)
        fn = po.METRICS[metric]
        for i in range(12):
            for j in range(12):
                w = fn(F[i], F[j])
                if np.isnan(w):
                    assert np.isnan(M[i, j])
                else:
                    assert abs(M[i, j] - w) <= 1e-13 * max(1.0, abs(w)), (metric, i, j)
