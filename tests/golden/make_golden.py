#!/usr/bin/env python3
"""Generate the committed golden vectors by running the REFERENCE's own function
bodies (AST-extracted from /root/reference, see oracle/ref_extract.py).

Run once in the build container (where /root/reference is mounted):

    python tests/golden/make_golden.py

Outputs (committed):
    tests/golden/profile_golden.json   sequences, pattern, strand -> counts/total/freq (hex float64)
    tests/golden/distance_golden.npz   profile matrices -> Eucl / JSD (1-D f64 form, 2-D f32 form)

Strand preparation uses the oracle's reverse complement (Bio.Seq is third-party
and absent here); everything after it -- window cutting, counting, count2freq,
KL/Eucl/JSD -- is the reference code itself.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import phylo_oracle as po  # noqa: E402
from oracle import ref_extract  # noqa: E402
from phyloligo_b200 import synth  # noqa: E402


def profile_cases():
    rng = np.random.default_rng(20261018)
    seqs = [
        "ACGTNACGTA",
        "AACG",
        "",
        "N",
        "ACG",
        "acgtacgtnnACGTTGCA",
        "ACGTRYKMACGTSWBDHVACGTTTGA",
        "A" * 40,
        "ACGT" * 30 + "NNNN" + "TTGACCA" * 9,
        "GATTACA",
        "ACGTUACGT-ACGT*ACGT",
    ]
    for L in (5, 9, 17, 33, 64, 150, 333, 1000, 2500):
        s = synth._bases(L, float(rng.uniform(0.3, 0.7)), rng)
        if L >= 64:
            p = int(rng.integers(0, L - 12))
            s[p:p + int(rng.integers(1, 12))] = ord("N")
            p = int(rng.integers(0, L - 30))
            s[p:p + 25] |= 0x20
        seqs.append(s.tobytes().decode())
    patterns = ["1", "11", "1111", "11111", "101", "11001", "110101", "111010011", "1000000001", "100", "0110"]
    cases = []
    for si, s in enumerate(seqs):
        for pat in patterns:
            if len(s) > 400 and pat in ("1", "100", "0110"):
                continue
            for strand in ("plus", "minus", "both"):
                cases.append((si, pat, strand))
    return seqs, cases


def main():
    ref = ref_extract.load()
    seqs, cases = profile_cases()
    out_cases = []
    for si, pat, strand in cases:
        prepared = po.select_strand(seqs[si], strand).upper()
        words, total = ref["cut_sequence_and_count_pattern"](prepared, pat)
        k = pat.count("1")
        freq = ref["count2freq"](words, total, k)
        counts = np.zeros(4 ** k, dtype=np.int64)
        for w, c in words.items():
            counts[po.word_index(w)] = c
        nz = np.nonzero(counts)[0]
        out_cases.append({
            "seq": si, "pattern": pat, "strand": strand, "total": int(total),
            "nz_index": nz.tolist(), "nz_count": counts[nz].tolist(),
            # float64 frequencies of the non-zero bins, exact hex form
            "nz_freq_hex": [float(freq[i]).hex() for i in nz],
            "freq_dtype": str(freq.dtype),
        })
    with open(os.path.join(HERE, "profile_golden.json"), "w") as fh:
        json.dump({"sequences": seqs, "cases": out_cases}, fh)

    # ---- distances -------------------------------------------------------
    rng = np.random.default_rng(7)
    mats = {}
    # realistic profiles: counts from synthetic contigs via the reference bodies
    contigs = synth.make_sequences(24, 3000, seed=11)
    rows = []
    for s in contigs:
        prepared = po.select_strand(s.decode(), "both").upper()
        words, total = ref["cut_sequence_and_count_pattern"](prepared, "1111")
        rows.append(ref["count2freq"](words, total, 4).astype(np.float64))
    mats["real_k4"] = np.vstack(rows)
    # sparse profiles with many exact zeros and one all-zero row
    X = rng.random((16, 64))
    X[rng.random(X.shape) < 0.4] = 0.0
    X[3] = 0.0
    X[5] = X[4]  # duplicate rows -> distance exactly 0
    s = X.sum(axis=1, keepdims=True)
    s[s == 0] = 1.0
    mats["sparse_64"] = X / s
    # one-hot rows: JSD = ln 2 between different bins
    mats["onehot_16"] = np.eye(16)[:6]
    out = {}
    for name, X in mats.items():
        n = X.shape[0]
        e = np.zeros((n, n))
        j = np.zeros((n, n))
        for a in range(n):
            for b in range(n):
                e[a, b] = ref["Eucl"](X[a].copy(), X[b].copy())
                j[a, b] = ref["JSD"](X[a].copy(), X[b].copy())
        X32 = X.astype(np.float32)
        j32 = ref["JSD"](X32.copy(), X32[: max(2, n // 2)].copy())  # 2-D form, rows index second arg
        out[name + "_X"] = X
        out[name + "_Eucl"] = e
        out[name + "_JSD"] = j
        out[name + "_JSD2d_f32"] = np.asarray(j32)
    np.savez_compressed(os.path.join(HERE, "distance_golden.npz"), **out)
    print("wrote", len(out_cases), "profile cases and", len(mats), "distance matrices")


if __name__ == "__main__":
    main()
