#!/usr/bin/env python3
"""Generate tests/golden/kount_golden.json by running the REFERENCE's own Kount.py function
bodies (AST-extracted from /root/reference/phylopackage/bin/Kount.py; nothing is copied).

    python tests/golden/make_kount_golden.py

Bio.SeqIO / Bio.Seq are third-party and absent here: SeqIO.parse is replaced by the oracle's FASTA
reader (records with .id / .seq) and Seq.reverse_complement by the oracle's reverse complement.
Everything else -- window cutting and coordinates (make_genome_chunk), the N filter, counting,
count2freq, the whole-genome composition, KL / Eucl / JSD -- is the reference code itself.
"""
import ast
import json
import os
import re
import sys
from collections import Counter
from itertools import product

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import kount_oracle as ko  # noqa: E402
from oracle import phylo_oracle as po  # noqa: E402
from phyloligo_b200 import synth  # noqa: E402

KOUNT = os.path.join(os.environ.get("PHYLOLIGO_REFERENCE", "/root/reference"), "phylopackage", "bin", "Kount.py")
NAMES = ["posdef_check_value", "KL", "Eucl", "JSD", "select_strand", "cut_sequence_and_count_pattern", "count2freq",
         "compute_frequency", "compute_distance_joblib", "make_genome_chunk"]


class _Seq:
    def __init__(self, s):
        self.s = str(s)

    def reverse_complement(self):
        return _Seq(po.reverse_complement(self.s))

    def __str__(self):
        return self.s


class _Record:
    def __init__(self, name, seq):
        self.id, self.seq = name, _Seq(seq)


class _SeqIO:
    @staticmethod
    def parse(path, fmt):
        assert fmt == "fasta"
        for name, seq in ko.read_records(path):
            yield _Record(name, seq)


def load_reference():
    tree = ast.parse(open(KOUNT).read(), filename=KOUNT)
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in NAMES]
    assert {n.name for n in wanted} == set(NAMES)
    ns = {"re": re, "np": np, "Counter": Counter, "product": product, "sys": sys, "Seq": _Seq, "SeqIO": _SeqIO,
          "min_nb_w_per_fasta_for_mul_cpu": 20}
    np.seterr(divide="ignore", invalid="ignore")  # Kount.py:59
    exec(compile(ast.Module(body=wanted, type_ignores=[]), KOUNT, "exec"), ns)
    return ns


def whole_composition(ns, path, pattern, strand):
    """compute_whole_composition (Kount.py:303-314) without joblib: the same three statements"""
    counts = [ns["cut_sequence_and_count_pattern"](str(r.seq), pattern, strand) for r in _SeqIO.parse(path, "fasta")]
    total = Counter()
    for c in counts:
        for word in c:
            total[word] += c[word]
    return ns["count2freq"](total, pattern.count("1"))


def assembly(seed, lengths, line):
    rng = np.random.default_rng(seed)
    recs = []
    for i, n in enumerate(lengths):
        s = synth._bases(n, float(rng.uniform(0.35, 0.65)), rng) if n else np.zeros(0, np.uint8)
        if n >= 200:
            for _ in range(int(rng.integers(1, 4))):
                p = int(rng.integers(0, n - 150))
                s[p:p + int(rng.integers(5, 150))] = ord("N")
            p = int(rng.integers(0, n - 60))
            s[p:p + 50] |= 0x20  # lower case, including n (not counted by the N filter)
        recs.append((("ctg%d some description" % i), s.tobytes().decode()))
    text = "".join(">%s\n%s\n" % (h, "\n".join(s[k:k + line] for k in range(0, len(s), line))) for h, s in recs)
    return text


def main():
    ns = load_reference()
    cases = []
    specs = [
        # (seed, lengths, line width, window, step, pattern, strand, n_max)
        (1, [120, 300, 301, 650, 999, 1000, 1001, 2500, 40], 70, 300, 50, "1111", "both", 0.4),
        (2, [700, 1500, 3100], 60, 300, 50, "11", "plus", 0.1),
        (3, [4000, 12000, 30500, 9999, 10000], 80, 5000, 500, "1111", "both", 0.4),
        (4, [900, 2600], 61, 250, 100, "1111", "minus", 0.05),
    ]
    tmp = os.path.join(HERE, "_kount_tmp.fasta")
    for seed, lengths, line, w, t, pattern, strand, n_max in specs:
        text = assembly(seed, lengths, line)
        with open(tmp, "w") as fh:
            fh.write(text)
        mcp = whole_composition(ns, tmp, pattern, strand)
        case = {"fasta": text, "window": w, "step": t, "pattern": pattern, "strand": strand, "n_max": n_max,
                "mcp": [float(v).hex() for v in np.asarray(mcp, dtype=np.float64)], "rows": {}}
        for dist in ("JSD", "KL", "Eucl"):
            rows = []
            for info, seqs in ns["make_genome_chunk"](tmp, w, t, None, 50000):
                for (sid, a, b), s in zip(info, seqs):
                    d = ns["compute_distance_joblib"](dist, mcp, s, pattern, strand, n_max)
                    rows.append([sid, int(a), int(b), float(d).hex()])
            case["rows"][dist] = rows
        cases.append(case)
    os.unlink(tmp)
    with open(os.path.join(HERE, "kount_golden.json"), "w") as fh:
        json.dump(cases, fh)
    print("wrote %d cases, %d rows" % (len(cases), sum(len(c["rows"]["JSD"]) for c in cases)))


if __name__ == "__main__":
    main()
