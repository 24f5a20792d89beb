"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, never the product path.

A plain Python / numpy restatement of the two PhylOligo stages that
``phyloligo_b200`` accelerates.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import
this module.  The shipped package never does (it fails loudly when the CUDA
library is missing instead of falling back to this code).

Pinning status
--------------
* profiling (``cut_sequence_and_count_pattern``, ``count2freq``), ``Eucl``,
  ``KL`` and ``JSD``: PINNED.  ``tests/golden/make_golden.py`` executed the
  reference's own function bodies (AST-extracted from
  ``/root/reference/phylopackage``) on seeded inputs and committed the outputs
  under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them here.
* ``BC``: PINNED against scipy's ``braycurtis`` (the routine the reference
  dispatches to, ``core/phylodist.py:79`` / ``bin/phyloligo.py:381``).
* ``SC``: the reference body raises ``NameError`` (``core/phylodist.py:82-85``
  uses ``spearmanr`` without importing it).  Restated as intended and PINNED
  against ``scipy.stats.spearmanr``.
* ``KT``: PARITY UNPINNED against the real Biopython C library
  (``Bio.Cluster`` is not installable here).  Restated from the published
  C Clustering Library algorithm ``kendall()`` (Biopython >= 1.68,
  ``Bio/Cluster/cluster.c``) and cross-checked against
  ``scipy.stats.kendalltau`` (tau-b) wherever both are defined.
* reverse complement / FASTA reading follow Biopython's documented behaviour
  (``Bio.Seq.reverse_complement`` with the ambiguous-DNA table;
  ``SeqIO.parse(..., "fasta")``), third-party code absent from the checkout.

All ``file:line`` citations point into ``/root/reference/phylopackage``.
"""
from __future__ import annotations

import math
import re
from collections import Counter
from itertools import product

import numpy as np

ALPHABET = ("C", "G", "A", "T")  # bin/phyloligo.py:653 -- word order of the frequency vector

# Biopython ambiguous_dna_complement (Bio/Data/IUPACData.py), both cases.
_COMP_SRC = "ACGTMRWSYKVHDBNacgtmrwsykvhdbn"
_COMP_DST = "TGCAKYWSRMBDHVNtgcakywsrmbdhvn"
_COMP_TABLE = str.maketrans(_COMP_SRC, _COMP_DST)


# ----------------------------------------------------------------------------
# FASTA (Bio.SeqIO.parse(..., "fasta") call sites bin/phyloligo.py:87,114,154,869,914,959)
# ----------------------------------------------------------------------------
def read_fasta(path):
    """Yield the sequence string of each record, Biopython-style.

    Lines before the first '>' are ignored; sequence lines are stripped of
    surrounding whitespace and concatenated, inner blanks removed.
    """
    seq_parts = None
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith(">"):
                if seq_parts is not None:
                    yield "".join(seq_parts)
                seq_parts = []
            elif seq_parts is not None:
                seq_parts.append("".join(line.split()))
    if seq_parts is not None:
        yield "".join(seq_parts)


# ----------------------------------------------------------------------------
# Profiling
# ----------------------------------------------------------------------------
def reverse_complement(seq: str) -> str:
    """Bio.Seq.reverse_complement as used at bin/phyloligo.py:141,143."""
    return seq.translate(_COMP_TABLE)[::-1]


def select_strand(seq: str, strand: str = "both") -> str:
    """bin/phyloligo.py:124-149.  'both' is seq + revcomp(seq), no separator."""
    if strand == "both":
        return seq + reverse_complement(seq)
    if strand == "minus":
        return reverse_complement(seq)
    if strand == "plus":
        return seq
    raise ValueError("strand must be one of 'both', 'minus', 'plus'")


def cut_sequence_and_count_pattern(seq: str, pattern: str):
    """bin/phyloligo.py:601-631, literal restatement.

    Split on runs of non-ACGT, slide a len(pattern) window over every run that
    is long enough, keep the characters under the '1's.
    """
    pattern = str(pattern)
    ones = [i for i, c in enumerate(pattern) if c == "1"]
    width = len(pattern)
    words = Counter()
    for run in re.split("[^ACGT]+", seq):
        if len(run) >= width:
            for start in range(len(run) - (width - 1)):
                words["".join(run[start + o] for o in ones)] += 1
    return words, sum(words.values())


def count2freq(count_words, kword_count, ksize):
    """bin/phyloligo.py:633-661.  C,G,A,T product order; int/int true division."""
    if kword_count > 0:
        feats = []
        for letters in product(ALPHABET, repeat=ksize):
            w = "".join(letters)
            feats.append(count_words[w] / kword_count if w in count_words else 0)
    else:
        feats = [0 for _ in range(4 ** ksize)]
    return np.array(feats)


def compute_frequency(seq: str, pattern="1111", strand="both"):
    """bin/phyloligo.py:663-691."""
    pattern = str(pattern)
    s = select_strand(seq, strand).upper()
    words, total = cut_sequence_and_count_pattern(s, pattern)
    return count2freq(words, total, pattern.count("1"))


def word_index(word: str) -> int:
    """Position of `word` in the C,G,A,T product order (bin/phyloligo.py:653)."""
    idx = 0
    for ch in word:
        idx = idx * 4 + ALPHABET.index(ch)
    return idx


def count_vector(seq: str, pattern="1111", strand="both"):
    """Integer counts in frequency-vector order plus the total (literal path)."""
    pattern = str(pattern)
    s = select_strand(seq, strand).upper()
    words, total = cut_sequence_and_count_pattern(s, pattern)
    out = np.zeros(4 ** pattern.count("1"), dtype=np.int64)
    for w, c in words.items():
        out[word_index(w)] = c
    return out, total


# numpy-vectorised restatement of the same counting, for inputs too large for
# the literal loop.  tests/test_oracle_golden.py checks it against the literal one.
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _ch in enumerate("CGAT"):
    _CODE[ord(_ch)] = _i
    _CODE[ord(_ch.lower())] = _i  # .upper() at bin/phyloligo.py:683


def _count_codes(codes: np.ndarray, pattern: str) -> np.ndarray:
    ones = [i for i, c in enumerate(pattern) if c == "1"]
    width, k = len(pattern), len(ones)
    dim = 4 ** k
    n = codes.shape[0] - width + 1
    if n <= 0:
        return np.zeros(dim, dtype=np.int64)
    bad = (codes == 255).astype(np.int64)
    csum = np.concatenate(([0], np.cumsum(bad)))
    ok = (csum[width:width + n] - csum[:n]) == 0
    word = np.zeros(n, dtype=np.int64)
    for o in ones:
        word = word * 4 + (codes[o:o + n] & 3)
    return np.bincount(word[ok], minlength=dim).astype(np.int64)


def count_vector_np(seq, pattern="1111", strand="both"):
    """Same result as count_vector, vectorised.  `seq` is str or bytes."""
    pattern = str(pattern)
    raw = np.frombuffer(seq.encode("latin-1") if isinstance(seq, str) else bytes(seq), dtype=np.uint8)
    codes = _CODE[raw]
    if strand == "plus":
        full = codes
    else:
        rc = codes[::-1].copy()
        good = rc != 255
        rc[good] ^= 1  # C<->G, A<->T in the C,G,A,T code
        full = rc if strand == "minus" else np.concatenate((codes, rc))
        if strand not in ("minus", "both"):
            raise ValueError("strand must be one of 'both', 'minus', 'plus'")
    counts = _count_codes(full, pattern)
    return counts, int(counts.sum())


def frequency_np(seq, pattern="1111", strand="both", dtype=np.float64):
    """float64 = count/total correctly rounded (bin/phyloligo.py:656); float32 is
    the cast of that float64 quotient (bin/phyloligo.py:720,777-786)."""
    counts, total = count_vector_np(seq, pattern, strand)
    if total == 0:
        return np.zeros(counts.shape[0], dtype=dtype)
    return (counts.astype(np.float64) / np.float64(total)).astype(dtype)


# ----------------------------------------------------------------------------
# Distances (core/phylodist.py)
# ----------------------------------------------------------------------------
def _scrub(d):
    """posdef_check_value, core/phylodist.py:12-14."""
    d[np.isnan(d)] = 0
    d[np.isinf(d)] = 0


def KL(a, b):
    """core/phylodist.py:18-24 (1-D branch): sum a*ln(a/b), NaN/Inf terms -> 0."""
    with np.errstate(divide="ignore", invalid="ignore"):
        d = a * np.log(a / b)
    _scrub(d)
    return np.sum(d)


def Eucl(a, b):
    """core/phylodist.py:36-41."""
    d = pow(a - b, 2)
    _scrub(d)
    return np.sqrt(np.sum(d))


def JSD(a, b):
    """core/phylodist.py:43-68.  1-D pair form and the 2-D x 2-D broadcast form
    (rows of the result index `b`)."""
    if a.ndim == 1 and b.ndim == 1:
        h = 0.5 * (a + b)
        return 0.5 * (KL(a, h) + KL(b, h))
    if a.ndim == 2 and b.ndim == 2:
        with np.errstate(divide="ignore", invalid="ignore"):
            h = 0.5 * (a[np.newaxis, :] + b[:, np.newaxis])
            d1 = a[np.newaxis, :] * np.log(a[np.newaxis, :] / h)
            _scrub(d1)
            d1 = np.sum(d1, axis=2)
            d2 = b[:, np.newaxis] * np.log(b[:, np.newaxis] / h)
            _scrub(d2)
            d2 = np.sum(d2, axis=2)
        return 0.5 * (d1 + d2)
    raise ValueError("JSD oracle handles 1-D x 1-D and 2-D x 2-D")


def BC(a, b):
    """Bray-Curtis as scipy computes it (reached from core/phylodist.py:79 and
    bin/phyloligo.py:381): sum|a-b| / sum|a+b|."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.abs(a - b).sum() / np.abs(a + b).sum()


def kendall_distance(a, b):
    """C Clustering Library kendall() (Biopython Bio/Cluster/cluster.c): the
    'k' distance = 1 - tau_b, with 1.0 when either variable is constant and
    0.0 when there is no element pair at all."""
    n = len(a)
    con = dis = exx = exy = 0
    flag = False
    for i in range(n):
        for j in range(i):
            x1, x2, y1, y2 = a[i], a[j], b[i], b[j]
            if x1 < x2 and y1 < y2:
                con += 1
            if x1 > x2 and y1 > y2:
                con += 1
            if x1 < x2 and y1 > y2:
                dis += 1
            if x1 > x2 and y1 < y2:
                dis += 1
            if x1 == x2 and y1 != y2:
                exx += 1
            if x1 != x2 and y1 == y2:
                exy += 1
            flag = True
    if not flag:
        return 0.0
    denomx = con + dis + exx
    denomy = con + dis + exy
    if denomx == 0 or denomy == 0:
        return 1.0
    tau = (con - dis) / math.sqrt(float(denomx) * float(denomy))
    return 1.0 - tau


def kendall_counts_np(a, b):
    """(con - dis, pairs with a untied, pairs with b untied), vectorised."""
    a = np.asarray(a)
    b = np.asarray(b)
    iu = np.triu_indices(len(a), 1)
    sa = np.sign(a[iu[0]] - a[iu[1]]).astype(np.int64)
    sb = np.sign(b[iu[0]] - b[iu[1]]).astype(np.int64)
    # con+dis+exx = pairs where b is untied; con+dis+exy = pairs where a is untied
    return int((sa * sb).sum()), int((sb != 0).sum()), int((sa != 0).sum())


def KT(a, b):
    """core/phylodist.py:71-74: 1 - distancematrix((a,b), dist='k')[1][0],
    i.e. Kendall's tau_b itself (0 when either row is constant)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if len(a) < 2:
        return 1.0
    s, denomx, denomy = kendall_counts_np(a, b)
    if denomx == 0 or denomy == 0:
        return 0.0
    return 1.0 - (1.0 - s / math.sqrt(float(denomx) * float(denomy)))


def rank_average(a):
    """scipy.stats.rankdata(method='average') restated: 1-based, ties share the mean rank."""
    a = np.asarray(a)
    order = np.argsort(a, kind="stable")
    sa = a[order]
    ranks = np.empty(len(a), dtype=np.float64)
    i = 0
    n = len(a)
    while i < n:
        j = i
        while j + 1 < n and sa[j + 1] == sa[i]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def SC(a, b):
    """Intended meaning of core/phylodist.py:82-85: 1 - spearmanr(a, b).correlation
    = 1 - Pearson(average ranks).  NaN when either row is constant (scipy)."""
    ra = rank_average(a)
    rb = rank_average(b)
    ra = ra - ra.mean()
    rb = rb - rb.mean()
    den = math.sqrt(float((ra * ra).sum()) * float((rb * rb).sum()))
    if den == 0.0:
        return float("nan")
    return 1.0 - float((ra * rb).sum()) / den


METRICS = {"Eucl": Eucl, "JSD": JSD, "KT": KT, "BC": BC, "SC": SC}


def pairwise(X, metric):
    """Full N x N matrix via the pair functions (what sklearn.pairwise_distances
    with a callable does at bin/phyloligo.py:388-390: upper triangle, mirrored,
    diagonal from metric(x, x))."""
    X = np.asarray(X, dtype=np.float64)
    fn = METRICS[metric]
    n = X.shape[0]
    out = np.zeros((n, n), dtype=np.float64)
    for i in range(n):
        for j in range(i, n):
            out[i, j] = out[j, i] = fn(X[i], X[j])
    return out


def pairwise_np(X, metric, block=256):
    """Vectorised float64 N x N matrix for Eucl / JSD / BC (same formulas,
    fp64 accumulation) -- for oracle runs at sizes the pair loop cannot reach."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    out = np.empty((n, n), dtype=np.float64)
    for r0 in range(0, n, block):
        A = X[r0:r0 + block][:, None, :]
        for c0 in range(0, n, block):
            B = X[c0:c0 + block][None, :, :]
            if metric == "Eucl":
                out[r0:r0 + block, c0:c0 + block] = np.sqrt(((A - B) ** 2).sum(axis=2))
            elif metric == "BC":
                with np.errstate(divide="ignore", invalid="ignore"):
                    out[r0:r0 + block, c0:c0 + block] = np.abs(A - B).sum(axis=2) / np.abs(A + B).sum(axis=2)
            elif metric == "JSD":
                with np.errstate(divide="ignore", invalid="ignore"):
                    H = 0.5 * (A + B)
                    d1 = A * np.log(A / H)
                    d2 = B * np.log(B / H)
                d1[~np.isfinite(d1)] = 0
                d2[~np.isfinite(d2)] = 0
                out[r0:r0 + block, c0:c0 + block] = 0.5 * (d1.sum(axis=2) + d2.sum(axis=2))
            else:
                raise ValueError(metric)
    return out
