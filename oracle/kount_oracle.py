"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY, never the product path.

Plain Python / numpy restatement of the sliding-window stage of the reference's
``phylopackage/bin/Kount.py`` (windows of every contig -> composition profile ->
distance to one reference profile).  Only ``tests/`` may import this module.

Pinning status: PINNED.  ``tests/golden/make_kount_golden.py`` executed the
reference's own function bodies (AST-extracted from ``bin/Kount.py``:
``make_genome_chunk``, ``compute_frequency``, ``count2freq``, ``KL``, ``Eucl``,
``JSD``, ``compute_distance_joblib``, with ``SeqIO.parse`` / ``Seq`` replaced by
the oracle's FASTA reader and reverse complement -- Biopython is absent here) on
seeded assemblies and committed the outputs (``tests/golden/kount_golden.json``);
``tests/test_oracle_golden.py`` replays them here.

All ``file:line`` citations point into ``/root/reference/phylopackage/bin/Kount.py``.
"""
from __future__ import annotations

import numpy as np

from . import phylo_oracle as po

MIN_NB_W_PER_FASTA_FOR_MUL_CPU = 20  # :64


def read_records(path):
    """(id, sequence) of every record: SeqIO.parse(genome, "fasta") with record.id = the first
    blank-delimited token of the header (call sites :305, :348, :476)."""
    out = []
    name, parts = None, None
    with open(path, "r") as fh:
        for line in fh:
            if line.startswith(">"):
                if parts is not None:
                    out.append((name, "".join(parts)))
                fields = line[1:].split()
                name, parts = (fields[0] if fields else ""), []
            elif parts is not None:
                parts.append("".join(line.split()))
    if parts is not None:
        out.append((name, "".join(parts)))
    return out


def _scrub(d):
    """posdef_check_value :67-69"""
    d = np.array(d, dtype=np.float64)
    d[np.isnan(d)] = 0
    d[np.isinf(d)] = 0
    return d


def KL(a, b):
    """1-D branch of KL :71-86 (natural log, NaN / Inf terms zeroed)"""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(np.sum(_scrub(a * np.log(a / b))))


def Eucl(a, b):
    """:88-92, scaled by 1000"""
    with np.errstate(invalid="ignore"):
        return float(np.sqrt(np.sum(_scrub(np.power(a - b, 2)))) * 1000)


def JSD(a, b):
    """1-D branch of JSD :94-123, scaled by 1000"""
    h = 0.5 * (a + b)
    return 0.5 * (KL(a, h) + KL(b, h)) * 1000


def count2freq(counts, total, dim):
    """:230-272 -- count / sum(counts) in product(("C","G","A","T")) order, zeros when nothing counted"""
    if total > 0:
        return np.asarray(counts, dtype=np.float64) / float(total)
    return np.zeros(dim, dtype=np.float64)


def compute_frequency(seq, n_max_freq_in_windows=1.0, pattern="1111", strand="both"):
    """:274-301.  A window with more than n_max upper-case 'N' gets an all-NaN vector (whose
    length the reference writes as ksize**4, a crash for k not in {2, 4}; here 4**ksize)."""
    pattern = str(pattern)
    dim = 4 ** pattern.count("1")
    if (seq.count("N") / len(seq)) <= float(n_max_freq_in_windows):
        counts, total = po.count_vector_np(seq, pattern, strand)
        return count2freq(counts, total, dim)
    return np.full(dim, np.nan)


def compute_distance(mth_dist, mcp, seq, pattern, strand, n_max_freq_in_windows):
    """compute_distance_joblib :317-324"""
    freq = compute_frequency(seq, n_max_freq_in_windows, pattern, strand)
    if mth_dist == "JSD":
        return JSD(freq, mcp)
    if mth_dist == "KL":
        return KL(freq, mcp)
    return Eucl(freq, mcp)


def compute_whole_composition(records, pattern, strand):
    """:303-314 -- counts of every record summed, then count2freq"""
    pattern = str(pattern)
    dim = 4 ** pattern.count("1")
    counts = np.zeros(dim, dtype=np.int64)
    total = 0
    for _, seq in records:
        c, t = po.count_vector_np(seq, pattern, strand)
        counts += c
        total += t
    return count2freq(counts, total, dim)


def make_windows(records, windows_size, windows_step):
    """make_genome_chunk :343-407 without the chunking: a list of
    (seq_id, displayed_start, displayed_stop, window_string)."""
    out = []
    w, t = windows_size, windows_step
    for seq_id, seq in records:
        n = len(seq)
        if n < w:  # :351-353
            out.append((seq_id, 0, int(n), seq))
        elif n < MIN_NB_W_PER_FASTA_FOR_MUL_CPU * t:  # :359-382
            for s in range(0, n - w, t):
                start = 1 if s == 0 else int(s + w / 2 - t / 2)
                stop = n if s == n - w else int(s + w / 2 + t / 2)
                out.append((seq_id, start, stop, seq[s:s + w]))
        else:  # :388-403
            for s in range(0, n - w, t):
                start, stop = int(s + w / 2 - t / 2), int(s + w / 2 + t / 2)
                dstart = 1 if start == (w / 2 - t / 2) else start
                edge = stop - t / 2 + w / 2
                dstop = n if (edge >= n - t and edge <= n) else stop
                out.append((seq_id, dstart, dstop, seq[s:s + w]))
    return out


def sliding_windows_distances(records, mcp, mth_dist, pattern, windows_size, windows_step, strand, n_max):
    """:409-453 -- rows [seq_id, displayed_start, displayed_stop, distance]"""
    rows = []
    for seq_id, start, stop, window in make_windows(records, windows_size, windows_step):
        rows.append([seq_id, start, stop, compute_distance(mth_dist, mcp, window, pattern, strand, n_max)])
    return rows
