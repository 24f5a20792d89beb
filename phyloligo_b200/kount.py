#!/usr/bin/env python3
"""Kount.py on B200: sliding-window composition distances (reference
phylopackage/bin/Kount.py), same command line, same output files.

Every window of every contig (``make_genome_chunk``, reference :343-407) is a
virtual record of the profiling kernel: the assembly is loaded once, its sequences
are laid out in HBM without line breaks, and a batch of windows is three launches --
``po_profile_batch`` (the window profiles), ``po_window_count_byte`` (the 'N'
filter of ``compute_frequency``, :294) and ``po_window_distances`` (KL / Eucl / JSD
to the reference profile, ``compute_distance_joblib`` :317-324).  The whole-genome /
host / contaminant profile (``compute_whole_composition``, :303-314) is one
``po_profile_batch`` over the records with the counts summed.

Worker-level names of the reference are kept: ``KL``, ``Eucl``, ``JSD``,
``compute_frequency``, ``compute_whole_composition``, ``compute_distance_joblib``,
``make_genome_chunk``, ``sliding_windows_distances``, ``get_cmd``, ``main``.

Deliberate handling of reference defects: a window over the 'N' limit gets an
all-NaN profile in the reference, whose NaN terms are then zeroed -- its distance is
0.0 for every metric (and a shape error for k other than 2 and 4, ``ksize**4``, :300);
here it is 0.0 for every k.  An empty record divides by zero in the reference (:294);
here it is profiled as an all-zero vector.
"""
from __future__ import annotations

import argparse
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

from . import _lib, engine
from ._lib import PhyloligoError

MIN_NB_W_PER_FASTA_FOR_MUL_CPU = 20   # reference :64
WINDOW_METRICS = {"JSD": 0, "KL": 1, "Eucl": 2}
BATCH_WINDOWS = 1 << 18               # windows per device batch
YIELD_WINDOWS = 50000                 # rows per yielded chunk (reference :424)


# ---------------------------------------------------------------------------
# assembly in memory
# ---------------------------------------------------------------------------
class Assembly:
    """Records of a FASTA file: ids, and one uint8 buffer holding every sequence without line
    breaks (records separated by one newline), with the offset and length of each."""

    def __init__(self, genome):
        text = np.fromfile(genome, dtype=np.uint8)
        begin, end = engine.fasta_index(text)
        n = int(begin.shape[0])
        self.ids = []
        parts, offsets, lengths = [], np.zeros(n, dtype=np.int64), np.zeros(n, dtype=np.int64)
        pos = 0
        sep = np.array([10], dtype=np.uint8)
        starts = engine.fasta_header_starts(text, begin, end)
        for i in range(n):
            header = bytes(text[int(starts[i]):int(begin[i])])
            fields = header[1:].split()
            self.ids.append(fields[0].decode("latin-1") if fields else "")
            seg = text[int(begin[i]):int(end[i])]
            seq = seg[(seg != 32) & ((seg < 9) | (seg > 13))]  # blanks are transparent, as in the profiling kernels
            offsets[i], lengths[i] = pos, seq.shape[0]
            parts.append(seq)
            parts.append(sep)
            pos += seq.shape[0] + 1
        self.n = n
        self.text = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)
        self.offsets, self.lengths = offsets, lengths
        self._device = None

    def device_text(self):
        if self._device is None:
            self._device = engine.text_to_device(self.text)
        return self._device

    def sequence(self, i):
        o, m = int(self.offsets[i]), int(self.lengths[i])
        return self.text[o:o + m].tobytes().decode("latin-1")


_ASSEMBLIES = {}


def _assembly(genome):
    key = (os.path.abspath(genome), os.path.getmtime(genome), os.path.getsize(genome))
    if key not in _ASSEMBLIES:
        _ASSEMBLIES.clear()  # one assembly resident at a time
        _ASSEMBLIES[key] = Assembly(genome)
    return _ASSEMBLIES[key]


# ---------------------------------------------------------------------------
# device calls
# ---------------------------------------------------------------------------
def _ptr(t):
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def window_count_byte(d_text, d_begin, d_end, value):
    lib = _lib.load()
    n = int(d_begin.shape[0])
    out = torch.empty((n,), dtype=torch.int64, device=d_text.device)
    _lib.check(lib.po_window_count_byte(_ptr(d_text), _ptr(d_begin), _ptr(d_end), n, int(value), _ptr(out), _stream()),
               "po_window_count_byte")
    return out


def window_distances(metric, d_freq, d_ref):
    """distance of every row of d_freq (float64, n x dim) to d_ref (float64, dim): float64 (n,)"""
    lib = _lib.load()
    if metric not in WINDOW_METRICS:
        raise PhyloligoError("Error, unknown method {}".format(metric))
    n, dim = int(d_freq.shape[0]), int(d_freq.shape[1])
    out = torch.empty((n,), dtype=torch.float64, device=d_freq.device)
    _lib.check(lib.po_window_distances(WINDOW_METRICS[metric], _ptr(d_freq), n, dim, int(d_freq.stride(0)), _ptr(d_ref),
                                       _ptr(out), _stream()), "po_window_distances")
    return out


def _pair(metric, a, b):
    engine.require_cuda()
    A = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).reshape(1, -1)).cuda()
    B = torch.from_numpy(np.ascontiguousarray(np.asarray(b, dtype=np.float64)).ravel()).cuda()
    if A.shape[1] != B.shape[0]:
        raise PhyloligoError("profiles must have the same length")
    return float(window_distances(metric, A, B)[0].item())


def KL(a, b):
    """1-D KL of the reference (:71-86)"""
    return _pair("KL", a, b)


def Eucl(a, b):
    """Euclidean distance x1000 (:88-92)"""
    return _pair("Eucl", a, b)


def JSD(a, b):
    """1-D JSD x1000 (:94-123)"""
    return _pair("JSD", a, b)


# ---------------------------------------------------------------------------
# profiles
# ---------------------------------------------------------------------------
def _check_strand(strand):
    if strand not in ("both", "minus", "plus"):
        print("Error, strand parameter of selectd_strand() should be choose from {'both', 'minus', 'plus'}",
              file=sys.stderr)
        sys.exit(1)


def compute_frequency(seq, n_max_freq_in_windows=1.0, pattern="1111", strand="both"):
    """Frequency vector of one window (:274-301); all-NaN above the 'N' limit."""
    pattern = str(pattern)
    _check_strand(strand)
    dim = 4 ** pattern.count("1")
    if len(seq) and (seq.count("N") / len(seq)) > float(n_max_freq_in_windows):
        return np.array([np.nan] * dim)
    res = engine.profile_sequences([seq], pattern, strand, want=("freq64",))
    return res["freq64"][0].cpu().numpy()


def compute_whole_composition(genome, pattern, strand, nb_jobs=1):
    """Counts of every record summed, then frequencies (:303-314)."""
    pattern = str(pattern)
    _check_strand(strand)
    engine.require_cuda()
    asm = _assembly(genome)
    dim = 4 ** pattern.count("1")
    if asm.n == 0:
        return np.zeros(dim, dtype=np.float64)
    d_begin = torch.from_numpy(asm.offsets).cuda()
    d_end = torch.from_numpy(asm.offsets + asm.lengths).cuda()
    res = engine.profile_device(asm.device_text(), d_begin, d_end, pattern, strand, want=("counts", "totals"))
    counts = res["counts"].cpu().numpy().astype(np.int64).sum(axis=0)
    total = int(res["totals"].cpu().numpy().astype(np.int64).sum())
    if total == 0:
        return np.zeros(dim, dtype=np.float64)
    return counts.astype(np.float64) / float(total)  # int / int true division of the reference (:263)


def compute_distance_joblib(mth_dist, mcp, seq, pattern, strand, n_max_freq_in_windows):
    """One window (:317-324)."""
    freq = compute_frequency(seq, n_max_freq_in_windows, pattern, strand)
    if np.isnan(freq).any():
        return 0.0  # every NaN term is zeroed by posdef_check_value
    if mth_dist == "JSD":
        return JSD(freq, mcp)
    if mth_dist == "KL":
        return KL(freq, mcp)
    return Eucl(freq, mcp)


# ---------------------------------------------------------------------------
# windows
# ---------------------------------------------------------------------------
def window_table(lengths, windows_size, windows_step):
    """Windows of every record as arrays (record, start, length, displayed_start, displayed_stop):
    the three cases of make_genome_chunk (:343-407), vectorised per record."""
    w, t = int(windows_size), int(windows_step)
    rec, start, size, dstart, dstop = [], [], [], [], []
    for i, n in enumerate(int(v) for v in lengths):
        if n < w:  # :351-353, one window, no sliding
            rec.append(np.array([i]))
            start.append(np.array([0]))
            size.append(np.array([n]))
            dstart.append(np.array([0]))
            dstop.append(np.array([n]))
            continue
        s = np.arange(0, n - w, t, dtype=np.int64)
        if s.shape[0] == 0:
            continue
        a = (s + w / 2 - t / 2).astype(np.int64)  # int() truncates, the values are positive
        b = (s + w / 2 + t / 2).astype(np.int64)
        if n < MIN_NB_W_PER_FASTA_FOR_MUL_CPU * t:  # :359-382
            da = np.where(s == 0, 1, a)
            db = np.where(s == n - w, n, b)
        else:  # :388-403
            da = np.where(a == (w / 2 - t / 2), 1, a)
            edge = b - t / 2 + w / 2
            db = np.where((edge >= n - t) & (edge <= n), n, b)
        rec.append(np.full(s.shape[0], i, dtype=np.int64))
        start.append(s)
        size.append(np.full(s.shape[0], w, dtype=np.int64))
        dstart.append(da)
        dstop.append(db)
    if not rec:
        z = np.zeros(0, dtype=np.int64)
        return z, z, z, z, z
    return tuple(np.concatenate(x).astype(np.int64) for x in (rec, start, size, dstart, dstop))


def make_genome_chunk(genome, windows_size, windows_step, options, nbchunk=500):
    """(chunk_info, chunk_sequences) like the reference's generator (:343-407)."""
    asm = _assembly(genome)
    rec, start, size, dstart, dstop = window_table(asm.lengths, windows_size, windows_step)
    for c0 in range(0, rec.shape[0], nbchunk):
        c1 = min(rec.shape[0], c0 + nbchunk)
        info = [[asm.ids[int(rec[k])], int(dstart[k]), int(dstop[k])] for k in range(c0, c1)]
        seqs = []
        for k in range(c0, c1):
            o = int(asm.offsets[rec[k]] + start[k])
            seqs.append(asm.text[o:o + int(size[k])].tobytes().decode("latin-1"))
        yield info, seqs


def window_distance_vector(asm, rec, start, size, mcp, mth_dist, pattern, strand, n_max):
    """Distances of the given windows to `mcp` (numpy float64), batched on the device."""
    engine.require_cuda()
    d_text = asm.device_text()
    d_ref = torch.from_numpy(np.ascontiguousarray(np.asarray(mcp, dtype=np.float64))).cuda()
    out = np.empty(rec.shape[0], dtype=np.float64)
    wb = asm.offsets[rec] + start
    for c0 in range(0, rec.shape[0], BATCH_WINDOWS):
        c1 = min(rec.shape[0], c0 + BATCH_WINDOWS)
        d_begin = torch.from_numpy(np.ascontiguousarray(wb[c0:c1])).cuda()
        d_end = torch.from_numpy(np.ascontiguousarray(wb[c0:c1] + size[c0:c1])).cuda()
        freq = engine.profile_device(d_text, d_begin, d_end, pattern, strand, want=("freq64",))["freq64"]
        if freq.shape[1] != d_ref.shape[0]:
            raise PhyloligoError("the reference profile has %d bins, the pattern gives %d" % (d_ref.shape[0], freq.shape[1]))
        dist = window_distances(mth_dist, freq, d_ref).cpu().numpy()
        n_count = window_count_byte(d_text, d_begin, d_end, ord("N")).cpu().numpy()
        m = size[c0:c1]
        with np.errstate(divide="ignore", invalid="ignore"):
            over = (n_count / m) > float(n_max)  # count / len <= n_max keeps the window (:294), same division
        over &= m > 0
        dist[over] = 0.0
        out[c0:c1] = dist
    return out


def sliding_windows_distances(genome, mcp_comparison, mth_dist="JSD", pattern="1111", windows_size=5000,
                              windows_step=500, options=None):
    """Rows [seq_id, displayed_start, displayed_stop, distance], yielded in chunks (:409-453)."""
    strand = getattr(options, "strand", "both")
    n_max = getattr(options, "n_max_freq_in_windows", 0.4)
    _check_strand(strand)
    asm = _assembly(genome)
    rec, start, size, dstart, dstop = window_table(asm.lengths, windows_size, windows_step)
    dist = window_distance_vector(asm, rec, start, size, mcp_comparison, mth_dist, str(pattern), strand, n_max)
    ids = asm.ids
    for c0 in range(0, rec.shape[0], YIELD_WINDOWS):
        c1 = min(rec.shape[0], c0 + YIELD_WINDOWS)
        yield [[ids[r], a, b, d] for r, a, b, d in zip(rec[c0:c1].tolist(), dstart[c0:c1].tolist(),
                                                       dstop[c0:c1].tolist(), dist[c0:c1].tolist())]


def vector_to_matrix(profile):
    """:125-126"""
    return list((zip(*(iter(profile),) * int(math.sqrt(len(profile))))))


# ---------------------------------------------------------------------------
# command line (flag surface of reference :482-518)
# ---------------------------------------------------------------------------
def get_cmd(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-i", "--assembly", action="store", required=True, dest="genome",
                        help="multifasta of the genome assembly")
    parser.add_argument("-c", "--conta", action="store", dest="conta",
                        help="multifasta of the contaminant species training set")
    parser.add_argument("-r", "--host", action="store", dest="host",
                        help="optional host species training set in multifasta")
    parser.add_argument("-n", "--n_max_freq_in_windows", action="store", type=float, dest="n_max_freq_in_windows",
                        default=0.4, help="maximum proportion of N tolerated in a window [0~1]")
    parser.add_argument("-k", "--lgMot", action="store", dest="k", type=int, default=4,
                        help="word wise/ kmer lenght/ k [default:%(default)d]")
    parser.add_argument("-p", "--pattern", action="store", dest="pattern",
                        help="pattern to use for frequency computation")
    parser.add_argument("-w", "--windows_size", action="store", dest="windows_size", type=int, default=5000,
                        help="Sliding windows size (bp)")
    parser.add_argument("-t", "--windows_step", action="store", dest="windows_step", type=int, default=500,
                        help="Sliding windows step size(bp)")
    parser.add_argument("-d", "--distance", action="store", dest="dist", choices=["JSD", "Eucl", "KL"], default="JSD",
                        help="distance method between two signatures [default:%(default)s]")
    parser.add_argument("-s", "--strand", action="store", default="both", choices=["both", "plus", "minus"],
                        help="strand used to compute microcomposition. [default:%(default)s]")
    parser.add_argument("-u", "--cpu", action="store", dest="threads_max", type=int, default=4,
                        help="accepted for compatibility [default:%(default)d]")
    parser.add_argument("-W", "--workdir", action="store", dest="workdir", default="", help="working directory")
    return parser.parse_args(argv)


def _write_rows(path, genome, mcp, options):
    with open(path, "w") as outf:
        for res in sliding_windows_distances(genome, mcp_comparison=mcp, mth_dist=options.dist, pattern=options.pattern,
                                             windows_size=options.windows_size, windows_step=options.windows_step,
                                             options=options):
            outf.write("".join("\t".join(map(str, t)) + "\n" for t in res))


def main(argv=None):
    options = get_cmd(argv)
    print("Genome : {}".format(options.genome))
    base_genome = os.path.basename(options.genome)
    if options.workdir and not os.path.isdir(options.workdir):
        os.makedirs(options.workdir)

    if not options.conta:
        print("Contaminant : {}".format(None))
        output = os.path.join(options.workdir, base_genome + ".mcp_windows_vs_whole_" + options.dist + ".dist")
    else:
        base_conta = os.path.basename(options.conta)
        print("Contaminant : {} ".format(options.conta))
        output = base_genome + ".mcp_hostwindows_vs_"
        if options.host:
            base_host = os.path.basename(options.host)
            print("Host : {}".format(options.host))
            output = os.path.join(options.workdir, output + "host_" + base_host + "_" + options.dist + ".dist")
        else:
            print("Host : None, using whole genome")
            output = os.path.join(options.workdir, output + "wholegenome_" + options.dist + ".dist")

    if not options.pattern and options.k:
        options.pattern = "1" * options.k

    # the reference always compares with the whole-genome profile (:567, :590), also when -r is given
    genome = compute_whole_composition(options.genome, options.pattern, options.strand, nb_jobs=options.threads_max)

    if not options.conta:
        if not options.windows_size and not options.windows_step:
            print("Warning, no sliding window parameters (-w and -t )\n"
                  "The signature will be computed from the whole genome\n"
                  "Computing signature from the whole genome", file=sys.stderr)
            output = os.path.join(options.workdir, base_genome + ".microcomposition.mat")
            with open(output, "w") as outf:
                outf.write(str(vector_to_matrix([float(v) for v in genome])))
            return 0
        print("Computing microcomposition signaure and distances to genome")
    else:
        conta = compute_whole_composition(options.conta, options.pattern, options.strand, nb_jobs=options.threads_max)

    _write_rows(output, options.genome, genome, options)
    if options.conta:
        output = os.path.join(options.workdir,
                              base_genome + ".mcp_hostwindows_vs_conta_" + base_conta + "_" + options.dist + ".dist")
        _write_rows(output, options.genome, conta, options)
    return 0


if __name__ == "__main__":
    main()
    sys.exit(0)
