"""GPU parity: po_profile_batch (through the C ABI) against the reference golden
vectors and the oracle.  Counts, totals and float64/float32 frequencies are bit exact."""
import numpy as np
import pytest
import torch

from conftest import golden_case_arrays
from oracle import phylo_oracle as po
from phyloligo_b200 import engine, synth

pytestmark = pytest.mark.gpu


def _profile(seqs, pattern, strand):
    res = engine.profile_sequences(seqs, pattern, strand, want=("counts", "totals", "freq64", "freq32"))
    return {k: v.cpu().numpy() for k, v in res.items()}


def test_golden_profiles_bit_exact(profile_golden):
    seqs = profile_golden["sequences"]
    groups = {}
    for case in profile_golden["cases"]:
        groups.setdefault((case["pattern"], case["strand"]), []).append(case)
    for (pat, strand), cases in groups.items():
        batch = [seqs[c["seq"]] for c in cases]
        out = _profile(batch, pat, strand)
        for r, c in enumerate(cases):
            counts, total, freq = golden_case_arrays(c)
            assert int(out["totals"][r]) == total, (pat, strand, c["seq"])
            assert np.array_equal(out["counts"][r].astype(np.int64), counts), (pat, strand, c["seq"])
            assert np.array_equal(out["freq64"][r], freq), (pat, strand, c["seq"])
            assert np.array_equal(out["freq32"][r], freq.astype(np.float32))


@pytest.mark.parametrize("pattern", ["1111", "11111", "111010011", "110101", "1" * 7, "1000000000000000001",
                                     "11011000000000000000000000000011"])
@pytest.mark.parametrize("strand", ["plus", "minus", "both"])
def test_random_contigs_match_oracle(pattern, strand):
    seqs = synth.make_sequences(40, 3000, seed=21) + [b"", b"NNNN", b"ACGT", b"acgtn" * 50]
    out = _profile(seqs, pattern, strand)
    for r, s in enumerate(seqs):
        counts, total = po.count_vector_np(s, pattern, strand)
        assert int(out["totals"][r]) == total
        assert np.array_equal(out["counts"][r].astype(np.int64), counts)
        assert np.array_equal(out["freq64"][r], po.frequency_np(s, pattern, strand))


def test_fasta_text_with_line_breaks_and_long_contig():
    # wrapped FASTA, CRLF, a record longer than one chunk sweep, empty record
    seqs = synth.make_sequences(6, 60_000, seed=5) + [b""] + synth.make_sequences(3, 100, seed=6)
    text = synth.to_fasta_bytes(seqs, line=60).replace(b"\n", b"\r\n", 7)
    begin, end = engine.fasta_index(text)
    assert len(begin) == len(seqs)
    res = engine.profile_text(text, "1111", "both", want=("counts", "totals"), begin=begin, end=end)
    counts = res["counts"].cpu().numpy()
    for r, s in enumerate(seqs):
        c, t = po.count_vector_np(s, "1111", "both")
        assert np.array_equal(counts[r].astype(np.int64), c)
        assert int(res["totals"][r].item()) == t


def test_large_k_global_histogram():
    seqs = synth.make_sequences(5, 5000, seed=8)
    out = _profile(seqs, "1" * 9, "both")  # 4^9 bins: global-memory histogram path
    for r, s in enumerate(seqs):
        c, t = po.count_vector_np(s, "1" * 9, "both")
        assert int(out["totals"][r]) == t
        assert np.array_equal(out["counts"][r].astype(np.int64), c)


def test_property_total_is_checksum_at_scale():
    # size-independent property at a larger size: total == sum(counts) == windows in valid runs,
    # both == plus + minus + junction (junction <= width-1 words)
    text, nbases = synth.fast_fasta_bytes(2000, 20_000, seed=2)
    begin, end = engine.fasta_index(text)
    d_text = engine.text_to_device(text)
    d_b = torch.from_numpy(begin).cuda()
    d_e = torch.from_numpy(end).cuda()
    outs = {s: engine.profile_device(d_text, d_b, d_e, "1111", s, want=("counts", "totals")) for s in ("plus", "minus", "both")}
    for s in outs:
        assert torch.equal(outs[s]["counts"].to(torch.int64).sum(dim=1), outs[s]["totals"])
    assert torch.equal(outs["plus"]["totals"], outs["minus"]["totals"])
    junction = outs["both"]["counts"] - outs["plus"]["counts"] - outs["minus"]["counts"]
    assert int(junction.min().item()) >= 0 and int(junction.sum(dim=1).max().item()) <= 3
    # minus counts are the reverse-complement permutation of plus counts for a contiguous k-mer
    k = 4
    idx = np.arange(4 ** k)
    digits = [(idx >> (2 * j)) & 3 for j in range(k)]  # least significant first
    rc = sum(((digits[j] ^ 1) << (2 * (k - 1 - j))) for j in range(k))
    assert torch.equal(outs["minus"]["counts"], outs["plus"]["counts"][:, torch.from_numpy(rc).cuda()])


def _layout(seqs, rng, widths, crlf=False, blank_lines=False, trailing_space=False):
    """FASTA text whose sequence lines have the given widths (cycled / random)."""
    parts = []
    for i, s in enumerate(seqs):
        parts.append(b">r%d some description\n" % i)
        p = 0
        while p < len(s):
            w = int(widths[int(rng.integers(0, len(widths)))])
            line = s[p:p + w]
            p += w
            if trailing_space and rng.random() < 0.2:
                # what Biopython's line.rstrip() drops: blanks of every kind, also several of them
                line += [b" ", b"\t", b"\x0b", b"\x0c", b" \t ", b"\t\t"][int(rng.integers(0, 6))]
            parts.append(line + (b"\r\n" if crlf else b"\n"))
            if blank_lines and rng.random() < 0.1:
                parts.append(b"\n")
    return b"".join(parts)


def _adversarial_sequences(rng):
    seqs = synth.make_sequences(12, 4000, seed=77)
    seqs += [b"", b"A", b"ACG", b"ACGTACGTACGTACG", b"ACGTACGTACGTACGT", b"ACGTACGTACGTACGTA", b"N" * 300,
             b"ACGT" * 20 + b"N" + b"TTGCA" * 30, b"acgtnACGT" * 40, b"ACGT" * 500]
    # N runs placed around multiples of common line widths
    s = bytearray(synth.make_sequences(1, 3000, seed=78)[0])
    for p in (59, 60, 61, 79, 80, 81, 160, 1023, 1024, 1100):
        s[p:p + int(rng.integers(1, 20))] = b"N" * 19
    seqs.append(bytes(s[:3000]))
    return seqs


@pytest.mark.parametrize("kernel", ["auto", "general"])
@pytest.mark.parametrize("pattern,strand", [("1111", "both"), ("1111", "minus"), ("11111", "plus"),
                                            ("111010011", "plus"), ("1101011", "both"), ("1" * 6, "both"),
                                            ("1001", "both"), ("10101", "minus"),
                                            # patterns that are not their own mirror image: minus-strand words from the
                                            # reverse-complement register of the segment kernel
                                            ("111010011", "both"), ("111010011", "minus"), ("110101", "minus"),
                                            ("1101", "both"), ("1100000000000001", "both")])
def test_irregular_fasta_layouts(kernel, pattern, strand, monkeypatch):
    """Line-segment fast path against the oracle on layouts that break its assumptions:
    random line widths, widths below/at/above the segment limits, CRLF, blank lines,
    trailing blanks, N runs across line ends."""
    if kernel == "general":
        monkeypatch.setenv("PO_PROFILE_KERNEL", "general")
    else:
        monkeypatch.delenv("PO_PROFILE_KERNEL", raising=False)
    rng = np.random.default_rng(5)
    seqs = _adversarial_sequences(rng)
    expect = [po.count_vector_np(s, pattern, strand) for s in seqs]
    layouts = [([80], False, False, False), ([60], True, False, False), ([16], False, False, False),
               ([17], False, False, False), ([15], False, True, False), ([127], False, False, False),
               ([128], False, False, False), ([129], False, False, True), ([500], False, False, False),
               ([1, 2, 3, 5, 80, 81, 200], True, True, True), ([61], False, True, True), ([4], False, False, False)]
    for widths, crlf, blank, trail in layouts:
        text = _layout(seqs, rng, widths, crlf, blank, trail)
        begin, end = engine.fasta_index(text)
        assert len(begin) == len(seqs)
        res = engine.profile_text(text, pattern, strand, want=("counts", "totals", "freq64"), begin=begin, end=end)
        counts = res["counts"].cpu().numpy()
        totals = res["totals"].cpu().numpy()
        for r, (c, t) in enumerate(expect):
            assert int(totals[r]) == t, (widths, crlf, blank, trail, r)
            assert np.array_equal(counts[r].astype(np.int64), c), (widths, crlf, blank, trail, r)
        f = res["freq64"].cpu().numpy()
        for r in (0, 5, len(seqs) - 1):
            assert np.array_equal(f[r], po.frequency_np(seqs[r], pattern, strand))
