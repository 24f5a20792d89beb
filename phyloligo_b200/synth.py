"""Deterministic synthetic multi-FASTA for the BASELINE.json configurations.

Shapes follow SURVEY.md section 8(d): per-contig GC from a two-component mixture
(host GC ~ N(0.52, 0.03); 10 % "contaminant" contigs GC ~ N(0.66, 0.03)), i.i.d.
bases at that GC, log-normal lengths clipped to [0.5, 2] x mean (or fixed),
0.1 % of positions inside N-runs of 10-100, 5 % lower-case (soft-masked)
stretches, lines wrapped at 80 columns, headers ``>c{index}``.
"""
from __future__ import annotations

import numpy as np

LINE = 80

CONFIGS = {
    # name: (n_contigs, mean_len, seed, length_model)
    "C1": (1_000, 10_000, 1, "lognormal"),
    "C2": (100_000, 20_000, 2, "lognormal"),
    "C3": (50_000, 15_000, 3, "lognormal"),
    "C4": (20_000, 375, 4, "short"),
    "C5": (1_000_000, 5_000, 5, "lognormal"),
}


def contig_lengths(n, mean_len, rng, model="lognormal"):
    if model == "fixed":
        return np.full(n, mean_len, dtype=np.int64)
    if model == "short":  # C4: uniform 150-600 bp, ~0.5 % empty records
        ln = rng.integers(150, 601, size=n).astype(np.int64)
        ln[rng.random(n) < 0.005] = 0
        return ln
    sigma = 0.35
    ln = rng.lognormal(mean=np.log(mean_len) - 0.5 * sigma * sigma, sigma=sigma, size=n)
    return np.clip(ln, 0.5 * mean_len, 2.0 * mean_len).astype(np.int64)


def _bases(length, gc, rng):
    """i.i.d. bases: P(G)=P(C)=gc/2, P(A)=P(T)=(1-gc)/2, as ASCII bytes."""
    r = rng.integers(0, 1 << 16, size=length, dtype=np.uint16)
    is_gc = (r >> 1) < np.uint16(gc * 32768.0)
    low = (r & 1).astype(np.uint8)
    #            low=0 low=1
    # is_gc      C     G
    # not        A     T
    out = np.where(is_gc, np.where(low == 0, ord("C"), ord("G")), np.where(low == 0, ord("A"), ord("T")))
    return out.astype(np.uint8)


def make_sequences(n, mean_len, seed, model="lognormal", n_frac=0.001, lower_frac=0.05, all_n_frac=0.0):
    """Return a list of ``bytes`` sequences (no headers, no newlines)."""
    rng = np.random.default_rng(seed)
    lens = contig_lengths(n, mean_len, rng, model)
    contaminant = rng.random(n) < 0.10
    gc = np.where(contaminant, rng.normal(0.66, 0.03, n), rng.normal(0.52, 0.03, n))
    gc = np.clip(gc, 0.05, 0.95)
    seqs = []
    for i in range(n):
        L = int(lens[i])
        s = _bases(L, float(gc[i]), rng)
        if L and all_n_frac and rng.random() < all_n_frac:
            s[:] = ord("N")
        if L >= 200:
            # N-runs: expected n_frac of positions, run length 10-100
            n_runs = rng.poisson(L * n_frac / 55.0)
            for _ in range(n_runs):
                w = int(rng.integers(10, 101))
                p = int(rng.integers(0, max(1, L - w)))
                s[p:p + w] = ord("N")
            # soft-masked stretches: expected lower_frac of positions, 50-500 long
            n_low = rng.poisson(L * lower_frac / 275.0)
            for _ in range(n_low):
                w = int(rng.integers(50, 501))
                p = int(rng.integers(0, max(1, L - w)))
                s[p:p + w] |= 0x20
        seqs.append(s.tobytes())
    return seqs


def to_fasta_bytes(seqs, line=LINE):
    """Wrap sequences at `line` columns with ``>c{index}`` headers."""
    parts = []
    for i, s in enumerate(seqs):
        parts.append(b">c%d\n" % i)
        L = len(s)
        if L == 0:
            continue
        arr = np.frombuffer(s, dtype=np.uint8)
        full = L // line
        if full:
            body = np.empty((full, line + 1), dtype=np.uint8)
            body[:, :line] = arr[: full * line].reshape(full, line)
            body[:, line] = 10
            parts.append(body.tobytes())
        if L % line:
            parts.append(arr[full * line:].tobytes() + b"\n")
    return b"".join(parts)


def write_fasta(path, seqs, line=LINE):
    with open(path, "wb") as fh:
        fh.write(to_fasta_bytes(seqs, line))


def make_config(name, scale=1.0):
    """Sequences of one named configuration, optionally with the contig count scaled."""
    n, mean_len, seed, model = CONFIGS[name]
    n = max(2, int(round(n * scale)))
    return make_sequences(n, mean_len, seed, model)


def fast_fasta_bytes(n, mean_len, seed, model="lognormal", line=LINE):
    """Bulk generator for bench-sized inputs (GBs): one vectorised pass.

    Same distributional shape as make_sequences (mixture GC, clipped log-normal
    lengths, N-runs, soft-masked stretches) but drawn with whole-array numpy
    operations.  Returns (fasta_bytes: np.ndarray[uint8], total_bases: int).
    """
    rng = np.random.default_rng(seed)
    lens = contig_lengths(n, mean_len, rng, model)
    contaminant = rng.random(n) < 0.10
    gc = np.where(contaminant, rng.normal(0.66, 0.03, n), rng.normal(0.52, 0.03, n))
    gc = np.clip(gc, 0.05, 0.95)
    total = int(lens.sum())
    thr = np.repeat((gc * 32768.0).astype(np.uint16), lens)
    seq = np.empty(total, dtype=np.uint8)
    step = 1 << 26
    lut = np.array([ord("A"), ord("T"), ord("C"), ord("G")], dtype=np.uint8)
    for s0 in range(0, total, step):
        s1 = min(total, s0 + step)
        r = rng.integers(0, 1 << 16, size=s1 - s0, dtype=np.uint16)
        idx = (((r >> 1) < thr[s0:s1]).astype(np.uint8) << 1) | (r & 1).astype(np.uint8)
        seq[s0:s1] = lut[idx]
    del thr
    # N-runs (~0.1 % of positions) and soft-masked stretches (~5 %), placed globally
    n_runs = int(total * 0.001 / 55.0)
    starts = rng.integers(0, max(1, total - 100), size=n_runs)
    widths = rng.integers(10, 101, size=n_runs)
    for p, w in zip(starts.tolist(), widths.tolist()):
        seq[p:p + w] = ord("N")
    n_low = int(total * 0.05 / 275.0)
    starts = rng.integers(0, max(1, total - 500), size=n_low)
    widths = rng.integers(50, 501, size=n_low)
    for p, w in zip(starts.tolist(), widths.tolist()):
        seq[p:p + w] |= 0x20
    # lay out as FASTA: header + wrapped lines
    headers = [b">c%d\n" % i for i in range(n)]
    hlen = np.fromiter((len(h) for h in headers), dtype=np.int64, count=n)
    nl = (lens + line - 1) // line
    rec_bytes = hlen + lens + nl
    rec_off = np.concatenate(([0], np.cumsum(rec_bytes)))
    out = np.full(int(rec_off[-1]), 10, dtype=np.uint8)
    seq_off = np.concatenate(([0], np.cumsum(lens)))
    for i in range(n):
        o = int(rec_off[i])
        h = headers[i]
        out[o:o + len(h)] = np.frombuffer(h, dtype=np.uint8)
        o += len(h)
        L = int(lens[i])
        if L == 0:
            continue
        s = seq[int(seq_off[i]):int(seq_off[i]) + L]
        full = L // line
        if full:
            out[o:o + full * (line + 1)].reshape(full, line + 1)[:, :line] = s[: full * line].reshape(full, line)
        rem = L - full * line
        if rem:
            o2 = o + full * (line + 1)
            out[o2:o2 + rem] = s[full * line:]
    return out, total


def device_fasta(n, length, seed, device, line=LINE, chunk_elems=1 << 26):
    """Fixed-length variant ("fixed-length variant for roofline runs") generated on the
    device: n records of `length` bases, the same GC mixture, N runs and soft-masked stretches,
    80-column lines, headers ``>c{index}`` zero-padded to a fixed width.  For the configurations
    whose FASTA text (5 GB at C5) would take minutes to draw on the host.
    Returns (text uint8 tensor [n * record_bytes + 64], begin int64 [n], end int64 [n], total_bases)."""
    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    w = len(str(max(1, n - 1)))
    hlen = 2 + w + 1
    nl = -(-length // line)
    rec = hlen + length + nl
    text = torch.empty(n * rec + 64, dtype=torch.uint8, device=device)
    text[n * rec:] = 10
    recs = text[: n * rec].view(n, rec)
    idx = torch.arange(n, device=device, dtype=torch.int64)
    recs[:, 0] = ord(">")
    recs[:, 1] = ord("c")
    for d in range(w):
        recs[:, 2 + d] = ((idx // (10 ** (w - 1 - d))) % 10 + 48).to(torch.uint8)
    recs[:, 2 + w] = 10
    body = recs[:, hlen:]
    body[:, :] = 10  # newline columns; the base columns are overwritten below
    contaminant = torch.rand(n, generator=g, device=device) < 0.10
    gc = torch.where(contaminant, 0.66 + 0.03 * torch.randn(n, generator=g, device=device),
                     0.52 + 0.03 * torch.randn(n, generator=g, device=device)).clamp_(0.05, 0.95)
    col = torch.arange(length, device=device, dtype=torch.int64)
    dst = col + col // line  # column of base j inside the record body (one newline after every `line` bases)
    lut = torch.tensor([ord("A"), ord("T"), ord("C"), ord("G")], dtype=torch.uint8, device=device)
    rows_per = max(1, chunk_elems // max(1, length))
    for r0 in range(0, n, rows_per):
        r1 = min(n, r0 + rows_per)
        u = torch.randint(0, 1 << 16, (r1 - r0, length), generator=g, device=device, dtype=torch.int32)
        is_gc = (u >> 1).to(torch.float32) < (gc[r0:r1, None] * 32768.0)
        codes = (is_gc.to(torch.int64) << 1) | (u & 1).to(torch.int64)
        body[r0:r1].index_copy_(1, dst, lut[codes])
    # N runs (~0.1 % of positions, width 55) and soft-masked stretches (~5 %, width 275): at most one per record
    for prob, width, kind in ((length * 0.001 / 55.0, 55, "n"), (length * 0.05 / 275.0, 275, "low")):
        if length <= width:
            continue
        pick = torch.nonzero(torch.rand(n, generator=g, device=device) < min(1.0, prob)).flatten()
        if pick.numel() == 0:
            continue
        start = torch.randint(0, length - width, (pick.numel(),), generator=g, device=device)
        pos = dst[(start[:, None] + torch.arange(width, device=device)[None, :])]
        flat = (pick[:, None] * rec + hlen + pos).flatten()
        if kind == "n":
            text[flat] = ord("N")
        else:
            text[flat] = text[flat] | 0x20
    begin = idx * rec + hlen
    end = begin + length + nl
    return text, begin, end, n * length
