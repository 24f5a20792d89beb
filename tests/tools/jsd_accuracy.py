#!/usr/bin/env python3
"""Max relative error of the GPU JSD matrix against the float64 oracle on a few profile sets (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from phyloligo_b200 import engine, synth
from oracle import coracle

def report(name, X32):
    got = engine.distance_matrix_device(torch.from_numpy(X32).cuda(), "JSD", torch.float64, symmetric=False).cpu().numpy()
    want = coracle.pairwise_rows("JSD", X32.astype(np.float64))
    off = ~np.eye(len(X32), dtype=bool)
    rel = np.abs(got - want)[off] / np.maximum(want[off], 1e-300)
    print("%-28s n=%d dim=%d  max rel %.3e  mean rel %.3e  (min JSD %.3e)" % (name, X32.shape[0], X32.shape[1], rel.max(), rel.mean(), want[off].min()))

for k, L in ((4, 20000), (4, 2000), (5, 5000), (6, 15000), (3, 500)):
    seqs = synth.make_sequences(384, L, seed=7)
    text, b, e = engine.sequences_to_text(seqs)
    F = coracle.profile_batch(text, b, e, "1" * k, "both").astype(np.float32)
    report("synthetic k=%d L=%d" % (k, L), F)
rng = np.random.default_rng(1)
base = rng.dirichlet(np.ones(256), size=1)
near = (base * (1 + 1e-3 * rng.standard_normal((256, 256)))).astype(np.float32)
near /= near.sum(axis=1, keepdims=True)
report("near-identical (1e-3 noise)", near)
sparse = rng.dirichlet(np.full(4096, 0.05), size=200).astype(np.float32)
report("sparse dirichlet 4096", sparse)
