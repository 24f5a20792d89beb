#!/usr/bin/env python3
"""JSD kernel on sparse profiles (C5-like: k=5, 5 kb contigs generated on the device), symmetric resident matrix:
    python tools/bench_jsd_sparse.py --n 60000"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=60000)
ap.add_argument("--len", type=int, default=5000)
ap.add_argument("--pattern", default="11111")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
import torch
from phyloligo_b200 import engine, synth
from phyloligo_b200._lib import FLAG_MIRROR, FLAG_SKIP_LOWER
dev = torch.device("cuda", 0)
text, b, e, bases = synth.device_fasta(args.n, args.len, 5, dev)
X = engine.profile_device(text, b, e, args.pattern, "both", want=("freq32",))["freq32"]
del text
P, aux, dim = engine.prepare(X, "JSD")
out = torch.empty((args.n, args.n), dtype=torch.float32, device=dev)
fn = lambda: engine.distance_block("JSD", P, aux, dim, 0, args.n, 0, args.n, out, 0, 0, FLAG_SKIP_LOWER | FLAG_MIRROR)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
pairs = args.n * (args.n + 1) / 2
print("JSD sparse n=%d dim=%d: %.2f ms  %.3e pairs/s  %.3e terms/s  checksum %.6f" % (args.n, dim, ms, pairs / ms * 1e3, pairs * dim / ms * 1e3, float(out[:2000, :2000].double().sum())))
