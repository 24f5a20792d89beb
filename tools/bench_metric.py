#!/usr/bin/env python3
"""Time one metric's distance kernel on random profiles (device resident, CUDA events):
    python tools/bench_metric.py --metric Eucl --n 20000 --dim 4096 [--exact]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--metric", default="Eucl")
ap.add_argument("--n", type=int, default=20000)
ap.add_argument("--dim", type=int, default=4096)
ap.add_argument("--exact", action="store_true")
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
if args.metric == "Eucl" and not args.exact:
    args.metric = "EuclGram"
import numpy as np, torch
from phyloligo_b200 import engine
from phyloligo_b200._lib import FLAG_MIRROR, FLAG_SKIP_LOWER
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand((args.n, args.dim), device="cuda", generator=g) ** 4
X /= X.sum(dim=1, keepdim=True)
if args.metric in ("KT", "SC"):
    X = torch.round(X * 3 * args.dim) / (3 * args.dim)   # tie-heavy
out = torch.empty((args.n, args.n), dtype=torch.float32, device="cuda")
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for rep in range(args.reps + 1):
    e0.record()
    P, aux, dim = engine.prepare(X, args.metric)
    e1.record()
    engine.distance_block(args.metric, P, aux, dim, 0, args.n, 0, args.n, out, 0, 0, FLAG_SKIP_LOWER | FLAG_MIRROR)
    e2.record()
    torch.cuda.synchronize()
    if rep:
        pairs = args.n * (args.n + 1) / 2
        ms = e1.elapsed_time(e2)
        print("%s n=%d dim=%d%s: prepare %.2f ms, tiles %.2f ms -> %.3e pairs/s, %.1f TFLOP/s (2*D flop per pair)" % (
            args.metric, args.n, args.dim, " exact" if args.exact else "", e0.elapsed_time(e1), ms, pairs / ms * 1e3,
            pairs * 2 * args.dim / ms * 1e3 / 1e12))
