#!/usr/bin/env python3
"""Time the profiling kernel on device-generated fixed-length FASTA (CUDA events):
    python tools/bench_profile.py --n 50000 --len 15000 --pattern 111010011 --strand both"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--len", type=int, default=15000)
ap.add_argument("--pattern", default="111010011")
ap.add_argument("--strand", default="both")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--want", default="freq32")
args = ap.parse_args()
import torch
from phyloligo_b200 import engine, synth
dev = torch.device("cuda", 0)
text, b, e, bases = synth.device_fasta(args.n, args.len, 3, dev)
want = tuple(args.want.split(","))
fn = lambda: engine.profile_device(text, b, e, args.pattern, args.strand, want=want)
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps): fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.reps
dim = 4 ** args.pattern.count("1")
nbytes = int(text.shape[0]) + args.n * dim * 4
print("%s %s n=%d len=%d [%s]: %.3f ms  %.1f Gbase/s  %.0f GB/s algorithmic (%.3f of 6560)" % (
    args.pattern, args.strand, args.n, args.len, os.environ.get("PO_SEG_WARPREC_MAXDIM", "-"), ms, bases / ms / 1e6, nbytes / ms / 1e6, nbytes / ms / 1e6 / 6560))
