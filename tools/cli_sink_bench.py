#!/usr/bin/env python3
"""The distance stage of the command line into a raw memmap file, for several settings of the host
sink (PO_SINK_WARM / PO_SINK_WARM_THREADS / PO_SINK_COPY_THREADS / PO_SINK_SLOT_MB), fresh file and
rewrite:   python tools/cli_sink_bench.py --n 50000 [--dir /dev/shm]"""
import argparse, os, sys, time, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phyloligo_b200 import phyloligo

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=50000)
ap.add_argument("--dir", default="/dev/shm")
ap.add_argument("--settings", default="none:0:8:32,none:0:16:32,populate:4:8:32,fallocate:4:8:32,fallocate:2:8:32,fallocate:4:8:96,fallocate:8:8:32")
args = ap.parse_args()
torch.cuda.set_device(0)
g = torch.Generator(device="cuda").manual_seed(1)
X = torch.rand((args.n, 256), device="cuda", generator=g) ** 2
X /= X.sum(dim=1, keepdim=True)
Xh = X.cpu().numpy()
work = tempfile.mkdtemp(dir=args.dir)
out = os.path.join(work, "d.mat")
nbytes = args.n * args.n * 4
print("n %d, output %.1f GB under %s" % (args.n, nbytes / 1e9, args.dir), flush=True)
for setting in args.settings.split(","):
    warm, wt, ct, slot = setting.split(":")
    if warm == "default":  # fresh file: fallocate only; existing file: premap
        os.environ.pop("PO_SINK_WARM", None)
    else:
        os.environ["PO_SINK_WARM"] = warm
    os.environ["PO_SINK_WARM_THREADS"] = wt
    os.environ["PO_SINK_COPY_THREADS"] = ct
    os.environ["PO_SINK_SLOT_MB"] = slot
    for label in ("fresh", "rewrite"):
        if label == "fresh" and os.path.exists(out):
            os.unlink(out)
        fdir = tempfile.mkdtemp(dir=work)
        fname = os.path.join(fdir, "frequencies")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        phyloligo.compute_distances("joblib", "memmap", Xh, fname, out, "JSD", 16, 250, work)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("warm=%-9s warm_threads=%s copy_threads=%-2s slot=%3s MB  %-7s  %.3f s  %.2f GB/s" % (warm, wt, ct, slot, label, dt, nbytes / dt / 1e9), flush=True)
m = np.memmap(out, dtype=np.float32, mode="r", shape=(args.n, args.n))
print("check: diag", float(np.abs(np.diagonal(m[:2000, :2000])).max()), "sym", bool(np.array_equal(m[:1000, :1000], m[:1000, :1000].T)), "row0 sum", float(m[0].sum()))
shutil.rmtree(work, ignore_errors=True)
