#!/usr/bin/env python3
"""Sliding-window mode (Kount.py) on one GPU: windows/s and window-bases/s on a synthetic assembly,
next to the Python restatement of the reference timed on a sample of the same windows.

    python tests/tools/bench_kount.py --contigs 2000 --mean-len 100000 -w 5000 -t 500 -d JSD"""
import argparse, json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from phyloligo_b200 import kount, synth, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--contigs", type=int, default=2000)
ap.add_argument("--mean-len", type=int, default=100000)
ap.add_argument("-w", type=int, default=5000)
ap.add_argument("-t", type=int, default=500)
ap.add_argument("-d", default="JSD")
ap.add_argument("--cpu-windows", type=int, default=300)
args = ap.parse_args()

d = tempfile.mkdtemp(prefix="po_kount_")
path = os.path.join(d, "asm.fasta")
fasta, total = synth.fast_fasta_bytes(args.contigs, args.mean_len, seed=7)
np.asarray(fasta).tofile(path)


class Opt:
    strand, n_max_freq_in_windows = "both", 0.4


def run():
    t0 = time.perf_counter()
    mcp = kount.compute_whole_composition(path, "1111", "both")
    t1 = time.perf_counter()
    rows = 0
    for chunk in kount.sliding_windows_distances(path, mcp, args.d, "1111", args.w, args.t, Opt):
        rows += len(chunk)
    torch.cuda.synchronize()
    return mcp, rows, t1 - t0, time.perf_counter() - t1


run()                      # warm-up (CUDA context, assembly load)
kount._ASSEMBLIES.clear()  # time the load again
l0 = _lib.launch_count()
mcp, rows, t_whole, t_windows = run()
launches = _lib.launch_count() - l0
# device-only time of the window stage (assembly resident)
asm = kount._assembly(path)
rec, start, size, _, _ = kount.window_table(asm.lengths, args.w, args.t)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
kount.window_distance_vector(asm, rec, start, size, mcp, args.d, "1111", "both", 0.4)
e1.record()
torch.cuda.synchronize()
dev_s = e0.elapsed_time(e1) * 1e-3
# CPU: the Python restatement of the reference on a sample of the same windows, one core
from oracle import kount_oracle as ko, phylo_oracle as po
records = [(asm.ids[i], asm.sequence(i)) for i in range(min(asm.n, 3))]
wins = ko.make_windows(records, args.w, args.t)[: args.cpu_windows]
t0 = time.perf_counter()
for sid, a, b, s in wins:  # the reference's per-window work: Counter over Python strings, count2freq, 1-D distance
    f = po.compute_frequency(s, "1111", "both") if (s.count("N") / len(s)) <= 0.4 else np.full(256, np.nan)
    {"JSD": ko.JSD, "KL": ko.KL, "Eucl": ko.Eucl}[args.d](np.asarray(f, dtype=np.float64), mcp)
cpu_s = time.perf_counter() - t0
print(json.dumps({
    "workload": "Kount.py windows: %d contigs x %d kb, -w %d -t %d -d %s, k=4 both" % (args.contigs, args.mean_len // 1000, args.w, args.t, args.d),
    "bases": int(total), "windows": int(rows), "window_bases": int(size.sum()),
    "whole_composition_s": t_whole, "windows_e2e_s": t_windows, "windows_device_s": dev_s,
    "windows_per_s_e2e": rows / t_windows, "windows_per_s_device": rows / dev_s,
    "window_gbases_per_s_device": float(size.sum()) / dev_s / 1e9, "gpu_launches": int(launches),
    "cpu_port_windows_per_s_1core": len(wins) / cpu_s, "cpu_sample_windows": len(wins),
}))
import shutil; shutil.rmtree(d)
