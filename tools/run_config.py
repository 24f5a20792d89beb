#!/usr/bin/env python3
"""Run one of the BASELINE.json configurations (C1..C5, optionally scaled) through the public
host API on one GPU and print a JSON line: profiling Gbases/s, distance pairs/s, end-to-end time.

    python tools/run_config.py C3 --scale 1.0 --metric Eucl
    python tools/run_config.py C5 --scale 0.05 --sink discard

The matrix is streamed to the host in row panels exactly as the --large memmap path does
(engine.PanelStreamer); --sink memmap writes a real raw float32 file under --workdir."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from phyloligo_b200 import engine, synth

CFG = {  # pattern, strand, metrics
    "C1": ("1111", "both", ["Eucl"]),
    "C2": ("1111", "both", ["JSD"]),
    "C3": ("111010011", "both", ["Eucl", "BC"]),
    "C4": ("1111", "both", ["KT", "SC"]),
    "C5": ("11111", "both", ["JSD"]),
}
ap = argparse.ArgumentParser()
ap.add_argument("config", choices=sorted(CFG))
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--metric", default=None)
ap.add_argument("--sink", default="discard", choices=["discard", "memmap"])
ap.add_argument("--workdir", default="/tmp")
args = ap.parse_args()
pattern, strand, metrics = CFG[args.config]
if args.metric:
    metrics = [args.metric]
n, mean_len, seed, model = synth.CONFIGS[args.config]
n = max(64, int(round(n * args.scale)))
t0 = time.perf_counter()
if model == "short":
    fasta = np.frombuffer(synth.to_fasta_bytes(synth.make_sequences(n, mean_len, seed, model, all_n_frac=0.005)), dtype=np.uint8)
    total = None
else:
    fasta, total = synth.fast_fasta_bytes(n, mean_len, seed, model)
gen_s = time.perf_counter() - t0
torch.cuda.set_device(0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
out = {"config": args.config, "contigs": n, "pattern": pattern, "strand": strand, "fasta_bytes": int(fasta.shape[0])}
for rep in range(2):  # second pass is the measured one
    t0 = time.perf_counter()
    begin, end = engine.fasta_index(fasta)
    t_index = time.perf_counter() - t0
    ev[0].record()
    res = engine.profile_text(fasta, pattern, strand, want=("freq32",), begin=begin, end=end)
    ev[1].record()
    torch.cuda.synchronize()
out["records"] = int(begin.shape[0])
bases = int((end - begin).sum()) if total is None else total
out["bases"] = bases
out["index_host_ms"] = 1e3 * t_index
out["profile_h2d_and_kernel_ms"] = ev[0].elapsed_time(ev[1])
X = res["freq32"]
out["dim"] = int(X.shape[1])
for metric in metrics:
    dev_metric = "EuclGram" if metric == "Eucl" else metric
    sink_arr = None
    if args.sink == "memmap":
        path = os.path.join(args.workdir, "run_config_%s_%s.bin" % (args.config, metric))
        sink_arr = np.memmap(path, dtype=np.float32, shape=(n, n), mode="w+")
    checks = {"sum": 0.0}
    def sink(r0, r1, host):
        if sink_arr is not None:
            sink_arr[r0:r1] = host
        else:
            checks["sum"] += float(host[:, ::997].sum())   # touch the data (discard sink)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = engine.PanelStreamer(X, dev_metric, torch.float32)
    pairs_computed = st.run(sink)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out[metric] = {"e2e_seconds": dt, "unique_pairs_per_s": n * (n + 1) / 2 / dt, "entries_written_per_s": n * n / dt,
                   "pairs_computed": int(pairs_computed), "symmetric": bool(st.symmetric), "sink": args.sink}
    if sink_arr is not None:
        sink_arr.flush(); del sink_arr; os.unlink(path)
    del st
    torch.cuda.empty_cache()
out["profiling_gbases_per_s_incl_h2d"] = bases / (out["profile_h2d_and_kernel_ms"] * 1e-3) / 1e9
print(json.dumps(out))
