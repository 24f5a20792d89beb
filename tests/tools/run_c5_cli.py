#!/usr/bin/env python3
"""C5 through the command line on N GPUs: JSD k=5 on 5 kb contigs with --large memmap, every rank
streaming its rows into the shared file.  The full configuration (1 M contigs, a 4 TB matrix) does
not fit any scratch disk here; --contigs sets the size actually run.

    python tests/tools/run_c5_cli.py --gpus 8 --contigs 60000 --workdir /dev/shm

Prints one JSON line: wall time of the torchrun command, stage times reported by rank 0, and a check
of the file (exact zero diagonal, symmetry and sampled entries against the float64 oracle)."""
import argparse, json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=8)
ap.add_argument("--contigs", type=int, default=60000)
ap.add_argument("--mean-len", type=int, default=5000)
ap.add_argument("--workdir", default="/dev/shm")
args = ap.parse_args()

from phyloligo_b200 import synth
d = tempfile.mkdtemp(prefix="po_c5_", dir=args.workdir)
fasta_path, out_path = os.path.join(d, "asm.fasta"), os.path.join(d, "jsd.mat")
fasta, total = synth.fast_fasta_bytes(args.contigs, args.mean_len, seed=5)
np.asarray(fasta).tofile(fasta_path)
cmd = [sys.executable]
if args.gpus > 1:
    cmd += ["-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
            "--master-port", "29561"]
cmd += ["-m", "phyloligo_b200.phyloligo", "-i", fasta_path, "-k", "5", "-d", "JSD", "--method", "joblib", "--large",
        "memmap", "-o", out_path, "-w", d]
t0 = time.perf_counter()
res = subprocess.run(cmd, cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT, PO_VERBOSE="1"), capture_output=True, text=True)
wall = time.perf_counter() - t0
stages = [l for l in res.stderr.splitlines() if l.startswith("phyloligo_b200:")]
out = {"config": "C5 (scaled): JSD k=5 both, %d contigs x %d kb, --large memmap, %d GPUs" % (args.contigs, args.mean_len // 1000, args.gpus),
       "returncode": res.returncode, "wall_s": wall, "stages": stages[-1] if stages else None,
       "fasta_bytes": int(len(fasta)), "matrix_bytes": args.contigs * args.contigs * 4}
if res.returncode != 0:
    out["stderr_tail"] = res.stderr[-2000:]
else:
    n = args.contigs
    M = np.memmap(out_path, dtype=np.float32, mode="r", shape=(n, n))
    rng = np.random.default_rng(1)
    rows = np.sort(rng.choice(n, 64, replace=False))
    sym = all(np.array_equal(M[r, rows], M[rows, r]) for r in rows)
    diag = bool((M[rows, rows] == 0).all())
    # sampled entries against the oracle on the reference's float32 profiles
    from oracle import phylo_oracle as po
    text = np.fromfile(fasta_path, dtype=np.uint8)
    from phyloligo_b200 import engine
    begin, end = engine.fasta_index(text)
    pick = rows[:12]
    prof = {int(r): po.frequency_np(bytes(text[begin[r]:end[r]]).replace(b"\n", b"").decode(), "11111", "both").astype(np.float32).astype(np.float64) for r in pick}
    worst = 0.0
    for a in pick:
        for b in pick:
            want = po.JSD(prof[int(a)], prof[int(b)])
            got = float(M[a, b])
            if want > 0:
                worst = max(worst, abs(got - want) / want)
    out.update({"pairs_per_s_wall": n * (n + 1) / 2 / wall, "symmetric_sample": bool(sym), "zero_diagonal_sample": diag,
                "max_rel_err_sample": worst})
print(json.dumps(out))
import shutil; shutil.rmtree(d, ignore_errors=True)
