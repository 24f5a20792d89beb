#!/usr/bin/env python3
"""python tools/check_cli_multi_gpu.py [N]: run the phyloligo.py command line on one GPU and under
``torchrun --nproc-per-node N`` on the same synthetic assembly, for every output mode, and compare
the files byte for byte (same kernels and tile grid => identical matrices)."""
import filecmp, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from phyloligo_b200 import synth

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
d = tempfile.mkdtemp(prefix="po_cli_")
fasta = os.path.join(d, "asm.fasta")
synth.write_fasta(fasta, synth.make_sequences(2100, 3000, seed=5) + [b"", b"NNNN"], line=80)
cases = [("None", "JSD", ["-k", "4"]), ("memmap", "JSD", ["-k", "4"]), ("h5py", "Eucl", ["-p", "1101011"]),
         ("memmap", "KT", ["-k", "3"]), ("None", "SC", ["-k", "4", "-s", "plus"]), ("h5py", "BC", ["-k", "5"])]
env = dict(os.environ, PYTHONPATH=ROOT)
ok = True
for large, metric, extra in cases:
    outs = []
    for tag, launcher in (("one", [sys.executable]),
                          ("many", [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node",
                                    str(world), "--master-addr", "127.0.0.1", "--master-port", "29533"])):
        out = os.path.join(d, "%s_%s_%s.out" % (tag, large, metric))
        freq = out + ".freq"
        cmd = launcher + ["-m", "phyloligo_b200.phyloligo", "-i", fasta, "-d", metric, "--method", "joblib", "--large", large,
                          "-o", out, "-q", freq, "-w", d] + extra
        res = subprocess.run(cmd, env=env, capture_output=True, text=True, cwd=ROOT)
        if res.returncode != 0:
            print(tag, large, metric, "FAILED\n", res.stdout[-2000:], res.stderr[-4000:])
            ok = False
        outs.append((out, freq))
    same = all(os.path.exists(a) and os.path.exists(b) and filecmp.cmp(a, b, shallow=False)
               for a, b in zip(outs[0], outs[1]))
    print("%-6s %-4s %s: %d-GPU files identical to the 1-GPU files: %s" % (large, metric, " ".join(extra), world, same), flush=True)
    ok = ok and same
sys.exit(0 if ok else 1)
