#!/usr/bin/env python3
"""Generate tests/golden/kount_cli_golden.npz: output files of the reference's UNMODIFIED bin/Kount.py
(through oracle/run_reference_cli.py --script Kount.py) on a small synthetic assembly.

    python tests/golden/make_kount_cli_golden.py
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from phyloligo_b200 import synth  # noqa: E402

RUNNER = os.path.join(ROOT, "oracle", "run_reference_cli.py")
CASES = [
    ("jsd_k4", ["-k", "4", "-w", "1000", "-t", "500", "-d", "JSD"]),
    ("kl_k3_plus", ["-k", "3", "-w", "1500", "-t", "700", "-d", "KL", "-s", "plus"]),
    ("eucl_k4_minus", ["-k", "4", "-w", "2000", "-t", "1000", "-d", "Eucl", "-s", "minus"]),
]


def assembly():
    seqs = synth.make_sequences(7, 5000, seed=23) + [b"ACGTTGCA" * 90]
    return synth.to_fasta_bytes(seqs, line=70)


def main():
    work = tempfile.mkdtemp(prefix="po_kount_golden_")
    path = os.path.join(work, "asm.fasta")
    open(path, "wb").write(assembly())
    out, names = {}, []
    for name, args in CASES:
        outdir = os.path.join(work, name)
        res = subprocess.run([sys.executable, RUNNER, "--script", "Kount.py", "-i", path, "-u", "2", "-W", outdir] + args,
                             cwd=work, capture_output=True, text=True)
        files = sorted(os.listdir(outdir)) if os.path.isdir(outdir) else []
        if res.returncode != 0 or not files:
            print(name, "FAILED", res.stderr[-400:])
            continue
        out[name + "_args"] = np.array(" ".join(args))
        out[name + "_files"] = np.array(files)
        for f in files:
            out[name + "_" + f] = np.frombuffer(open(os.path.join(outdir, f), "rb").read(), dtype=np.uint8)
        names.append(name)
        print(name, files, [os.path.getsize(os.path.join(outdir, f)) for f in files])
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "kount_cli_golden.npz"), **out)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
