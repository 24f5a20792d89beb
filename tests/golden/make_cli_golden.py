#!/usr/bin/env python3
"""Generate tests/golden/cli_golden.npz by running the reference's UNMODIFIED command-line script
(/root/reference/phylopackage/bin/phyloligo.py, through oracle/run_reference_cli.py: stand-ins for the
absent third-party modules only) on a small synthetic assembly, for every mode of it that runs.

    python tests/golden/make_cli_golden.py

Stored per case: the argument list, the -q frequency matrix and the distance matrix exactly as the
reference wrote them (text parsed back with np.loadtxt, raw float32 memmap read back).  The assembly
itself is regenerated from its seed by the tests (phyloligo_b200.synth is deterministic); its bytes'
SHA-256 is stored to catch drift.
"""
import hashlib
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

CASES = [  # name, arguments (without -i / -o / -q / -w), how the matrix is stored
    ("eucl_k4_joblib", ["-k", "4", "-d", "Eucl", "--method", "joblib"], "text"),
    ("jsd_k4_joblib", ["-k", "4", "-d", "JSD", "--method", "joblib"], "text"),
    ("jsd_spaced_plus_joblib", ["-p", "110101", "-s", "plus", "-d", "JSD", "--method", "joblib"], "text"),
    ("eucl_k3_minus_joblib", ["-k", "3", "-s", "minus", "-d", "Eucl", "--method", "joblib"], "text"),
    ("bc_k4_joblib", ["-k", "4", "-d", "BC", "--method", "joblib"], "text"),
    ("kt_k2_joblib", ["-k", "2", "-d", "KT", "--method", "joblib"], "text"),
    ("eucl_k4_scoop", ["-k", "4", "-d", "Eucl", "--method", "scoop"], "text"),
    ("jsd_k5_scoop", ["-k", "5", "-d", "JSD", "--method", "scoop"], "text"),
    ("kt_k2_scoop", ["-k", "2", "-d", "KT", "--method", "scoop"], "text"),
    ("eucl_k4_memmap", ["-k", "4", "-d", "Eucl", "--method", "joblib", "--large", "memmap"], "memmap"),
    ("jsd_k4_memmap", ["-k", "4", "-d", "JSD", "--method", "joblib", "--large", "memmap"], "memmap"),
]


def assembly():
    seqs = synth.make_sequences(36, 1200, seed=19) + [b"NNNNNNNNNNNN", b"acgtnnACGTTGCAgg" * 30]
    return synth.to_fasta_bytes(seqs, line=70), len(seqs)


def main():
    fasta, n = assembly()
    work = tempfile.mkdtemp(prefix="po_cli_golden_")
    path = os.path.join(work, "asm.fasta")
    open(path, "wb").write(fasta)
    out = {"fasta_sha256": np.array(hashlib.sha256(fasta).hexdigest()), "n_records": np.int64(n)}
    names = []
    for name, args, kind in CASES:
        mat, freq = os.path.join(work, name + ".mat"), os.path.join(work, name + ".freq")
        cmd = [sys.executable, RUNNER, "-i", path, "-o", mat, "-q", freq, "-w", work, "-c", "2"] + args
        env = dict(os.environ, PO_REF_JOBLIB_THREADS="1" if kind == "memmap" else "0")
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=work, env=env)
        if res.returncode != 0 or not os.path.exists(mat):
            print("%-28s the reference does not run this mode here: %s" % (name, (res.stderr.strip().splitlines() or ["?"])[-1][:150]))
            continue
        if kind == "text":
            M = np.loadtxt(mat)
        else:
            M = np.fromfile(mat, dtype=np.float32).reshape(n, n)
        F = np.loadtxt(freq)
        out[name + "_args"] = np.array(" ".join(args))
        out[name + "_matrix"] = M
        out[name + "_freq"] = F
        names.append(name)
        print("%-28s matrix %s %s  freq %s  stdout lines: %s" % (name, M.shape, M.dtype, F.shape, res.stdout.strip().splitlines()))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "cli_golden.npz"), **out)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
