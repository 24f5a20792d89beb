#!/usr/bin/env python3
"""Generate tests/golden/select_cli_golden.npz: the reference's UNMODIFIED bin/phyloselect.py (through
oracle/run_reference_cli.py --script phyloselect.py) clustering matrices that the reference's UNMODIFIED
bin/phyloligo.py wrote for the assembly of make_cli_golden.py -- the whole chain is reference code.

    python tests/golden/make_select_cli_golden.py

Stored per case: arguments, the matrix file's bytes (text) or array (memmap), the bytes of
data_cluster_indexes.dat and of every data_fasta_*.fa.
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
sys.path.insert(0, HERE)
import make_cli_golden as mk  # noqa: E402

RUNNER = os.path.join(ROOT, "oracle", "run_reference_cli.py")
CASES = [  # name, phyloligo arguments, phyloselect arguments
    ("jsd_text_k2", ["-k", "4", "-d", "JSD", "--method", "joblib"], ["-m", "kmedoids", "-k", "2"]),
    ("jsd_text_k3", ["-k", "4", "-d", "JSD", "--method", "joblib"], ["-m", "kmedoids", "-k", "3"]),
    ("eucl_text_k4", ["-k", "3", "-d", "Eucl", "--method", "joblib"], ["-m", "kmedoids", "-k", "4"]),
    ("jsd_memmap_k3", ["-k", "4", "-d", "JSD", "--method", "joblib", "--large", "memmap"], ["-m", "kmedoids", "-k", "3", "--large", "memmap"]),
]


def main():
    fasta, n = mk.assembly()
    work = tempfile.mkdtemp(prefix="po_select_golden_")
    path = os.path.join(work, "asm.fasta")
    open(path, "wb").write(fasta)
    out = {}
    names = []
    for name, ligo, sel in CASES:
        mat = os.path.join(work, name + ".mat")
        env = dict(os.environ, PO_REF_JOBLIB_THREADS="1" if "--large" in ligo else "0")
        subprocess.run([sys.executable, RUNNER, "-i", path, "-o", mat, "-w", work, "-c", "2"] + ligo, check=True, cwd=work,
                       env=env, capture_output=True)
        outdir = os.path.join(work, name + "_sel")
        res = subprocess.run([sys.executable, RUNNER, "--script", "phyloselect.py", "-i", mat, "-o", outdir, "-f", path, "--noX"] + sel,
                             cwd=work, capture_output=True, text=True)
        if res.returncode != 0:
            print(name, "FAILED", res.stderr[-500:])
            continue
        out[name + "_ligo_args"] = np.array(" ".join(ligo))
        out[name + "_select_args"] = np.array(" ".join(sel))
        out[name + "_matrix_bytes"] = np.frombuffer(open(mat, "rb").read(), dtype=np.uint8)
        out[name + "_indexes"] = np.frombuffer(open(os.path.join(outdir, "data_cluster_indexes.dat"), "rb").read(), dtype=np.uint8)
        fa = sorted(f for f in os.listdir(outdir) if f.startswith("data_fasta_"))
        out[name + "_fasta_names"] = np.array(fa)
        for f in fa:
            out[name + "_" + f] = np.frombuffer(open(os.path.join(outdir, f), "rb").read(), dtype=np.uint8)
        names.append(name)
        labels = [int(l.split()[0]) for l in open(os.path.join(outdir, "data_cluster_indexes.dat"))]
        print("%-16s stdout %s  cluster sizes %s  files %s" % (name, res.stdout.strip().splitlines(), np.bincount(labels).tolist(), fa))
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "select_cli_golden.npz"), **out)
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
