"""Build libphyloligo_b200.so in-tree with nvcc for sm_100a (and nothing else)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libphyloligo_b200.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-O3",
]
OBJ_DIR = os.path.join(PKG_DIR, "lib", "obj")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(PKG_DIR), "include", "phyloligo_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build_library(force=False, verbose=False, extra_flags=()):
    """Compile every .cu under csrc/ into lib/libphyloligo_b200.so."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libphyloligo_b200.so")
    os.makedirs(OBJ_DIR, exist_ok=True)
    # compile the translation units in parallel, then link
    jobs = []
    objs = []
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if not f.endswith(".cu")]
    headers.append(os.path.join(os.path.dirname(PKG_DIR), "include", "phyloligo_b200.h"))
    hdr_time = max(os.path.getmtime(h) for h in headers)
    for src in sources():
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.isfile(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and os.path.getmtime(obj) > hdr_time):
            continue
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + os.environ.get("PO_NVCC_EXTRA", "").split() + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        jobs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, proc in jobs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s" % (src, out))
        if verbose and out:
            print(out, file=sys.stderr)
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    flags = ["-Xptxas", "-v"] if "-v" in sys.argv else []
    print(build_library(force=True, verbose=True, extra_flags=flags))
