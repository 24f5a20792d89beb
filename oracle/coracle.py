"""ctypes wrapper of oracle/_build/liboracle.so (TEST / BASELINE INFRASTRUCTURE ONLY)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle.so")
METRICS = {"Eucl": 0, "JSD": 1, "KT": 2, "BC": 3, "SC": 4}
STRANDS = {"plus": 0, "minus": 1, "both": 2}
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB):
            subprocess.run(["make", "-s", "-C", HERE], check=True)
        lib = C.CDLL(LIB)
        lib.oracle_count.restype = C.c_uint64
        lib.oracle_count.argtypes = [C.c_void_p, C.c_int64, C.c_char_p, C.c_int, C.c_void_p]
        lib.oracle_profile_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_char_p, C.c_int,
                                             C.c_void_p, C.c_int]
        lib.oracle_pairwise_rows.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64,
                                             C.c_void_p, C.c_int]
        for f in ("oracle_eucl", "oracle_jsd", "oracle_bc", "oracle_kt", "oracle_sc"):
            getattr(lib, f).restype = C.c_double
            getattr(lib, f).argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _lib = lib
    return _lib


def _threads(threads):
    return int(threads or os.cpu_count() or 1)


def count(seq: bytes, pattern: str, strand: str):
    lib = load()
    dim = 4 ** pattern.count("1")
    counts = np.zeros(dim, dtype=np.uint64)
    buf = np.frombuffer(seq, dtype=np.uint8) if len(seq) else np.zeros(1, dtype=np.uint8)
    total = lib.oracle_count(buf.ctypes.data, len(seq), pattern.encode(), STRANDS[strand], counts.ctypes.data)
    return counts.astype(np.int64), int(total)


def profile_batch(text: np.ndarray, begin, end, pattern: str, strand: str, threads=None):
    """Records are byte ranges of `text` holding bare sequence bytes (no line breaks)."""
    lib = load()
    n = len(begin)
    dim = 4 ** pattern.count("1")
    freq = np.zeros((n, dim), dtype=np.float64)
    begin = np.ascontiguousarray(begin, dtype=np.int64)
    end = np.ascontiguousarray(end, dtype=np.int64)
    lib.oracle_profile_batch(text.ctypes.data, begin.ctypes.data, end.ctypes.data, n, pattern.encode(),
                             STRANDS[strand], freq.ctypes.data, _threads(threads))
    return freq


def pair(metric: str, a, b):
    lib = load()
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    fn = getattr(lib, "oracle_" + metric.lower())
    return fn(a.ctypes.data, b.ctypes.data, a.shape[0])


def pairwise_rows(metric: str, X, r0=0, r1=None, threads=None):
    lib = load()
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, d = X.shape
    r1 = n if r1 is None else r1
    out = np.empty((r1 - r0, n), dtype=np.float64)
    lib.oracle_pairwise_rows(METRICS[metric], X.ctypes.data, n, d, r0, r1, out.ctypes.data, _threads(threads))
    return out
