import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def profile_golden():
    with open(os.path.join(GOLDEN, "profile_golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def distance_golden():
    return dict(np.load(os.path.join(GOLDEN, "distance_golden.npz")))


def golden_case_arrays(case):
    """(counts int64[dim], total, freq float64[dim]) of one golden profile case."""
    dim = 4 ** case["pattern"].count("1")
    counts = np.zeros(dim, dtype=np.int64)
    freq = np.zeros(dim, dtype=np.float64)
    idx = np.asarray(case["nz_index"], dtype=np.int64)
    if idx.size:
        counts[idx] = case["nz_count"]
        freq[idx] = [float.fromhex(h) for h in case["nz_freq_hex"]]
    return counts, case["total"], freq
