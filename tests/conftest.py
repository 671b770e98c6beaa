"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    """Golden vectors produced by the imported reference (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "ntxent_golden.npz")
    return np.load(path)


GOLDEN_CASES = [
    "c1_b256_d128_t05",
    "aligned_b100_d64_t01",
    "ragged_b37_d20_t05",
    "b192_d256_t01",
    "b130_d128_t005_aligned",
    "b64_d128_t1_scaled",
    "b1_d16_t05",
]


def rel_fro(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def rel_max(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())
