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
def kern_mat():
    return np.load(os.path.join(GOLDEN, "kern_matfiles.npz"))


@pytest.fixture(scope="session")
def matrix_mat():
    return np.load(os.path.join(GOLDEN, "matrix_matfiles.npz"))


@pytest.fixture(scope="session")
def gp_ref():
    return np.load(os.path.join(GOLDEN, "gp_reference.npz"))


@pytest.fixture(scope="session")
def rand_ref():
    return np.load(os.path.join(GOLDEN, "random_reference.npz"))


CASES = {  # tag -> component types (tests/golden/make_golden.py CASES)
    "c_rbf_white": ["rbf", "white"],
    "c_rbfard_white": ["rbfard", "white"],
    "c_m52_white": ["matern52", "white"],
    "c_m32_bias_white": ["matern32", "bias", "white"],
    "c_lin_poly_white": ["lin", "poly", "white"],
    "c_all": ["rbf", "rbfard", "matern32", "matern52", "lin", "poly", "bias", "white"],
}
SINGLE = ["rbf", "rbfard", "matern32", "matern52", "lin", "poly", "white", "bias"]


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1.0)))
