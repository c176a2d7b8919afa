import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "triple-tensor-decomposition-with-admm_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests", "golden"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(x, ref):
    """Relative error in the Frobenius norm, the measure north_star's 1e-8 bar is stated in."""
    x = np.asarray(x, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    d = np.linalg.norm((x - ref).ravel())
    n = np.linalg.norm(ref.ravel())
    return d / n if n > 0 else d


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
