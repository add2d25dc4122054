import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests")
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


def golden_files():
    return sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def load_golden(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def libc_rand():
    """glibc srand/rand, the generator the reference uses (common.h:103-110)."""
    import ctypes
    libc = ctypes.CDLL("libc.so.6")

    def draw(seed, n):
        libc.srand(seed)
        return np.array([libc.rand() for _ in range(n)], np.int32)
    return draw
