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
def golden_dir():
    return GOLDEN


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def curveball_cases():
    z = load_npz("curveball_replay.npz")
    names = sorted({k.split("__")[0] for k in z.files if "__" in k})
    return {n: {k.split("__")[1]: z[k] for k in z.files if k.startswith(n + "__")} for n in names}, str(z["python"])


def rows_of(dense):
    """Packed rows over the shorter dimension, as find_presences orients them
    (src/curveball.py:10-11): samples are rows when F >= S, else features."""
    from superdendrix_b200 import bitpack
    S, F = dense.shape
    return bitpack.pack_rows(dense if F >= S else dense.T)
