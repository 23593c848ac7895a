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


def burn_cases():
    """Matrices (S x F, samples x features) that take the four branches of the burn-in plan:
    name -> (dense, burn-in on the transpose of the trade orientation?, dense-favouring pairs?)."""
    import numpy as np
    rng = np.random.default_rng(17)
    out = {}
    # one feature carried by 60 % of the samples, the others sparse; samples balanced: transpose, uniform pairs (the C0 situation)
    d = (rng.random((150, 60)) < 0.06).astype(np.uint8)
    d[:, 3] = rng.random(150) < 0.6
    out["dense_feature"] = (d, True, False)
    # one sample with most of the features (a hypermutator) and balanced features: the trade rows (features) are the balanced side
    d = (rng.random((150, 60)) < 0.06).astype(np.uint8)
    d[7, :] = rng.random(60) < 0.8
    out["dense_sample"] = (d, False, False)
    # both: a frequent feature and a hypermutated sample, the sample rows being the less unbalanced side
    d = (rng.random((150, 60)) < 0.025).astype(np.uint8)
    d[:, 3] = rng.random(150) < 0.7
    d[7, :] = rng.random(60) < 0.45
    out["dense_both_transpose"] = (d, True, True)
    d = (rng.random((150, 60)) < 0.025).astype(np.uint8)
    d[:, 3] = rng.random(150) < 0.25
    d[7, :] = rng.random(60) < 0.95
    out["dense_both_trade_rows"] = (d, False, True)
    return out
