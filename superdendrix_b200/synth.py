"""Synthetic inputs of the DepMap-20Q2 shape.

Matrix: S x F Bernoulli with per-feature frequency clip(0.6 * j**-0.9, 0.004, 0.7),
all-zero features dropped; sample ids ``ACH-%06d``; feature names ``G{j}_{A|I|O}_MUT``.
Weights: 2C-score-like fp64, 2*N(0,1) clipped to +-20 with 2 % exact zeros.
Seeds are fixed by the callers (matrix 20, weights 3) so every run sees the same data.
"""
from __future__ import annotations

import numpy as np

CONFIGS = {
    # name: (S samples, F features before dropping empties)
    "C0": (770, 400),
    "C1": (770, 400),
    "C2": (1000, 2000),
    "C3": (1000, 5000),
}


def feature_matrix(S: int, F: int, seed: int = 20):
    """Returns (dense uint8 S x F', sample ids, feature names)."""
    rng = np.random.default_rng(seed)
    freq = np.clip(0.6 * np.arange(1, F + 1, dtype=np.float64) ** -0.9, 0.004, 0.7)
    dense = (rng.random((S, F)) < freq).astype(np.uint8)
    keep = dense.sum(axis=0) > 0
    idx = np.nonzero(keep)[0]
    dense = np.ascontiguousarray(dense[:, keep])
    samples = ["ACH-%06d" % (i + 1) for i in range(S)]
    kinds = "AIO"
    features = ["G%d_%s_MUT" % (j + 1, kinds[j % 3]) for j in idx]
    return dense, samples, features


def weights(S: int, n_profiles: int = 1, seed: int = 3) -> np.ndarray:
    """(n_profiles, S) fp64 weight vectors, already sign-flipped (src/superdendrix.py:105-106)."""
    rng = np.random.default_rng(seed)
    w = np.clip(2.0 * rng.standard_normal((n_profiles, S)), -20.0, 20.0)
    w[rng.random((n_profiles, S)) < 0.02] = 0.0
    return w


def pick_module(dense: np.ndarray, k: int = 3) -> np.ndarray:
    """Deterministic module: the features ranked 2..k+1 by carrier count (stable)."""
    order = np.argsort(-dense.sum(axis=0), kind="stable")
    return np.sort(order[1:1 + k]).astype(np.int32)


def write_features_tsv(path: str, dense: np.ndarray, samples, features) -> None:
    """sample -> events TSV (format of src/i_o.py:54-63)."""
    with open(path, "w") as fh:
        fh.write("#Samples\tEvents\n")
        for i, s in enumerate(samples):
            fh.write("\t".join([s] + [features[j] for j in np.nonzero(dense[i])[0]]) + "\n")


def write_profile_tsv(path: str, samples, w: np.ndarray, names) -> None:
    """profile table: header row, first column sample id (src/superdendrix.py:85-89)."""
    with open(path, "w") as fh:
        fh.write("\t".join(["sample"] + list(names)) + "\n")
        for i, s in enumerate(samples):
            fh.write("\t".join([s] + [repr(float(w[p, i])) for p in range(w.shape[0])]) + "\n")
