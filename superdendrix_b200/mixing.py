"""Run-time check that the burn-in pool is mixed.

The pool's burn-in rule (which orientation of the matrix the rounds trade, how many rounds: DESIGN.md 3b) was derived on mutation-matrix shapes.  This check costs a few
hundred extra nulls and needs nothing but the public engine calls: chains of `rounds` nulls are started from the pool and the
number of ones each null shares with the OBSERVED matrix is compared between the first and the last null of the chains.  From
stationary start states the expected overlap does not depend on the position in the chain; from start states that still
remember the observed matrix it keeps falling (at C0 without burn-in: 1 285 ones shared after one round, 717 after four,
558 when stationary).  The reference needs no such check: its nulls are late states of one long chain
(src/generate_null_matrices.py:71-74)."""
from __future__ import annotations

import numpy as np

from . import bitpack


def trade_rows(feature_rows, n_samples: int):
    """The uploaded matrix in the trade orientation of the library (rows over the shorter dimension, src/curveball.py:10-11)."""
    fr = np.ascontiguousarray(feature_rows, dtype=np.uint32)
    F = fr.shape[0]
    return bitpack.pack_rows(bitpack.unpack_rows(fr, n_samples).T) if F >= n_samples else fr


def overlap_drift(engine, feature_rows, n_samples: int, seed: int = 1764, n_chains: int = 96, rounds: int = 4, null0: int = 1 << 40):
    """Returns dict(first, last, drift, sigma, z): mean ones shared with the observed matrix (feature_rows, as uploaded) by the
    first / last null of n_chains chains of `rounds` nulls (current burn-in and pair-group settings of the engine), their
    difference, its standard error and z = drift / sigma.  null0: where in the null index space the probe chains live (far
    from the nulls of a run; a multiple of 32 * rounds so that whole blocks of chains are used)."""
    null0 -= null0 % (32 * rounds)
    obs = trade_rows(feature_rows, n_samples)
    rows, _ = engine.curveball(seed, null0, n_chains * rounds, chain_len=rounds, want_applied=False)
    shared = bitpack.popcount_rows(rows & obs[None, :, :]).sum(axis=1).reshape(n_chains, rounds)
    first, last = shared[:, 0].astype(np.float64), shared[:, -1].astype(np.float64)
    d = first - last
    sigma = float(d.std(ddof=1) / np.sqrt(len(d))) if len(d) > 1 else float("inf")
    return {"first": float(first.mean()), "last": float(last.mean()), "drift": float(d.mean()), "sigma": sigma,
            "z": float(d.mean() / sigma) if sigma > 0 else 0.0}


def check(engine, feature_rows, n_samples: int, seed: int = 1764, n_chains: int = 96, rounds: int = 4, z_max: float = 4.0, log=None):
    """True when the pool looks stationary; otherwise logs / warns that the burn-in is too short for this matrix."""
    r = overlap_drift(engine, feature_rows, n_samples, seed, n_chains, rounds)
    ok = r["z"] <= z_max
    msg = ("burn-in check: nulls share %.1f ones with the observed matrix one round after their pool state and %.1f after %d rounds "
           "(drift %.1f +- %.1f, z = %.1f): %s" % (r["first"], r["last"], rounds, r["drift"], r["sigma"], r["z"],
                                                    "pool is stationary" if ok else "the pool still remembers the observed matrix — raise --burn-in"))
    if log:
        log(msg)
    elif not ok:
        import warnings
        warnings.warn(msg)
    return ok
