"""Individual contributions of the features of a module
(``utils/compute_individual_contribution_SELECT.py:69-85``).

* ``pair_contributions``: for feature pairs, ``sum_{s in g_i} w_s / sum_{s in g1 ∪ g2} w_s`` — the two ratios the
  reference script appends to each result row (it uses the raw profile, blanks as 0.0).
* ``leave_one_out``: ``W(M) - W(M \\ {g})`` for every feature of a module (the form the genome-wide screen uses).
"""
from __future__ import annotations

import numpy as np


def pair_contributions(eng, pairs, profile_of_pair=None):
    """pairs: (n, 2) feature indices.  Returns (m1_contribution, m2_contribution), each (n,)."""
    pairs = np.atleast_2d(np.asarray(pairs, dtype=np.int32))
    n = len(pairs)
    prof = np.zeros(n, dtype=np.int32) if profile_of_pair is None else np.asarray(profile_of_pair, dtype=np.int32)
    total = eng.union_sums(pairs, prof)
    m1 = eng.union_sums(pairs[:, :1], prof)
    m2 = eng.union_sums(pairs[:, 1:2], prof)
    with np.errstate(divide="ignore", invalid="ignore"):
        return m1 / total, m2 / total


def leave_one_out(eng, modules, profile_of_job=None):
    """modules: (n, k) feature indices, -1 padded.  Returns (n, k): W(M) - W(M without that feature); NaN for empty slots."""
    mods = np.atleast_2d(np.asarray(modules, dtype=np.int32))
    n, k = mods.shape
    prof = np.zeros(n, dtype=np.int32) if profile_of_job is None else np.asarray(profile_of_job, dtype=np.int32)
    full = eng.score_sets(mods, prof)["W"]
    out = np.full((n, k), np.nan)
    for j in range(k):
        reduced = mods.copy()
        reduced[:, j] = -1
        wj = eng.score_sets(reduced, prof)["W"]
        out[:, j] = np.where(mods[:, j] >= 0, full - wj, np.nan)
    return out
