"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the reference's ILP restated for scipy.optimize.milp (HiGHS).

The reference solves its model with Gurobi 9.1.2 (environment.yml:49), which is not installed here; HiGHS solves the
same model, so its optimum certifies the exhaustive scan (pin G4).

Model (src/superdendrix.py:294-313, objective build_obj :393-401), x_g = feature g selected, y_p = sample p covered:

    maximise   sum_{p: w_p > 0} 2 w_p y_p  -  sum_p |w_p| c_p ,      c_p = sum_{g in G(p)} x_g
    subject to sum_g x_g <= k                                        (:306)
               y_p <= c_p                       for every sample p   (:311)
               y_p >= x_g   for g in G(p)       when w_p < 0         (:313)

(:393-401 builds  sum_{pos} w y - (sum_{pos} w c - sum_{pos} w y) - sum_{non-pos} (-w) c , which is the line above.)
"""
import numpy as np
from scipy import sparse
from scipy.optimize import Bounds, LinearConstraint, milp


def ilp_optimum(dense, w, k, all_samples=True, time_limit=300.0):
    """dense: S x F 0/1, w: S weights (already sign-flipped), k: cardinality bound.
    all_samples=False drops y_p and its constraints for samples with w_p <= 0: their objective coefficient is zero and
    (:311, :313) only tie them to x, so the optimum is unchanged (used to keep the C0-size model small).
    Returns (objective, sorted feature indices)."""
    dense = np.asarray(dense) != 0
    S, F = dense.shape
    w = np.asarray(w, dtype=np.float64)
    ys = np.arange(S) if all_samples else np.flatnonzero(w > 0)
    ny = len(ys)
    ypos = {int(p): F + i for i, p in enumerate(ys)}
    n = F + ny
    c = np.zeros(n)
    c[:F] = (np.abs(w)[:, None] * dense).sum(0)               # + sum_p |w_p| c_p  (milp minimises)
    for p in ys:
        if w[p] > 0:
            c[ypos[int(p)]] = -2.0 * w[p]
    rows, cols, vals, lo, hi = [], [], [], [], []
    r = 0
    rows += [r] * F; cols += list(range(F)); vals += [1.0] * F; lo.append(-np.inf); hi.append(float(k)); r += 1      # :306
    for p in ys:
        g = np.flatnonzero(dense[p])
        rows += [r] * (len(g) + 1); cols += [ypos[int(p)]] + list(g); vals += [1.0] + [-1.0] * len(g)          # :311  y - c <= 0
        lo.append(-np.inf); hi.append(0.0); r += 1
        if w[p] < 0:
            for gg in g:                                                                                      # :313  x - y <= 0
                rows += [r, r]; cols += [int(gg), ypos[int(p)]]; vals += [1.0, -1.0]
                lo.append(-np.inf); hi.append(0.0); r += 1
    A = sparse.csr_matrix((vals, (rows, cols)), shape=(r, n))
    res = milp(c, constraints=LinearConstraint(A, np.array(lo), np.array(hi)), integrality=np.ones(n), bounds=Bounds(0, 1),
               options={"time_limit": time_limit, "mip_rel_gap": 0.0})
    if res.x is None:
        raise RuntimeError("HiGHS returned no solution: %s" % res.message)
    feats = [int(g) for g in np.flatnonzero(res.x[:F] > 0.5)]
    return -float(res.fun), feats, bool(res.status == 0)
