"""Full-matrix vs restricted-matrix randomisation: the reference randomises the FULL feature matrix
(src/generate_null_matrices.py:53) and restricts every null to the profiled samples when it loads it (src/superdendrix.py:186);
randomising the matrix already restricted to the profiled samples is a different null model.  For profiles that cover 100 %, 90 %,
70 % and 50 % of the C0 samples: permutation p-value and mean / sd of the null statistic of the bench module under both scopes
(20 000 nulls each, shipped defaults), with the z-score of the difference of the p-values.
    python tools/exp_null_scope.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from superdendrix_b200 import bitpack, defaults, synth
from superdendrix_b200.engine import Engine

dense, _, _ = synth.feature_matrix(770, 400, seed=20)
S, F = dense.shape
w = synth.weights(S, 1, seed=3)[0]
module = synth.pick_module(dense, 3)
n = 20000
rng = np.random.default_rng(5)
with Engine(0) as eng:
    for frac in (1.0, 0.9, 0.7, 0.5):
        keep = np.ones(S, dtype=bool) if frac == 1.0 else rng.random(S) < frac
        out = {}
        # full: the whole matrix is randomised, the profile is a mask
        eng.upload_dense(dense)
        eng.upload_weights(np.where(keep, w, 0.0)[None, :], keep[None, :])
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
        r = eng.null_pvalue(1764, 0, n, module[None, :].astype(np.int32), chain_len=1, want_scores=True)
        out["full"] = (r["n_ge"][0] / n, r["null_scores"][0].mean(), r["null_scores"][0].std(), r["z_obs"][0])
        # restricted: the matrix of the profiled samples is randomised (features without a carrier among them dropped, src/i_o.py:62-67)
        sub = dense[keep]
        cols = sub.sum(0) > 0
        idx = np.cumsum(cols) - 1
        eng.upload_dense(sub[:, cols])
        eng.upload_weights(w[keep][None, :])
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
        r = eng.null_pvalue(1764, 0, n, idx[module][None, :].astype(np.int32), chain_len=1, want_scores=True)
        out["restricted"] = (r["n_ge"][0] / n, r["null_scores"][0].mean(), r["null_scores"][0].std(), r["z_obs"][0])
        pf, pr = out["full"][0], out["restricted"][0]
        infl = 1 + (n / defaults.POOL_SIZE - 1) * 0.31                     # nulls that share a pool state
        se = np.sqrt((pf * (1 - pf) + pr * (1 - pr)) / n * infl)
        print("profile covers %3.0f %% of the samples (%d): full p %.4f mean %.2f sd %.2f | restricted p %.4f mean %.2f sd %.2f | z_obs %.2f / %.2f | dp %.4f = %.1f sigma" % (
            100 * frac, keep.sum(), pf, out["full"][1], out["full"][2], pr, out["restricted"][1], out["restricted"][2], out["full"][3], out["restricted"][3],
            pf - pr, (pf - pr) / se), flush=True)
