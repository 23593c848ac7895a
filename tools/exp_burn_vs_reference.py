"""How many burn-in rounds does the pool need on the bench workload?  For burn = 1..8: 100 000 GPU nulls with the shipped
layout (one round from a pool state, pair groups) against the stationary part of the reference's own chain
(tests/golden/nulldist_c0.npz): p-value, mean, sd and the z-scores of the differences (cluster / autocorrelation inflated).
    python tools/exp_burn_vs_reference.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from superdendrix_b200 import defaults, synth
from superdendrix_b200.engine import Engine

z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "nulldist_c0.npz"))
ref, z_obs = z["scores"][200:], float(z["z_obs"])
x = ref - ref.mean()
v = float(np.dot(x, x)) / len(x)
tau = 1.0
for k in range(1, 200):
    r = float(np.dot(x[:-k], x[k:])) / (len(x) - k) / v
    if r <= 0:
        break
    tau += 2 * r
dense, _, _ = synth.feature_matrix(770, 400, seed=20)
w = synth.weights(dense.shape[0], 1, seed=3)
module = synth.pick_module(dense, 3)[None, :].astype(np.int32)
n, P = 100_000, defaults.POOL_SIZE
print("reference chain: %d nulls, mean %.2f sd %.2f p %.4f tau %.2f" % (len(ref), ref.mean(), ref.std(), (ref >= z_obs).mean(), tau))
with Engine(0) as eng:
    eng.upload_dense(dense)
    eng.upload_weights(w)
    eng.set_pair_group(defaults.PAIR_GROUP)
    for burn in (1, 2, 3, 4, 5, 6, 8):
        eng.set_burn_in(burn, P)
        sc = eng.null_pvalue(1764, 0, n, module, chain_len=1, want_scores=True)["null_scores"][0]
        per = sc[: (n // P) * P].reshape(-1, P)
        se_m = np.sqrt(ref.var() * tau / len(ref) + per.mean(axis=0).var(ddof=1) / P)
        out = ["burn %d: mean %.2f (z %+.2f) sd %.2f" % (burn, sc.mean(), (sc.mean() - ref.mean()) / se_m, sc.std())]
        for thr in [z_obs] + [float(np.quantile(ref, q)) for q in (0.1, 0.5, 0.9)]:
            pr, pn = (ref >= thr).mean(), (sc >= thr).mean()
            se = np.sqrt(pr * (1 - pr) * tau / len(ref) + (per >= thr).mean(axis=0).var(ddof=1) / P)
            out.append("p(>=%.0f) %.4f vs %.4f (z %+.2f)" % (thr, pn, pr, (pn - pr) / se))
        print("; ".join(out), flush=True)
