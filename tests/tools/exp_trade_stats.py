"""Trade statistics of the C0 workload (stationary start), for the cost model in profiles/r01_k_trade_ncu_summary.md:
row sizes, symmetric-difference sizes n, picks k, share of trades that involve a dense row.  CPU only."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from superdendrix_b200 import synth, bitpack

dense, _, _ = synth.feature_matrix(770, 400)
S, F = dense.shape
obs = bitpack.pack_rows(dense if F >= S else dense.T)
R = obs.shape[0]
L = max(S, F)
T = 5 * R
nulls, _ = oracle.curveball_nulls(obs, 5, 0, 40, T, chain_len=40)           # state after 40 rounds ~ stationary
state = bitpack.unpack_rows(nulls[-1], L).astype(bool)
sizes = state.sum(1)
print("rows", R, "row size: mean %.1f median %d p90 %d max %d; rows > 16: %d, > 32: %d, > 64: %d" % (
    sizes.mean(), np.median(sizes), np.percentile(sizes, 90), sizes.max(), (sizes > 16).sum(), (sizes > 32).sum(), (sizes > 64).sum()))
rng = np.random.default_rng(1)
ns, ks, cls = [], [], []
for _ in range(20000):
    a, b = rng.choice(R, 2, replace=False)
    A, B = state[a], state[b]
    la, lb = int((A & ~B).sum()), int((B & ~A).sum())
    ns.append(la + lb); ks.append(min(la, lb) if la and lb else 0)
    cls.append((sizes[a] > 16) + (sizes[b] > 16))
ns, ks, cls = np.array(ns), np.array(ks), np.array(cls)
print("n = |a^b|: mean %.1f median %d p90 %d p99 %d max %d;  n <= 32: %.1f%%  n <= 64: %.1f%%" % (
    ns.mean(), np.median(ns), np.percentile(ns, 90), np.percentile(ns, 99), ns.max(), 100 * (ns <= 32).mean(), 100 * (ns <= 64).mean()))
print("k picks: mean %.2f median %d p90 %d p99 %d max %d; skipped (k = 0): %.1f%%" % (
    ks.mean(), np.median(ks), np.percentile(ks, 90), np.percentile(ks, 99), ks.max(), 100 * (ks == 0).mean()))
for c, name in ((0, "list x list (both <= 16)"), (1, "one row > 16"), (2, "both rows > 16")):
    m = cls == c
    print("%-26s %5.1f%% of trades, mean n %.1f, mean k %.2f, share of sum(n) %.1f%%, share of sum(k) %.1f%%" % (
        name, 100 * m.mean(), ns[m].mean() if m.any() else 0, ks[m].mean() if m.any() else 0, 100 * ns[m].sum() / ns.sum(), 100 * ks[m].sum() / ks.sum()))
# lock-step cost of 8 trades per warp (max over the group) vs mean, after sorting by the plan's size classes
def lockstep(vals, width):
    v = np.sort(vals)[::-1]
    pad = (-len(v)) % width
    v = np.concatenate([v, np.zeros(pad, dtype=v.dtype)]).reshape(-1, width)
    return v.max(1).sum() / max(vals.sum() / width, 1e-9)
print("lock-step inflation of k (sorted into rounds): 8 per round %.2fx, 32 per round %.2fx" % (lockstep(ks[:47 * 40], 8), lockstep(ks[:47 * 40], 32)))
