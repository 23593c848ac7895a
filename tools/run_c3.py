"""C3 slice (1000 x 5000 matrix, rows in global memory / L2): nulls per second of the trade kernel alone and with 500 profiles."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine
dense, _, _ = synth.feature_matrix(1000, 5000, seed=20)
S, F = dense.shape
n_prof = int(sys.argv[1]) if len(sys.argv) > 1 else 500
w = synth.weights(S, n_prof, seed=5)
rng = np.random.default_rng(2)
order = np.argsort(-dense.sum(0), kind="stable")[:60]
mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
with Engine(0) as eng:
    eng.upload_matrix(bitpack.pack_rows(dense.T), S)
    eng.upload_weights(w)
    eng.null_pvalue(1764, 0, 64, mods)
    for n in (1024, 4096):
        r = eng.null_pvalue(1764, 0, n, mods)
        t = eng.timing()
        print("n=%d profiles=%d total %.2f ms trade %.2f ms plan %.2f ms -> %.0f nulls/s  n_ge mean %.3f" % (
            n, n_prof, t["total_ms"], t["trade_ms"], t["plan_ms"], n / (t["total_ms"] / 1e3), float(np.mean(r["n_ge"]))))
