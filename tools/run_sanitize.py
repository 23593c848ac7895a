"""Small cases for compute-sanitizer (memcheck / racecheck): the lane kernel and k_trade on a 300 x 120 matrix with dense rows,
from a burn-in pool, with fused scoring.  python tools/run_sanitize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDX_LANE_MIN_CHAINS", "1")
import numpy as np

from superdendrix_b200 import synth
from superdendrix_b200.engine import Engine

dense, _, _ = synth.feature_matrix(300, 120)
w = synth.weights(dense.shape[0], 2, seed=4)
order = np.argsort(-dense.sum(0), kind="stable")
mods = np.array([[order[0], order[3], order[50]], [order[1], order[70], -1]], dtype=np.int32)
for pg in (32, 1):
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        eng.set_pair_group(pg)
        eng.set_burn_in(2, 32)
        r = eng.null_pvalue(5, 0, 64, mods, chain_len=1, want_scores=True)
        rows, ap = eng.curveball(5, 64, 40, chain_len=2)
        print("pair_group", pg, "n_ge", r["n_ge"].tolist(), "plan_ms", eng.timing()["plan_ms"], "applied", int(ap.sum()))
