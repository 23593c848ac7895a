"""Experiment: cost of the fused scoring tail of k_trade (benchmarks only)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine

dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(770, 1)
order = np.argsort(-dense.sum(0), kind="stable")
with Engine(0) as eng:
    eng.upload_matrix(bitpack.pack_rows(dense.T), 770)
    eng.upload_weights(w)
    n = 32768
    for _ in range(2):
        eng.curveball(1, 0, n, want_rows=False, want_applied=False)
    print("generate only            : trade %.2f ms" % eng.timing()["trade_ms"])
    for name, mod in (("module = ranks 2-4 (bench)", np.sort(order[1:4])), ("module = 3 densest", np.sort(order[:3])),
                      ("module = 3 sparse", np.sort(order[200:203]))):
        for _ in range(2):
            eng.null_pvalue(1, 0, n, mod[None, :].astype(np.int32))
        print("%-28s: trade %.2f ms  (carriers %s)" % (name, eng.timing()["trade_ms"], dense[:, mod].sum(0).tolist()))
