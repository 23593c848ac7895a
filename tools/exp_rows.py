"""Experiment: how much of the Curveball time is due to the few dense rows?  (benchmarks only)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine

dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(770, 1)
order = np.argsort(-dense.sum(0), kind="stable")
with Engine(0) as eng:
    for name, keep in (("all rows", np.arange(dense.shape[1])), ("without the 8 densest", np.sort(order[8:])),
                       ("without the 16 densest", np.sort(order[16:])), ("without the 32 densest", np.sort(order[32:]))):
        d = dense[:, keep]
        eng.upload_matrix(bitpack.pack_rows(d.T), 770)
        eng.upload_weights(w)
        mod = np.array([[0, 1, 2]], dtype=np.int32)
        eng.null_pvalue(1, 0, 32768, mod)
        eng.null_pvalue(1, 0, 32768, mod)
        t = eng.timing()
        print("%-24s F=%d ones=%d max row=%d : trade %.2f ms plan %.2f ms per 32768 nulls -> %.2f M nulls/s" % (
            name, d.shape[1], d.sum(), d.sum(0).max(), t["trade_ms"], t["plan_ms"], 32768 / t["total_ms"] / 1e3))
