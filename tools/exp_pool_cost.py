"""Experiment: cost of building the burn-in pool and of chained launches (benchmarks only)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine
dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(770, 1)
module = synth.pick_module(dense, 3)[None, :]
rows = bitpack.pack_rows(dense.T)
with Engine(0) as eng:
    for L, K, B in ((32, 1024, 32), (16, 1024, 32), (16, 2048, 32), (8, 1024, 32), (1, 1024, 0)):
        eng.upload_matrix(rows, 770); eng.upload_weights(w)
        eng.set_burn_in(B, K)
        t0 = time.perf_counter(); eng.null_pvalue(1764, 0, 100000, module, chain_len=L); t1 = time.perf_counter()
        first = eng.timing()
        eng.null_pvalue(1764, 100000, 100000, module, chain_len=L); t2 = time.perf_counter()
        second = eng.timing()
        print("L=%d K=%d B=%d: first call %.1f ms wall (incl. pool), second %.1f ms wall; lib total_ms %.1f / %.1f, trade %.1f / %.1f, launches %d / %d" % (
            L, K, B, 1e3 * (t1 - t0), 1e3 * (t2 - t1), first["total_ms"], second["total_ms"], first["trade_ms"], second["trade_ms"], first["launches"], second["launches"]))
