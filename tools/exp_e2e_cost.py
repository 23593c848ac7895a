"""Experiment: where the end-to-end step spends its time (benchmarks only)."""
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
    eng.set_burn_in(6, 1480)
    for it in range(4):
        t0 = time.perf_counter(); eng.upload_matrix(rows, 770)
        t1 = time.perf_counter(); eng.upload_weights(w)
        t2 = time.perf_counter(); eng.null_pvalue(1764, it * 100000, 100000, module, chain_len=16)
        t3 = time.perf_counter()
        print("upload_matrix %.2f ms, upload_weights %.2f ms, null_pvalue %.2f ms (lib total %.2f)" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), eng.timing()["total_ms"]))
