"""Runs the max_k null statistic on C0 (used for launch lists / ncu)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import SDX_STAT_MAXK, Engine
dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(770, 1)
module = synth.pick_module(dense, 3)[None, :]
with Engine(0) as eng:
    eng.upload_dense(dense); eng.upload_weights(w)
    for _ in range(2):
        r = eng.null_pvalue(1764, 0, 1000, module, stat=SDX_STAT_MAXK)
        print(eng.timing(), int(r["n_ge"][0]))
