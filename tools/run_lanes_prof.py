"""One lane-kernel launch for ncu: python tools/run_lanes_prof.py [n_nulls] [chain_len]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import defaults, synth
from superdendrix_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
cl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(dense.shape[0], 1)
module = synth.pick_module(dense, 3)
with Engine(0) as eng:
    eng.upload_dense(dense)
    eng.upload_weights(w)
    eng.set_pair_group(32)
    eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
    for _ in range(2):
        r = eng.null_pvalue(1764, 0, n, [module], chain_len=cl)
    print(r["n_ge"], eng.timing())
