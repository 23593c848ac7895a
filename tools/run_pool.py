"""Cost of the burn-in pool on the bench matrix: plan / trade / total ms of sdx_pool_begin + build + commit for a few
(burn rounds, states) settings, and the run-time mixing check of the result.  python tools/run_pool.py"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, defaults, mixing, synth      # noqa: E402
from superdendrix_b200.engine import Engine                          # noqa: E402

dense, _, _ = synth.feature_matrix(770, 400, seed=20)
S, F = dense.shape
rows = bitpack.pack_rows(dense.T)
with Engine(0) as eng:
    eng.upload_matrix(rows, S)
    eng.set_pair_group(defaults.PAIR_GROUP)
    for burn, states in ((defaults.BURN_ROUNDS, defaults.POOL_SIZE), (defaults.BURN_ROUNDS, 184), (defaults.BURN_ROUNDS, 736), (2, 1472), (6, 1472), (1, 32)):
        eng.set_burn_in(burn, defaults.POOL_SIZE)
        best = None
        for rep in range(4):
            t0 = time.perf_counter()
            eng.pool_begin(1764 + rep)
            t1 = time.perf_counter()
            eng.pool_build(0, states)
            t2 = time.perf_counter()
            tm = eng.timing()
            cur = (1e3 * (t2 - t0), 1e3 * (t1 - t0), tm["plan_ms"], tm["trade_ms"], tm["total_ms"])
            best = cur if best is None or cur[0] < best[0] else best
        print("burn %d, %4d states: wall %.2f ms (begin %.2f), k_plan %.2f ms, k_trade %.2f ms, device total %.2f ms" % ((burn, states) + best), flush=True)
    eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
    print(mixing.overlap_drift(eng, rows, S))
    eng.set_burn_in(1, defaults.POOL_SIZE)
    print("burn 1:", mixing.overlap_drift(eng, rows, S))
