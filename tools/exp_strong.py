"""What one rank of an 8-GPU strong-scaled C1 run does: 12 500 nulls (391 groups) on the lane kernel vs k_trade, and
the burn-in of 184 pool states.  python tools/exp_strong.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from superdendrix_b200 import defaults, synth
from superdendrix_b200.engine import Engine

dense, _, _ = synth.feature_matrix(770, 400)
w = synth.weights(dense.shape[0], 1)
module = synth.pick_module(dense, 3)[None, :].astype(np.int32)
for lanes in ("1", "0"):
    os.environ["SDX_LANES"] = lanes
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
        for n in (3125, 6250, 12500, 25000, 50000):
            for _ in range(3):
                t0 = time.perf_counter()
                r = eng.null_pvalue(1764, 0, n, module, chain_len=1)
                dt = time.perf_counter() - t0
            tm = eng.timing()
            print("lanes=%s n=%6d: wall %.2f ms, trade %.2f ms, plan %.2f ms, total %.2f ms, n_ge %d" % (lanes, n, dt * 1e3, tm["trade_ms"], tm["plan_ms"], tm["total_ms"], r["n_ge"][0]), flush=True)
        for cnt in (1472, 736, 368, 184):
            ts = []
            for _ in range(3):
                eng.pool_begin(1764)
                t0 = time.perf_counter()
                eng.pool_build(0, cnt)
                ts.append(time.perf_counter() - t0)
            print("lanes=%s pool_build of %4d states: %.2f ms" % (lanes, cnt, min(ts) * 1e3), flush=True)
        eng.pool_build(0, defaults.POOL_SIZE); eng.pool_commit()
