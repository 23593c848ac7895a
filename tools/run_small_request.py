"""Latency of small requests on the bench matrix (what one rank of N sees when 100 000 nulls are sharded): python tools/run_small_request.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import defaults, synth      # noqa: E402
from superdendrix_b200.engine import Engine         # noqa: E402

dense, _, _ = synth.feature_matrix(770, 400, seed=20)
w = synth.weights(dense.shape[0], 1, seed=3)
module = synth.pick_module(dense, 3)
with Engine(0) as eng:
    eng.upload_dense(dense)
    eng.upload_weights(w)
    eng.set_pair_group(defaults.PAIR_GROUP)
    eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
    for n in (100000, 50000, 25000, 12500, 6250, 1024):
        best = None
        for _ in range(4):
            r = eng.null_pvalue(1764, 0, n, [module], chain_len=defaults.CHAIN_LEN)
            t = eng.timing()
            best = t if best is None or t["total_ms"] < best["total_ms"] else best
        print("%6d nulls: total %.2f ms (plan %.2f, trade %.2f), %d launches, n_ge %d" % (n, best["total_ms"], best["plan_ms"], best["trade_ms"], best["launches"], r["n_ge"][0]), flush=True)
