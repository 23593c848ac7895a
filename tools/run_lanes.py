"""Lane kernel on the GPU: parity against the oracle on small cases, then throughput at C1 for a few chain lengths.
    python tools/run_lanes.py [n_nulls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SDX_LANE_MIN_CHAINS", "1")
import numpy as np

import oracle
from superdendrix_b200 import bitpack, defaults, synth
from superdendrix_b200.engine import Engine


def check(S, F, n, chain_len, null0, burn=0, pool=7, dens=None, seed=1764):
    if dens is None:
        dense, _, _ = synth.feature_matrix(S, F)
    else:
        dense = (np.random.default_rng(S * 1000 + F).random((S, F)) < dens).astype(np.uint8)
    S, F = dense.shape
    obs = bitpack.pack_rows(dense if F >= S else dense.T)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.set_pair_group(32)
        eng.set_burn_in(burn, pool)
        R = obs.shape[0]
        got, ap = eng.curveball(seed, null0, n, chain_len=chain_len)
        want, wap = oracle.curveball_nulls(obs, seed, null0, n, 5 * R, chain_len, burn=burn, pool_size=pool, pair_group=32, L=max(S, F))
        ok = np.array_equal(got, want) and np.array_equal(ap, wap)
        print("parity S=%d F=%d n=%d chain=%d null0=%d burn=%d: %s  (launches %d, trade %.2f ms)" % (
            S, F, n, chain_len, null0, burn, "OK" if ok else "MISMATCH", eng.timing()["trade_launches"], eng.timing()["trade_ms"]), flush=True)
        if not ok:
            bad = [i for i in range(n) if not np.array_equal(got[i], want[i])]
            print("   first bad nulls:", bad[:10], "applied", ap[:8], wap[:8])
        return ok


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    ok = True
    ok &= check(300, 120, 64, 1, 0)
    ok &= check(300, 120, 70, 3, 5)
    ok &= check(200, 60, 96, 2, 64, dens=0.03)
    ok &= check(70, 40, 40, 1, 0, dens=0.5)
    ok &= check(130, 50, 64, 2, 0, burn=2, dens=0.25)
    ok &= check(40, 90, 64, 1, 0, dens=0.15)
    ok &= check(770, 400, 64, 2, 0)
    if not ok:
        print("PARITY FAILED"); sys.exit(1)
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    w = synth.weights(S, 1)
    module = synth.pick_module(dense, 3)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
        res = {}
        for pg, cl in ((1, 16), (32, 16), (32, 4), (32, 3), (32, 2), (32, 1)):
            eng.set_pair_group(pg)
            for rep in range(3):
                t0 = time.perf_counter()
                r = eng.null_pvalue(1764, 0, n, [module], chain_len=cl)
                dt = time.perf_counter() - t0
            tm = eng.timing()
            res[(pg, cl)] = r["n_ge"][0]
            print("pair_group %2d chain %2d: %8.2f ms wall, trade %.2f ms, plan %.2f ms, launches %d -> %.2f M nulls/s  n_ge %d" % (
                pg, cl, dt * 1e3, tm["trade_ms"], tm["plan_ms"], tm["trade_launches"], n / tm["trade_ms"] / 1e3, r["n_ge"][0]), flush=True)


if __name__ == "__main__":
    main()
