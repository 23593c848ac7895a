"""Randomised parity sweep of the burn-in pool: GPU nulls started from pool states vs the oracle, over random shapes, densities and
degenerate matrices (empty, full rows / columns, two rows), pair_group 1 and 32.  python tools/fuzz_burn.py [n_cases] [seed]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle                                                   # noqa: E402
from superdendrix_b200 import bitpack                           # noqa: E402
from superdendrix_b200.engine import Engine                     # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
bad = 0
seen = set()
with Engine(0) as eng:
    for case in range(n_cases):
        S, F = int(rng.integers(2, 200)), int(rng.integers(2, 120))
        kind = case % 8
        dens = float(rng.choice([0.01, 0.05, 0.15, 0.4]))
        d = (rng.random((S, F)) < dens).astype(np.uint8)
        if kind == 1:
            d[:, int(rng.integers(F))] = rng.random(S) < 0.8            # a frequent feature
        elif kind == 2:
            d[int(rng.integers(S)), :] = rng.random(F) < 0.9            # a hypermutated sample
        elif kind == 3:
            d[:, 0] = 1; d[0, :] = 1                                     # a full row and a full column
        elif kind == 4:
            d[:] = 0                                                     # empty matrix
        elif kind == 5:
            S, F = (2, int(rng.integers(2, 70))) if case % 16 < 8 else (int(rng.integers(2, 70)), 2)
            d = (rng.random((S, F)) < 0.5).astype(np.uint8)
        elif kind == 6:
            S, F = 32 * int(rng.integers(1, 4)), 32 * int(rng.integers(1, 4))   # word-aligned sizes
            d = (rng.random((S, F)) < dens).astype(np.uint8)
        S, F = d.shape
        obs = bitpack.pack_rows(d if F >= S else d.T)
        R, L = obs.shape[0], max(S, F)
        plan = oracle.burn_plan(obs, L)
        seen.add((plan["alt"], plan["dense_rows"] > 0))
        eng.upload_dense(d)
        for pg, burn, pool, chain_len, null0, n in ((1, 2, 5, 2, 1, 9), (32, 1, 40, 1, 32, 40), (32, 3, 33, 3, 0, 99)):
            eng.set_pair_group(pg)
            eng.set_burn_in(burn, pool)
            want, wap = oracle.curveball_nulls(obs, 1000 + case, null0, n, 5 * R, chain_len, burn=burn, pool_size=pool, pair_group=pg, L=L)
            got, gap = eng.curveball(1000 + case, null0, n, -1, chain_len)
            ok = np.array_equal(got, want) and np.array_equal(gap, wap)
            if not ok:
                bad += 1
                print("MISMATCH case %d kind %d S=%d F=%d dens=%.2f plan=%s pg=%d burn=%d pool=%d" % (case, kind, S, F, dens, plan, pg, burn, pool), flush=True)
        eng.set_burn_in(0)
        eng.set_pair_group(1)
print("cases %d, mismatches %d, burn-plan branches seen (transpose, dense-favouring): %s" % (n_cases, bad, sorted(seen)))
sys.exit(1 if bad else 0)
