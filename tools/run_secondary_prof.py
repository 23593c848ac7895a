"""One pass over the secondary kernels for ncu captures: C2 scan (k_pair_table, k_scan12, k_scan3), a C3 slice (k_plan_levels,
k_plan_sort, k_trade<SMEM=false>), a C4 slice (k_condperm*, k_ranksum).  python tools/run_secondary_prof.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth          # noqa: E402
from superdendrix_b200.engine import Engine           # noqa: E402

with Engine(0) as eng:
    dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
    eng.upload_matrix(bitpack.pack_rows(dense.T), 1000)
    eng.upload_weights(synth.weights(1000, 1, seed=3))
    for _ in range(2):
        r = eng.scan(0, 3, 8)
    print("C2 scan: %d subsets, %.3f ms" % (r["n_scored"], eng.timing()["total_ms"]))

    dense, _, _ = synth.feature_matrix(1000, 5000, seed=20)
    S, F = dense.shape
    n_prof = 500
    w = synth.weights(S, n_prof, seed=5)
    rng = np.random.default_rng(2)
    order = np.argsort(-dense.sum(0), kind="stable")[:60]
    mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
    eng.upload_matrix(bitpack.pack_rows(dense.T), S)
    eng.upload_weights(w)
    for _ in range(2):
        eng.null_pvalue(1764, 0, 1184, mods)
    t = eng.timing()
    print("C3 slice: 1184 nulls x 500 profiles, total %.2f ms (trade %.2f, plan %.2f)" % (t["total_ms"], t["trade_ms"], t["plan_ms"]))

    dense, _, _ = synth.feature_matrix(770, 400, seed=20)
    S, F = dense.shape
    n_prof = 2000
    w = synth.weights(S, n_prof, seed=11)
    eng.upload_matrix(bitpack.pack_rows(dense.T), S)
    eng.upload_weights(w)
    rng = np.random.default_rng(1)
    order = np.argsort(-dense.sum(0), kind="stable")[:40]
    mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
    pj = np.arange(n_prof, dtype=np.int32)
    for _ in range(2):
        eng.condperm_batch(mods, 10000, 1764, 0, pj)
    print("C4 slice: condperm %d jobs x 3 x 10000: %.2f ms" % (n_prof, eng.timing()["total_ms"]))
    for _ in range(2):
        eng.ranksum(mods, pj)
    print("C4 slice: ranksum %d jobs: %.3f ms" % (n_prof, eng.timing()["total_ms"]))
