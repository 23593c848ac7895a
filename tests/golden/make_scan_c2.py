#!/usr/bin/env python
"""tests/golden/scan_c2.npz: the 16 best subsets of the FULL C2 scan (BASELINE configs[2]: k <= 3 over the 1 000 x ~2 000
synthetic matrix, 1.27e9 subsets) computed once by the CPU oracle — oracle/sdx_oracle.c scores every subset with the plain
definition of SuperW (src/superdendrix.py:387-391), no inclusion-exclusion, so it shares no arithmetic with the GPU scan.
The enumeration is split by smallest feature index over the host cores (about 15 core-hours / cores); the merge order is
the scan's: W descending, then size, then lexicographic.

    python tests/golden/make_scan_c2.py [processes]
"""
import multiprocessing as mp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from superdendrix_b200 import bitpack, dist, synth  # noqa: E402

TOPN = 16


def _inputs():
    dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
    w = synth.weights(1000, 1, seed=3)[0]
    return bitpack.pack_rows(dense.T), dense.shape[0], w


def _job(rng):
    rows, S, w = _inputs()
    W, F, C, n = oracle.scan_range(rows, S, w, 3, TOPN, rng[0], rng[1])
    return {"W": W, "features": F, "coverage": C, "n": n}


if __name__ == "__main__":
    nproc = int(sys.argv[1]) if len(sys.argv) > 1 else max(1, (os.cpu_count() or 2) - 1)
    rows, S, w = _inputs()
    F = rows.shape[0]
    ranges = [r for r in dist.scan_feature_ranges(F, 3, nproc * 6) if r[1] > r[0]]
    with mp.get_context("fork").Pool(nproc) as pool:
        parts = pool.map(_job, ranges, chunksize=1)
    merged = dist.merge_hits(parts, TOPN)
    n = sum(p["n"] for p in parts)
    assert n == F + F * (F - 1) // 2 + F * (F - 1) * (F - 2) // 6
    np.savez_compressed(os.path.join(HERE, "scan_c2.npz"), W=merged["W"], features=merged["features"], coverage=merged["coverage"],
                        n_scored=n, F=F, S=S)
    print("scan_c2.npz:", n, "subsets; best", merged["features"][0], merged["W"][0])
