"""Runs the C2 exhaustive scan (2 000 features x 1 000 samples, k <= 3) a few times; used for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth          # noqa: E402
from superdendrix_b200.engine import Engine           # noqa: E402

dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
w = synth.weights(1000, 1, seed=3)
with Engine(0) as eng:
    eng.upload_matrix(bitpack.pack_rows(dense.T), 1000)
    eng.upload_weights(w)
    for _ in range(3):
        r = eng.scan(0, 3, 8)
        print("%d subsets in %.3f ms (kernels %.3f ms), best W %.6f %s" % (r["n_scored"], eng.timing()["total_ms"], eng.timing()["score_ms"],
                                                                        r["W"][0], r["features"][0].tolist()))
