"""A small pass over every kernel of libsdx.so (used under compute-sanitizer)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import SDX_STAT_MAXK, Engine

with Engine(0) as eng:
    for S, F in ((770, 400), (60, 300), (130, 37)):
        dense, _, _ = synth.feature_matrix(S, F, seed=S)
        w = synth.weights(S, 2, seed=F)
        eng.upload_matrix(bitpack.pack_rows(dense.T), S)
        eng.upload_weights(w)
        mods = np.array([[0, 1, 2], [3, 4, -1]], dtype=np.int32)
        rows, ap = eng.curveball(5, 0, 12, chain_len=4)
        r = eng.null_pvalue(5, 0, 40, mods, want_scores=True)
        if F <= 60:
            eng.null_pvalue(5, 0, 6, mods, stat=SDX_STAT_MAXK)
        eng.scan(0, 3, 8)
        eng.score_sets(mods, np.array([0, 1], dtype=np.int32))
        eng.feature_scores(1)
        eng.condperm_batch(mods, 200, 3, 0, np.array([0, 1], dtype=np.int32))
        eng.ranksum(mods, np.array([0, 1], dtype=np.int32))
        print(S, F, "ok", int(ap.sum()), r["n_ge"].tolist())
