"""Mixing after b burn-in rounds, measured on the GPU path: ones shared with the
observed matrix by nulls that are ONE ordinary round away from a pool state.  C0 shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from superdendrix_b200 import bitpack, synth
from superdendrix_b200.engine import Engine

dense, _, _ = synth.feature_matrix(770, 400)
S, F = dense.shape
g = int(np.argmax(dense.sum(0)))
mid = np.argsort(-dense.sum(0), kind="stable")[8:40]
with Engine(0) as eng:
    eng.upload_dense(dense)
    R, L, wpr, ras = eng.trade_shape()
    for burn in (0, 1, 2, 3, 4, 6, 8, 12, 32):
        eng.set_burn_in(burn, 400)
        rows, _ = eng.curveball(5, 0, 400, -1, 1, want_applied=False)
        d = np.stack([bitpack.unpack_rows(r, L) for r in rows]).astype(np.int64)      # n x F x S
        whole = (d * dense.T[None]).sum((1, 2)).mean()
        top = (d[:, g, :] * dense[:, g][None]).sum(1).mean()
        m = (d[:, mid, :] * dense.T[mid][None]).sum((1, 2)).mean()
        print("burn %2d: ones shared with the observed matrix %7.1f   densest feature %6.1f   rows 9..40 %6.1f" % (burn, whole, top, m))
