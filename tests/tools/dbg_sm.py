"""Debug driver for the persistent trade kernel: one small null against the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from superdendrix_b200 import bitpack
from superdendrix_b200.engine import Engine

shape = (40, 17)
rng = np.random.default_rng(shape[0])
mat = (rng.random(shape) < 0.3).astype(np.uint8)
S, F = shape
obs = bitpack.pack_rows(mat if F >= S else mat.T)
eng = Engine(0)
eng.upload_dense(mat)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cl = int(sys.argv[2]) if len(sys.argv) > 2 else 1
T = int(sys.argv[3]) if len(sys.argv) > 3 else 5 * min(shape)
rows, _ = eng.curveball(1764, 0, n, T, cl, want_applied=False)
want, _ = oracle.curveball_nulls(obs, 1764, 0, n, T, chain_len=cl)
print("equal:", np.array_equal(rows, want))
