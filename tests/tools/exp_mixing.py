"""Mixing measurement behind DESIGN.md §3b: ones shared with the observed matrix after r rounds of Curveball
(5*min(S,F) trades each) from the observed matrix, C0 shape, canonical streams, CPU oracle.  Test infrastructure only."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from superdendrix_b200 import synth, bitpack

dense, _, _ = synth.feature_matrix(770, 400)
S, F = dense.shape
obs = bitpack.pack_rows(dense if F >= S else dense.T)
CL, NCH = 48, int(sys.argv[1]) if len(sys.argv) > 1 else 60
t0 = time.time()
nulls, _ = oracle.curveball_nulls(obs, 5, 0, CL * NCH, 5 * min(S, F), chain_len=CL)
L = max(S, F)
g = int(np.argmax(dense.sum(0)))
tot, totg = np.zeros(CL), np.zeros(CL)
for i, m in enumerate(nulls):
    d = bitpack.unpack_rows(m, L)
    d = d if F >= S else d.T
    tot[i % CL] += (d * dense).sum()
    totg[i % CL] += (d[:, g] * dense[:, g]).sum()
rs = [1, 2, 4, 8, 12, 16, 24, 32, 40, 48]
print("chains", NCH, "seconds %.1f" % (time.time() - t0), "ones", int(dense.sum()), "densest feature", int(dense[:, g].sum()))
print("rounds           ", rs)
print("whole matrix     ", [float(round(tot[r - 1] / NCH, 1)) for r in rs])
print("densest feature  ", [float(round(totg[r - 1] / NCH, 1)) for r in rs])
