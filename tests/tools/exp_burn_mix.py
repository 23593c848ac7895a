"""Does a burn-in that favours pairs of dense rows mix faster?    The pair distribution of a Curveball chain may
be any fixed distribution (row sizes are invariants), the stationary distribution stays uniform.  C0, pure Python sets."""
import random, sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))))
from superdendrix_b200 import synth

dense, _, _ = synth.feature_matrix(770, 400)
S, F = dense.shape
R = F
rows0 = [set(np.flatnonzero(dense[:, g]).tolist()) for g in range(F)]
sizes = np.array([len(r) for r in rows0])
order = np.argsort(-sizes, kind="stable")
T = 5 * R

def trade(rows, a, b, rnd):
    A, B = rows[a], rows[b]
    ab = A & B
    la, lb = len(A) - len(ab), len(B) - len(ab)
    if la == 0 or lb == 0:
        return
    tot = sorted((A | B) - ab)
    rnd.shuffle(tot)
    rows[a] = ab | set(tot[:la])
    rows[b] = ab | set(tot[la:])

def run(frac_dense, K, rounds, chains, seed):
    rnd = random.Random(seed)
    top = [int(x) for x in order[:K]]
    ov_all, ov_top, ov_mid = np.zeros(rounds), np.zeros(rounds), np.zeros(rounds)
    mid = [int(x) for x in order[8:K]]
    for c in range(chains):
        rows = [set(r) for r in rows0]
        for r in range(rounds):
            for t in range(T):
                if rnd.random() < frac_dense:
                    a, b = rnd.sample(top, 2)
                else:
                    a, b = rnd.sample(range(R), 2)
                trade(rows, a, b, rnd)
            ov_all[r] += sum(len(rows[g] & rows0[g]) for g in range(R))
            ov_top[r] += len(rows[order[0]] & rows0[order[0]])
            ov_mid[r] += sum(len(rows[g] & rows0[g]) for g in mid)
    return ov_all / chains, ov_top / chains, ov_mid / chains

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for frac, K in ((0.0, 40), (0.25, 40), (0.5, 40), (0.5, 16)):
    a, t, m = run(frac, K, rounds, 6, 1)
    print("dense-pair fraction %.2f (top %d rows): whole %s" % (frac, K, [int(x) for x in a]))
    print("                                     densest %s" % [int(x) for x in t])
    print("                               rows 9..%d sum %s" % (K, [int(x) for x in m]))
