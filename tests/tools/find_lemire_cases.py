"""Finds null indices (seed 4242, two disjoint rows of `half` bits, one trade per null) whose picks enter the rejection
zone of Lemire's method / re-draw, by scanning the oracle.  Used to pin tests/test_gpu_parity.py::LEMIRE_CASES."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle
from superdendrix_b200 import bitpack


def search(half, max_nulls, step=1_000_000, want=2):
    L = max(2 * half, 64)
    d = np.zeros((2, L), np.uint8); d[0, :half] = 1; d[1, half:2 * half] = 1
    obs = bitpack.pack_rows(d)
    hits, lo = [], 0
    while lo < max_nulls and len(hits) < want:
        oracle.deal_stats(reset=True)
        oracle.curveball_nulls(obs, 4242, lo, step, 1, 1, want_matrices=False)
        if oracle.deal_stats()[0]:
            for s0 in range(lo, lo + step, 2000):
                oracle.deal_stats(reset=True)
                oracle.curveball_nulls(obs, 4242, s0, 2000, 1, 1, want_matrices=False)
                if oracle.deal_stats()[0]:
                    for i in range(s0, s0 + 2000):
                        oracle.deal_stats(reset=True)
                        oracle.curveball_nulls(obs, 4242, i, 1, 1, 1, want_matrices=False)
                        z, r = oracle.deal_stats()
                        if z:
                            hits.append((i, z, r))
        lo += step
    return hits


if __name__ == "__main__":
    for half, lim in ((400, 400_000), (32, 20_000_000), (16, 60_000_000)):
        print(half, search(half, lim))
