"""TEST INFRASTRUCTURE: an object with the Engine methods that superdendrix_b200.dist calls, computed by the
oracle on the CPU.  It lets the world_size-2 gloo tests exercise the sharding / reduction logic without a GPU."""
import numpy as np

import oracle
from superdendrix_b200 import bitpack


class OracleEngine:
    def __init__(self, dense, w):
        self.dense = np.asarray(dense)
        self.S, self.F = self.dense.shape
        self.w = np.atleast_2d(np.asarray(w, dtype=np.float64))
        self.P = len(self.w)
        self.frows = bitpack.pack_rows(self.dense.T)
        self.ras = self.F >= self.S
        self.trows = bitpack.pack_rows(self.dense if self.ras else self.dense.T)
        self.burn, self.pool, self.pair_group = 0, 1, 1

    def set_burn_in(self, burn_rounds, pool_size=1472):
        self.burn, self.pool = int(burn_rounds), int(pool_size)

    def set_pair_group(self, pair_group):
        self.pair_group = int(pair_group)

    def null_pvalue(self, seed, null0, n_nulls, modules, n_trades=-1, chain_len=1, stat=0, z_obs=None, want_scores=False):
        modules = np.atleast_2d(modules)
        R = self.trows.shape[0]
        T = 5 * R if n_trades < 0 else n_trades
        z = np.array([oracle.score_set(self.frows, self.S, m[m >= 0], self.w[p])[0] for p, m in enumerate(modules)])
        zo = z if z_obs is None else np.asarray(z_obs)
        n_ge = np.zeros(self.P, dtype=np.int64)
        if n_nulls:
            nulls, _ = oracle.curveball_nulls(self.trows, seed, null0, n_nulls, T, chain_len, burn=self.burn, pool_size=self.pool,
                                              pair_group=self.pair_group, L=max(self.S, self.F))
            for m in nulls:
                fr = bitpack.pack_rows(bitpack.unpack_rows(m, max(self.S, self.F)).T) if self.ras else m
                for p, mod in enumerate(modules):
                    n_ge[p] += oracle.score_set(fr, self.S, mod[mod >= 0], self.w[p])[0] >= zo[p]
        return {"n_ge": n_ge, "z_obs": z, "null_scores": None}

    def scan(self, profile=0, k=3, topn=8, g_lo=0, g_hi=None):
        g_hi = self.F if g_hi is None else g_hi
        W, Fe, C, n = oracle.scan(self.frows, self.S, self.w[profile], k, topn=4096)
        keep = (Fe[:, 0] >= g_lo) & (Fe[:, 0] < g_hi)
        W, Fe, C = W[keep][:topn], Fe[keep][:topn], C[keep][:topn]
        pad = topn - len(W)
        r = self.F - 1 - np.arange(g_lo, g_hi)
        cnt = len(r) + (r.sum() if k >= 2 else 0) + ((r * (r - 1) // 2).sum() if k >= 3 else 0)
        return {"W": np.concatenate([W, np.full(pad, -np.inf)]),
                "features": np.concatenate([Fe, np.full((pad, 3), -1, dtype=np.int32)]).astype(np.int32),
                "coverage": np.concatenate([C, np.zeros(pad, dtype=np.int32)]).astype(np.int32), "n_scored": int(cnt)}

    def condperm_batch(self, modules, n_perm, seed, perm_id0=0, profile_of_job=None):
        counts, real = [], []
        for j, m in enumerate(np.atleast_2d(modules)):
            p = 0 if profile_of_job is None else int(profile_of_job[j])
            valid = m[m >= 0]
            c, r = oracle.condperm(self.frows, self.S, valid, self.w[p], np.arange(self.S), seed, perm_id0, n_perm)
            pad = len(m) - len(valid)
            counts.append(np.concatenate([c, np.full(pad, -1, dtype=np.int64)]))
            real.append(np.concatenate([r, np.zeros(pad)]))
        return np.array(counts, dtype=np.int64), np.array(real)

    def ranksum(self, modules, profile_of_job=None):
        out = {"z": [], "p": [], "twice_rank_sum": [], "n1": []}
        for j, m in enumerate(np.atleast_2d(modules)):
            p = 0 if profile_of_job is None else int(profile_of_job[j])
            union = self.dense[:, m[m >= 0]].any(axis=1)
            z, pv, s2, n1 = oracle.ranksum(self.w[p], np.arange(self.S), bitpack.pack_mask(np.nonzero(union)[0], self.S))
            for k_, v in zip(("z", "p", "twice_rank_sum", "n1"), (z, pv, s2, n1)):
                out[k_].append(v)
        return {"z": np.array(out["z"]), "p": np.array(out["p"]), "twice_rank_sum": np.array(out["twice_rank_sum"], dtype=np.int64),
                "n1": np.array(out["n1"], dtype=np.int32)}
