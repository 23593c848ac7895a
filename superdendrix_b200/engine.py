"""Engine — a thin, typed Python face of one libsdx context (one process, one GPU).

Every method maps 1:1 onto an entry point of include/sdx.h; numpy arrays in,
numpy arrays out.  The reference-named shims (superdendrix_b200/curveball.py,
superdendrix_b200/superdendrix.py) and the CLIs are written on top of this.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi, bitpack
from ._capi import SDX_STAT_FIXED, SDX_STAT_MAXK, SdxError  # noqa: F401  (re-exported)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Engine:
    def __init__(self, device: int = 0):
        self._lib = _capi.load()
        self._ctx = C.c_void_p()
        rc = self._lib.sdx_create(int(device), C.byref(self._ctx))
        if rc != 0:
            msg = self._lib.sdx_last_error(None)
            raise SdxError(rc, msg.decode() if msg else "sdx_create failed")
        self.device = int(device)
        self.F = self.S = self.P = 0

    # -- plumbing ------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._lib.sdx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            msg = self._lib.sdx_last_error(self._ctx)
            raise SdxError(rc, msg.decode() if msg else "")

    def set_stream(self, cuda_stream: int | None):
        """Run on a caller-owned CUDA stream (e.g. ``torch.cuda.current_stream().cuda_stream``)."""
        self._check(self._lib.sdx_set_stream(self._ctx, C.c_void_p(cuda_stream or 0)))

    def set_burn_in(self, burn_rounds: int, pool_size: int = 1472):
        """Chains start from one of ``pool_size`` states that have already seen ``burn_rounds`` rounds of trades
        (0 = start from the observed matrix, as a single reference chain does at null 0)."""
        self._check(self._lib.sdx_set_burn_in(self._ctx, int(burn_rounds), int(pool_size)))

    def set_pair_group(self, pair_group: int):
        """1: every null draws its own row pairs.  32: aligned blocks of 32 chains share the schedule of row pairs
        (every null keeps its own deal stream) — the layout the one-thread-per-null kernel needs (include/sdx.h)."""
        self._check(self._lib.sdx_set_pair_group(self._ctx, int(pair_group)))

    def pool_begin(self, seed: int, n_trades: int = -1):
        """Start building the burn-in pool in parts (multi-GPU): returns (device address, bytes per state)."""
        ptr, nb = C.c_void_p(), C.c_size_t()
        self._check(self._lib.sdx_pool_begin(self._ctx, C.c_uint64(seed), C.c_int64(n_trades), C.byref(ptr), C.byref(nb)))
        return ptr.value, nb.value

    def pool_build(self, first_state: int, n_states: int):
        self._check(self._lib.sdx_pool_build(self._ctx, int(first_state), int(n_states)))

    def pool_valid(self, seed: int, n_trades: int = -1) -> bool:
        return bool(self._lib.sdx_pool_valid(self._ctx, C.c_uint64(seed), C.c_int64(n_trades)))

    def pool_commit(self):
        self._check(self._lib.sdx_pool_commit(self._ctx))

    def timing(self) -> dict:
        t = _capi.Timing()
        self._check(self._lib.sdx_get_timing(self._ctx, C.byref(t)))
        return {"plan_ms": t.plan_ms, "trade_ms": t.trade_ms, "score_ms": t.score_ms, "total_ms": t.total_ms,
                "launches": t.launches, "trade_launches": t.trade_launches}

    def total_launches(self) -> int:
        n = C.c_int64(0)
        self._check(self._lib.sdx_total_launches(self._ctx, C.byref(n)))
        return n.value

    def measure_int_peaks(self) -> dict:
        """Sustained 32-bit POPC and LOP3 rates of this GPU (ops/s), measured by two register-only micro-kernels."""
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._lib.sdx_measure_int_peaks(self._ctx, C.byref(a), C.byref(b)))
        return {"popc_per_s": a.value, "lop3_per_s": b.value}

    # -- uploads -------------------------------------------------------------
    def upload_matrix(self, feature_rows: np.ndarray, n_samples: int):
        """feature_rows: (F, ceil(S/32)) uint32 packed rows over samples."""
        rows = np.ascontiguousarray(feature_rows, dtype=np.uint32)
        if rows.ndim != 2:
            raise ValueError("feature_rows must be 2-D")
        F, wpr = rows.shape
        self._check(self._lib.sdx_upload_matrix(self._ctx, _ptr(rows, C.c_uint32), F, int(n_samples), wpr))
        self.F, self.S = F, int(n_samples)
        self.P = 0

    def upload_dense(self, dense: np.ndarray):
        """dense: (S, F) 0/1 sample x feature matrix (the layout of src/generate_null_matrices.py:62)."""
        d = np.asarray(dense)
        self.upload_matrix(bitpack.pack_rows(d.T), d.shape[0])

    def upload_weights(self, w: np.ndarray, sample_mask: np.ndarray | None = None):
        """w: (P, S) or (S,) fp64; sample_mask: optional (P, S) bool / (P, wpr) packed uint32."""
        w = np.ascontiguousarray(np.atleast_2d(np.asarray(w, dtype=np.float64)))
        P, S = w.shape
        mk = None
        if sample_mask is not None:
            m = np.asarray(sample_mask)
            if m.dtype != np.uint32:
                m = bitpack.pack_rows(np.atleast_2d(m))
            mk = np.ascontiguousarray(np.atleast_2d(m), dtype=np.uint32)
            if mk.shape != (P, bitpack.words_for(S)):
                raise ValueError("sample_mask shape mismatch")
        self._check(self._lib.sdx_upload_weights(self._ctx, _ptr(w, C.c_double), _ptr(mk, C.c_uint32), P, S))
        self.P = P

    def trade_shape(self):
        R, L, wpr, ras = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.sdx_trade_shape(self._ctx, C.byref(R), C.byref(L), C.byref(wpr), C.byref(ras)))
        return R.value, L.value, wpr.value, bool(ras.value)

    # -- scoring -------------------------------------------------------------
    def score_sets(self, sets, profile_of_set=None):
        """sets: (n, k) int32 feature indices (-1 padded).  Returns dict of W, coverage, overlap, act_cov."""
        s = np.ascontiguousarray(np.atleast_2d(np.asarray(sets, dtype=np.int32)))
        n, k = s.shape
        pr = None if profile_of_set is None else np.ascontiguousarray(profile_of_set, dtype=np.int32)
        W = np.zeros(n)
        cov = np.zeros(n, dtype=np.int32)
        ovl = np.zeros(n, dtype=np.int32)
        act = np.zeros(n, dtype=np.int32)
        self._check(self._lib.sdx_score_sets(self._ctx, _ptr(s, C.c_int32), _ptr(pr, C.c_int32), n, k,
                                             _ptr(W, C.c_double), _ptr(cov, C.c_int32), _ptr(ovl, C.c_int32),
                                             _ptr(act, C.c_int32)))
        return {"W": W, "coverage": cov, "overlap": ovl, "act_cov": act}

    def union_sums(self, sets, profile_of_set=None):
        """Sum of w over the union of each set (n, k) -> (n,) fp64."""
        s = np.ascontiguousarray(np.atleast_2d(np.asarray(sets, dtype=np.int32)))
        n, k = s.shape
        pr = None if profile_of_set is None else np.ascontiguousarray(profile_of_set, dtype=np.int32)
        out = np.zeros(n)
        self._check(self._lib.sdx_union_sums(self._ctx, _ptr(s, C.c_int32), _ptr(pr, C.c_int32), n, k, _ptr(out, C.c_double)))
        return out

    def feature_scores(self, profile: int = 0):
        F = self.F
        sw, W1 = np.zeros(F), np.zeros(F)
        cnt = np.zeros(F, dtype=np.int32)
        best = C.c_int32(-1)
        self._check(self._lib.sdx_feature_scores(self._ctx, int(profile), _ptr(sw, C.c_double), _ptr(W1, C.c_double),
                                                 _ptr(cnt, C.c_int32), C.byref(best)))
        return {"sum_w": sw, "W1": W1, "count": cnt, "best_single": best.value}

    # -- Curveball -----------------------------------------------------------
    def curveball(self, seed: int, null0: int, n_nulls: int, n_trades: int = -1, chain_len: int = 1,
                  want_rows: bool = True, want_applied: bool = True):
        """Returns (rows (n, R, wpr) uint32 in the trade orientation or None, applied (n,) int64 or None)."""
        R, _, wpr, _ = self.trade_shape()
        out = np.zeros((n_nulls, R, wpr), dtype=np.uint32) if want_rows else None
        ap = np.zeros(n_nulls, dtype=np.int64) if want_applied else None
        self._check(self._lib.sdx_curveball(self._ctx, C.c_uint64(seed), C.c_uint64(null0), n_nulls, n_trades,
                                            chain_len, _ptr(out, C.c_uint32), _ptr(ap, C.c_int64)))
        return out, ap

    def curveball_features(self, seed: int, null0: int, n_nulls: int, n_trades: int = -1, chain_len: int = 1):
        """The nulls of ``curveball`` as packed FEATURE rows over samples, (n, F, ceil(S/32)) uint32, whatever the trade
        orientation (transposed on the device when the trade rows are samples)."""
        out = np.zeros((n_nulls, self.F, bitpack.words_for(self.S)), dtype=np.uint32)
        self._check(self._lib.sdx_curveball_features(self._ctx, C.c_uint64(seed), C.c_uint64(null0), n_nulls, n_trades, chain_len,
                                                     _ptr(out, C.c_uint32)))
        return out

    def curveball_replay(self, a, b, to_a, start_rows=None):
        R, _, wpr, _ = self.trade_shape()
        a = np.ascontiguousarray(a, dtype=np.int32)
        b = np.ascontiguousarray(b, dtype=np.int32)
        m = np.ascontiguousarray(to_a, dtype=np.uint32)
        if m.shape != (len(a), wpr):
            raise ValueError("to_a must have shape (n_trades, words_per_row)")
        st = None if start_rows is None else np.ascontiguousarray(start_rows, dtype=np.uint32)
        if st is not None and st.shape != (R, wpr):
            raise ValueError("start_rows must have shape (R, words_per_row)")
        out = np.zeros((R, wpr), dtype=np.uint32)
        ap = C.c_int64(0)
        self._check(self._lib.sdx_curveball_replay(self._ctx, _ptr(st, C.c_uint32), _ptr(a, C.c_int32),
                                                   _ptr(b, C.c_int32), _ptr(m, C.c_uint32), len(a),
                                                   _ptr(out, C.c_uint32), C.byref(ap)))
        return out, ap.value

    def null_pvalue(self, seed: int, null0: int, n_nulls: int, modules, n_trades: int = -1, chain_len: int = 1,
                    stat: int = SDX_STAT_FIXED, z_obs=None, want_scores: bool = False):
        """Returns dict(n_ge (P,), z_obs (P,), null_scores (P, n) or None)."""
        mods = np.ascontiguousarray(np.atleast_2d(np.asarray(modules, dtype=np.int32)))
        P, k = mods.shape
        if P != self.P:
            raise ValueError("modules must have one row per uploaded profile (%d), got %d" % (self.P, P))
        zin = None if z_obs is None else np.ascontiguousarray(np.atleast_1d(z_obs), dtype=np.float64)
        n_ge = np.zeros(P, dtype=np.int64)
        scores = np.zeros((P, n_nulls)) if want_scores else None
        zout = np.zeros(P)
        self._check(self._lib.sdx_null_pvalue(self._ctx, C.c_uint64(seed), C.c_uint64(null0), n_nulls, n_trades,
                                              chain_len, int(stat), _ptr(mods, C.c_int32), k, _ptr(zin, C.c_double),
                                              _ptr(n_ge, C.c_int64), _ptr(scores, C.c_double), _ptr(zout, C.c_double)))
        return {"n_ge": n_ge, "z_obs": zout, "null_scores": scores}

    # -- scan / model selection / rank-sum -------------------------------------
    def null_pvalue_rows(self, feature_rows, modules, stat: int = SDX_STAT_FIXED, z_obs=None, want_scores: bool = False):
        """The null loop on caller-supplied nulls: feature_rows (n, F, ceil(S/32)) uint32 packed feature rows, every null
        scored for every uploaded profile.  Returns dict(n_ge (P,), z_obs (P,), null_scores (P, n) or None)."""
        rows = np.ascontiguousarray(feature_rows, dtype=np.uint32)
        if rows.ndim != 3 or rows.shape[1] != self.F or rows.shape[2] != bitpack.words_for(self.S):
            raise ValueError("feature_rows must be (n, %d, %d)" % (self.F, bitpack.words_for(self.S)))
        mods = np.ascontiguousarray(np.atleast_2d(np.asarray(modules, dtype=np.int32)))
        P, k = mods.shape
        if P != self.P:
            raise ValueError("modules must have one row per uploaded profile (%d), got %d" % (self.P, P))
        n = rows.shape[0]
        zin = None if z_obs is None else np.ascontiguousarray(np.atleast_1d(z_obs), dtype=np.float64)
        n_ge = np.zeros(P, dtype=np.int64)
        scores = np.zeros((P, n)) if want_scores else None
        zout = np.zeros(P)
        self._check(self._lib.sdx_null_pvalue_rows(self._ctx, _ptr(rows, C.c_uint32), n, int(stat), _ptr(mods, C.c_int32), k,
                                                   _ptr(zin, C.c_double), _ptr(n_ge, C.c_int64), _ptr(scores, C.c_double), _ptr(zout, C.c_double)))
        return {"n_ge": n_ge, "z_obs": zout, "null_scores": scores}

    def scan(self, profile: int = 0, k: int = 3, topn: int = 8, g_lo: int = 0, g_hi: int | None = None):
        """Exhaustive scan over subsets of size <= k whose smallest feature index is in [g_lo, g_hi)."""
        hits = (_capi.Hit * topn)()
        n = C.c_int64(0)
        self._check(self._lib.sdx_scan_range(self._ctx, int(profile), int(k), int(topn), int(g_lo),
                                             int(self.F if g_hi is None else g_hi), hits, C.byref(n)))
        W = np.array([h.W for h in hits])
        feats = np.array([[h.features[0], h.features[1], h.features[2]] for h in hits], dtype=np.int32)
        cov = np.array([h.coverage for h in hits], dtype=np.int32)
        return {"W": W, "features": feats, "coverage": cov, "n_scored": n.value}

    def condperm(self, module, n_perm: int, seed: int, perm_id0: int = 0, profile: int = 0):
        """Returns (counts (k,) int64, real_W (k,)) for one module of one profile."""
        m = np.ascontiguousarray(module, dtype=np.int32)
        counts = np.zeros(len(m), dtype=np.int64)
        real = np.zeros(len(m))
        self._check(self._lib.sdx_condperm(self._ctx, int(profile), _ptr(m, C.c_int32), len(m), C.c_uint64(seed),
                                           C.c_uint64(perm_id0), int(n_perm), _ptr(counts, C.c_int64),
                                           _ptr(real, C.c_double)))
        return counts, real

    def condperm_batch(self, modules, n_perm: int, seed: int, perm_id0: int = 0, profile_of_job=None):
        """modules: (n_jobs, k).  Returns (counts (n_jobs, k), real_W (n_jobs, k))."""
        mods = np.ascontiguousarray(np.atleast_2d(np.asarray(modules, dtype=np.int32)))
        n, k = mods.shape
        pj = np.zeros(n, dtype=np.int32) if profile_of_job is None else np.ascontiguousarray(profile_of_job, dtype=np.int32)
        counts = np.zeros((n, k), dtype=np.int64)
        real = np.zeros((n, k))
        self._check(self._lib.sdx_condperm_batch(self._ctx, _ptr(pj, C.c_int32), _ptr(mods, C.c_int32), n, k,
                                                 C.c_uint64(seed), C.c_uint64(perm_id0), int(n_perm),
                                                 _ptr(counts, C.c_int64), _ptr(real, C.c_double)))
        return counts, real

    def ranksum(self, modules, profile_of_job=None):
        mods = np.ascontiguousarray(np.atleast_2d(np.asarray(modules, dtype=np.int32)))
        n, k = mods.shape
        pj = np.zeros(n, dtype=np.int32) if profile_of_job is None else np.ascontiguousarray(profile_of_job, dtype=np.int32)
        z, p = np.zeros(n), np.zeros(n)
        s2 = np.zeros(n, dtype=np.int64)
        n1 = np.zeros(n, dtype=np.int32)
        self._check(self._lib.sdx_ranksum(self._ctx, _ptr(pj, C.c_int32), _ptr(mods, C.c_int32), n, k,
                                          _ptr(z, C.c_double), _ptr(p, C.c_double), _ptr(s2, C.c_int64),
                                          _ptr(n1, C.c_int32)))
        return {"z": z, "p": p, "twice_rank_sum": s2, "n1": n1}

    # -- self-test hooks -------------------------------------------------------
    def test_philox(self, ctr, key):
        c = np.ascontiguousarray(ctr, dtype=np.uint32).reshape(-1, 4)
        k = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros_like(c)
        self._check(self._lib.sdx_test_philox(self._ctx, _ptr(c, C.c_uint32), _ptr(k, C.c_uint32), len(c),
                                              _ptr(out, C.c_uint32)))
        return out

    def test_trade_pairs(self, seed, null_id, n_trades, n_rows):
        a = np.zeros(n_trades, dtype=np.int32)
        b = np.zeros(n_trades, dtype=np.int32)
        self._check(self._lib.sdx_test_trade_pairs(self._ctx, C.c_uint64(seed), C.c_uint64(null_id), n_trades, n_rows,
                                                   _ptr(a, C.c_int32), _ptr(b, C.c_int32)))
        return a, b
