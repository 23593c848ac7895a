"""oracle — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the SuperDendrix significance-and-search path, used only as
the checker by tests/, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.  Nothing under ``superdendrix_b200/``
imports this package.

Two parts:

* ``sdx_oracle.c`` (built to ``oracle/_build/libsdx_oracle.so`` by
  :func:`build`) — the canonical, Philox-driven restatement that the CUDA path
  must match bit for bit (integers) / within 1e-9 (fp64 weights).
* ``refport.py`` — a pure-Python port that follows the reference's own data
  structures (lists, sets, CPython ``random``); it is what the CPU baseline
  times, because that is what the reference runs.

Pinning: ``tests/test_oracle_golden.py`` checks both against fixtures produced
by importing the reference itself (``tests/golden/make_golden.py``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsdx_oracle.so")
_SRC = os.path.join(_HERE, "sdx_oracle.c")
_lib = None


def build(force: bool = False) -> str:
    """Compile sdx_oracle.c with gcc (no GPU, no reference sources involved)."""
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(
            ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fvisibility=hidden",
             "-ffp-contract=off", "-o", _SO, _SRC, "-lm"])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.sdxo_curveball.restype = C.c_int64
        _lib.sdxo_curveball_replay.restype = C.c_int64
        _lib.sdxo_scan.restype = C.c_int64
        _lib.sdxo_below_from_draws.restype = C.c_uint32
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a


# --------------------------------------------------------------------------
def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().sdxo_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def stream_draws(seed, kind, sid, t, n):
    out = np.zeros(n, dtype=np.uint32)
    lib().sdxo_stream_draws(C.c_uint64(seed), C.c_int(kind), C.c_uint64(sid), C.c_uint32(t), C.c_int(n),
                            _p(out, C.c_uint32))
    return out


def below_from_draws(draws, m):
    d = _u32(draws)
    used = C.c_int(0)
    v = lib().sdxo_below_from_draws(_p(d, C.c_uint32), C.c_int(len(d)), C.c_uint32(m), C.byref(used))
    return int(v), used.value


def deal_stats(reset=False):
    """(times the Lemire rejection zone was entered, draws rejected) by the canonical trade since the last reset."""
    z, r = C.c_uint64(0), C.c_uint64(0)
    lib().sdxo_deal_stats(C.byref(z), C.byref(r), C.c_int(1 if reset else 0))
    return z.value, r.value


def trade_pair(seed, null_id, t, R):
    a, b = C.c_uint32(0), C.c_uint32(0)
    lib().sdxo_trade_pair(C.c_uint64(seed), C.c_uint64(null_id), C.c_uint32(t), C.c_uint32(R),
                          C.byref(a), C.byref(b))
    return a.value, b.value


def curveball_nulls(observed, seed, null0, n, n_trades, chain_len=1, want_matrices=True, burn=0, pool_size=1024, pair_group=1, L=None):
    """observed: (R, wpr) uint32 packed rows over the longer dimension.
    burn > 0: chains start from the burn-in pool (see sdx_oracle.c); L = bits per row (the burn-in may run on the transpose,
    so the number of columns must be known exactly).
    pair_group > 1: aligned blocks of that many chains draw their row pairs from the stream of the block's first chain.
    Returns (nulls (n, R, wpr) or None, applied (n,))."""
    obs = _u32(observed)
    R, wpr = obs.shape
    if burn > 0 and L is None:
        raise ValueError("burn > 0 needs L (bits per row)")
    out = np.zeros((n, R, wpr), dtype=np.uint32) if want_matrices else None
    applied = np.zeros(n, dtype=np.int64)
    lib().sdxo_curveball_nulls_group(_p(obs, C.c_uint32), C.c_int(R), C.c_int(wpr), C.c_int(wpr * 32 if L is None else L), C.c_uint64(seed),
                                     C.c_uint64(null0), C.c_int64(n), C.c_int64(n_trades), C.c_int64(chain_len),
                                     C.c_int64(burn), C.c_int(pool_size), C.c_int(pair_group), _p(out, C.c_uint32),
                                     _p(applied, C.c_int64))
    return out, applied


def pool_state(observed, L, seed, burn, j):
    """start state j of the burn-in pool: the observed matrix (R rows of L bits) after `burn` burn-in rounds on reserved streams"""
    obs = _u32(observed)
    R, wpr = obs.shape
    out = np.zeros_like(obs)
    lib().sdxo_pool_state(_p(obs, C.c_uint32), C.c_int(R), C.c_int(wpr), C.c_int(L), C.c_uint64(seed),
                          C.c_int64(burn), C.c_int(j), _p(out, C.c_uint32))
    return out


def burn_plan(observed, L):
    """what the burn-in of this matrix looks like: dict(alt, rows, trades_per_round, dense_rows)"""
    obs = _u32(observed)
    R, wpr = obs.shape
    info = np.zeros(4, dtype=np.int64)
    lib().sdxo_burn_plan(_p(obs, C.c_uint32), C.c_int(R), C.c_int(wpr), C.c_int(L), _p(info, C.c_int64))
    return {"alt": bool(info[0]), "rows": int(info[1]), "trades_per_round": int(info[2]), "dense_rows": int(info[3])}


def curveball_replay(rows, a, b, to_a):
    rows = _u32(rows).copy()
    R, wpr = rows.shape
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    m = _u32(to_a)
    assert m.shape == (len(a), wpr)
    ap = lib().sdxo_curveball_replay(_p(rows, C.c_uint32), C.c_int(R), C.c_int(wpr), _p(a, C.c_int32),
                                     _p(b, C.c_int32), _p(m, C.c_uint32), C.c_int64(len(a)))
    return rows, int(ap)


def score_set(rows, S, feats, w, mask=None):
    rows = _u32(rows)
    f = np.ascontiguousarray(feats, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    mk = _u32(mask) if mask is not None else None
    W = C.c_double(0)
    cov, ov, ac = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    lib().sdxo_score_set(_p(rows, C.c_uint32), C.c_int(rows.shape[1]), C.c_int(S), _p(f, C.c_int32),
                         C.c_int(len(f)), _p(w, C.c_double), _p(mk, C.c_uint32),
                         C.byref(W), C.byref(cov), C.byref(ov), C.byref(ac))
    return W.value, cov.value, ov.value, ac.value


def feature_scores(rows, S, w, mask=None):
    rows = _u32(rows)
    F, wpr = rows.shape
    w = np.ascontiguousarray(w, dtype=np.float64)
    mk = _u32(mask) if mask is not None else None
    sw = np.zeros(F); W1 = np.zeros(F); cnt = np.zeros(F, dtype=np.int32)
    best = C.c_int32(-1)
    lib().sdxo_feature_scores(_p(rows, C.c_uint32), C.c_int(F), C.c_int(wpr), C.c_int(S), _p(w, C.c_double),
                              _p(mk, C.c_uint32), _p(sw, C.c_double), _p(W1, C.c_double), _p(cnt, C.c_int32),
                              C.byref(best))
    return sw, W1, cnt, best.value


def scan(rows, S, w, k, topn=8, mask=None):
    rows = _u32(rows)
    F, wpr = rows.shape
    w = np.ascontiguousarray(w, dtype=np.float64)
    mk = _u32(mask) if mask is not None else None
    oW = np.zeros(topn); oF = np.zeros((topn, 3), dtype=np.int32); oC = np.zeros(topn, dtype=np.int32)
    n = lib().sdxo_scan(_p(rows, C.c_uint32), C.c_int(F), C.c_int(wpr), C.c_int(S), _p(w, C.c_double),
                        _p(mk, C.c_uint32), C.c_int(k), C.c_int(topn), _p(oW, C.c_double), _p(oF, C.c_int32),
                        _p(oC, C.c_int32))
    return oW, oF, oC, int(n)


def scan_range(rows, S, w, k, topn, a_lo, a_hi, mask=None):
    """The scan restricted to subsets whose smallest feature index is in [a_lo, a_hi)."""
    rows = _u32(rows)
    F, wpr = rows.shape
    w = np.ascontiguousarray(w, dtype=np.float64)
    mk = _u32(mask) if mask is not None else None
    oW = np.zeros(topn); oF = np.zeros((topn, 3), dtype=np.int32); oC = np.zeros(topn, dtype=np.int32)
    lib().sdxo_scan_range.restype = C.c_int64
    n = lib().sdxo_scan_range(_p(rows, C.c_uint32), C.c_int(F), C.c_int(wpr), C.c_int(S), _p(w, C.c_double),
                              _p(mk, C.c_uint32), C.c_int(k), C.c_int(topn), C.c_int(a_lo), C.c_int(a_hi), _p(oW, C.c_double),
                              _p(oF, C.c_int32), _p(oC, C.c_int32))
    return oW, oF, oC, int(n)


def condperm(rows, S, feats, w, sample_idx, seed, perm_id0, P):
    rows = _u32(rows)
    f = np.ascontiguousarray(feats, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    si = np.ascontiguousarray(sample_idx, dtype=np.int32)
    counts = np.zeros(len(f), dtype=np.int64)
    realW = np.zeros(len(f))
    lib().sdxo_condperm(_p(rows, C.c_uint32), C.c_int(rows.shape[1]), C.c_int(S), _p(f, C.c_int32), C.c_int(len(f)),
                        _p(w, C.c_double), _p(si, C.c_int32), C.c_int(len(si)), C.c_uint64(seed),
                        C.c_uint64(perm_id0), C.c_int64(P), _p(counts, C.c_int64), _p(realW, C.c_double))
    return counts, realW


def ranksum(profile, sample_idx, union_bits):
    pr = np.ascontiguousarray(profile, dtype=np.float64)
    si = np.ascontiguousarray(sample_idx, dtype=np.int32)
    ub = _u32(union_bits)
    z, p = C.c_double(0), C.c_double(0)
    s2, n1 = C.c_int64(0), C.c_int32(0)
    lib().sdxo_ranksum(_p(pr, C.c_double), _p(si, C.c_int32), C.c_int(len(si)), _p(ub, C.c_uint32),
                       C.byref(z), C.byref(p), C.byref(s2), C.byref(n1))
    return z.value, p.value, s2.value, n1.value
