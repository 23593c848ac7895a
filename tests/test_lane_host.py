"""The one-thread-per-null trade code (csrc/sdx_lane_trade.cuh) on the CPU against the oracle — no GPU needed.

tests/tools/lane_host.cu compiles the very functions k_trade_lanes calls as host code and replaces the TMA staging by
memcpy; here a group of 32 chains is run and compared bit for bit with oracle.curveball_nulls(pair_group=32), which
restates src/curveball.py:17-40 with the canonical streams."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from superdendrix_b200 import bitpack, synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "tools", "lane_host.cu")
SO = os.path.join(HERE, "tools", "_build", "liblane_host.so")
HDRS = [os.path.join(HERE, "..", "superdendrix_b200", "csrc", f) for f in ("sdx_lane_trade.cuh", "sdx_common.cuh")]


def _lib():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    newest = max(os.path.getmtime(p) for p in [SRC] + HDRS)
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call([nvcc, "-x", "cu", "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-cudart", "static",
                               "-gencode", "arch=compute_100a,code=sm_100a", "-o", SO, SRC])
    lib = C.CDLL(SO)
    return lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def run_group(lib, start, ts, seed, chain0, chain_len, round0, rounds, T, score=None):
    start = np.ascontiguousarray(start, dtype=np.uint32)
    if start.ndim == 2:
        start = start[None]
    n_start, R, nw = start.shape
    out = np.zeros((rounds, 32, R, nw), dtype=np.uint32)
    ap = np.zeros((rounds, 32), dtype=np.int64)
    stats = np.zeros(3, dtype=np.int64)
    W = None
    mod = w = mask = None
    k = S = 0
    if score is not None:
        mod, w, mask, S = score
        mod = np.ascontiguousarray(mod, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint32)
        k = len(mod)
        W = np.zeros((rounds, 32))
    rc = lib.lane_host_run(_p(start, C.c_uint32), n_start, R, nw, ts, C.c_uint64(seed), C.c_uint64(chain0), C.c_int64(chain_len),
                           C.c_int64(round0), rounds, C.c_int64(T), _p(out, C.c_uint32), _p(ap, C.c_int64), _p(mod, C.c_int32), k,
                           _p(w, C.c_double), _p(mask, C.c_uint32), S, _p(W, C.c_double), _p(stats, C.c_int64))
    assert rc == 0
    return out, ap, W, stats


def oracle_group(obs, seed, chain0, chain_len, rounds, T):
    """nulls of chains chain0 .. chain0 + 31, rounds 0 .. rounds - 1, as (rounds, 32, R, nw)"""
    n = 32 * chain_len
    nulls, ap = oracle.curveball_nulls(obs, seed, chain0 * chain_len, n, T, chain_len, pair_group=32)
    nulls = nulls.reshape(32, chain_len, *obs.shape)[:, :rounds].transpose(1, 0, 2, 3)
    ap = ap.reshape(32, chain_len)[:, :rounds].T
    return nulls, ap


CASES = {
    # name: (S, F, density or None for the DepMap-like synth, ts)
    "c0_like_lists_and_bits": (300, 120, None, 0),
    "all_lists": (200, 60, 0.03, 0),
    "all_bits": (70, 40, 0.5, 4),
    "mixed_dense": (130, 50, 0.25, 0),
    "wide_masks": (400, 90, 0.08, 40),        # lists up to 40 entries: chosen ranks in memory (MREG = false)
    "tiny": (33, 5, 0.4, 0),
    "word_boundary": (64, 30, 0.2, 0),
    "rows_are_samples": (40, 90, 0.15, 0),
}


def _matrix(name):
    S, F, dens, ts = CASES[name]
    if dens is None:
        dense, _, _ = synth.feature_matrix(S, F)
    else:
        rng = np.random.default_rng(abs(hash(name)) % 1000 + 7)
        dense = (rng.random((S, F)) < dens).astype(np.uint8)
        dense[:, 0] = 1                       # a full row / column
        dense[:, 1] = 0                       # an empty one
    S, F = dense.shape
    obs = bitpack.pack_rows(dense if F >= S else dense.T)
    return dense, obs, ts


@pytest.mark.parametrize("name", list(CASES))
def test_lane_code_matches_oracle_bit_for_bit(name):
    lib = _lib()
    dense, obs, ts = _matrix(name)
    R = obs.shape[0]
    T = 5 * R
    chain_len, rounds = 3, 3
    for chain0 in (0, 64):
        got, got_ap, _, stats = run_group(lib, obs, ts, 1764, chain0, chain_len, 0, rounds, T)
        want, want_ap = oracle_group(obs, 1764, chain0, chain_len, rounds, T)
        assert np.array_equal(got, want), name
        assert np.array_equal(got_ap, want_ap)
        # margins (G1): row and column sums of every null equal the observed ones
        L = max(dense.shape)
        for nul in got[-1][:4]:
            bits = bitpack.unpack_rows(nul, L)
            ob = bitpack.unpack_rows(obs, L)
            assert np.array_equal(bits.sum(0), ob.sum(0)) and np.array_equal(bits.sum(1), ob.sum(1))
    assert stats[2] == T * rounds


def test_lane_code_from_pool_states():
    """chains start from the burn-in pool: chain c from state c % pool_size"""
    lib = _lib()
    dense, obs, ts = _matrix("mixed_dense")
    R = obs.shape[0]
    T = 5 * R
    burn, pool_size, chain_len, seed = 2, 7, 2, 99
    pool = np.stack([oracle.pool_state(obs, max(dense.shape), seed, burn, j) for j in range(pool_size)])
    got, got_ap, _, _ = run_group(lib, pool, ts, seed, 32, chain_len, 0, chain_len, T)
    want, want_ap = oracle.curveball_nulls(obs, seed, 32 * chain_len, 32 * chain_len, T, chain_len, burn=burn, pool_size=pool_size,
                                           pair_group=32, L=max(dense.shape))
    want = want.reshape(32, chain_len, *obs.shape).transpose(1, 0, 2, 3)
    assert np.array_equal(got, want)
    assert np.array_equal(got_ap, want_ap.reshape(32, chain_len).T)


def test_pair_group_changes_only_the_pair_stream():
    """pair_group = 32: lane 0 of a block is the pair_group = 1 null; the other lanes differ from their pair_group = 1 selves"""
    dense, obs, _ = _matrix("all_lists")
    T = 5 * obs.shape[0]
    a, _ = oracle.curveball_nulls(obs, 11, 0, 64, T, 1, pair_group=1)
    b, _ = oracle.curveball_nulls(obs, 11, 0, 64, T, 1, pair_group=32)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[32], b[32])
    assert not np.array_equal(a[1], b[1])
    L = max(dense.shape)
    ob = bitpack.unpack_rows(obs, L)
    for nul in b[:8]:
        bits = bitpack.unpack_rows(nul, L)
        assert np.array_equal(bits.sum(0), ob.sum(0)) and np.array_equal(bits.sum(1), ob.sum(1))


def test_lane_scorer_is_bit_identical_to_a_butterfly_sum():
    """ln_score forms the fp64 sums exactly as score_module (sdx_trade.cuh): 32 partials + butterfly"""
    lib = _lib()
    dense, obs, ts = _matrix("c0_like_lists_and_bits")
    S, F = dense.shape
    assert F < S                                  # rows are features
    w = synth.weights(S, 1)[0]
    mask = np.full(bitpack.words_for(S), 0xffffffff, dtype=np.uint32)
    if S % 32:
        mask[-1] = (1 << (S % 32)) - 1
    mod = np.argsort(-dense.sum(0), kind="stable")[1:4].astype(np.int32)      # dense rows (bit-packed) ...
    mod2 = np.array([int(np.argsort(dense.sum(0), kind="stable")[5]), int(mod[0]), -1], dtype=np.int32)   # ... and a list row, an empty slot
    R = obs.shape[0]
    for m in (mod, mod2):
        nulls, _, W, _ = run_group(lib, obs, ts, 7, 0, 1, 0, 1, 5 * R, score=(m, w, mask, S))
        for lane in range(32):
            rows = nulls[0, lane]
            # reference: partial p = words p, p + 32, ..; ascending samples; then the xor-butterfly 16, 8, 4, 2, 1
            pos = np.zeros(32); neg = np.zeros(32)
            for wi in range(rows.shape[1]):
                for b in range(32):
                    s = wi * 32 + b
                    if s >= S or not (mask[wi] >> b) & 1:
                        continue
                    mult = sum(int((rows[g, wi] >> b) & 1) for g in m if g >= 0)
                    if mult:
                        if w[s] > 0:
                            pos[wi & 31] += w[s]
                        neg[wi & 31] += mult * abs(w[s])
            for o in (16, 8, 4, 2, 1):
                for i in range(o):
                    pos[i] += pos[i + o]; neg[i] += neg[i + o]
            assert W[0, lane] == 2.0 * pos[0] - neg[0]
            want, *_ = oracle.score_set(rows, S, m[m >= 0], w)
            assert abs(W[0, lane] - want) <= 1e-9 * np.abs(w).sum()


def test_schedule_flags_count_shared_rows():
    lib = _lib()
    dense, obs, ts = _matrix("all_lists")
    R = obs.shape[0]
    T = 5 * R
    _, _, _, stats = run_group(lib, obs, ts, 3, 0, 1, 0, 1, T)
    # a trade shares a row with the one before it with probability ~ 4 / R
    assert 0.3 * 4 / R < stats[0] / stats[2] < 3 * 4 / R
