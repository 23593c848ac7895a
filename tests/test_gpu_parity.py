"""Parity of the CUDA path (libsdx.so through the C ABI) against the oracle and
the golden fixtures made from the reference.  Needs a B200: run with -m gpu."""
import json
import os

import numpy as np
import pytest

import oracle
from superdendrix_b200 import bitpack, defaults, synth
from superdendrix_b200.engine import SDX_STAT_FIXED, SDX_STAT_MAXK, Engine, SdxError

from conftest import GOLDEN, burn_cases, curveball_cases, load_npz, rows_of

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def upload_dense(eng, dense):
    S, F = dense.shape
    eng.upload_matrix(bitpack.pack_rows(dense.T), S)


# ---------------------------------------------------------------- random streams
def test_philox_known_answers_on_device(eng):
    ctr = [(0, 0, 0, 0), (0xffffffff,) * 4, (0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344)]
    key = [(0, 0), (0xffffffff,) * 2, (0xa4093822, 0x299f31d0)]
    want = [(0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)]
    out = eng.test_philox(ctr, key)
    assert [tuple(int(x) for x in r) for r in out] == want


@pytest.mark.parametrize("R", [2, 3, 400, 770, 65535])
def test_trade_pairs_match_oracle(eng, R):
    a, b = eng.test_trade_pairs(1764, 12345678901, 300, R)
    for t in range(300):
        assert (int(a[t]), int(b[t])) == oracle.trade_pair(1764, 12345678901, t, R)


# ---------------------------------------------------------------- Curveball: replay of the reference's own schedules
@pytest.mark.parametrize("name", ["tall_small", "wide_small", "square", "short_run", "c0_shape"])
def test_replay_reproduces_reference_bit_for_bit(eng, name):
    cases, _ = curveball_cases()
    c = cases[name]
    upload_dense(eng, c["dense"])
    rows = None
    for i in range(c["a"].shape[0]):
        rows, applied = eng.curveball_replay(c["a"][i], c["b"][i], c["to_a"][i], start_rows=rows)
        assert np.array_equal(rows, c["out"][i]), "null %d differs from the reference output" % i
        assert applied == int(c["to_a"][i].any(axis=1).sum())


# ---------------------------------------------------------------- Curveball: canonical streams vs the oracle
def _dense_cases():
    rng = np.random.default_rng(11)
    cases, _ = curveball_cases()
    out = {
        "c0_shape": cases["c0_shape"]["dense"],                        # 770 x ~384: rows = features, G=8
        "wide": (rng.random((60, 300)) < 0.1).astype(np.uint8),         # rows = samples
        "tall": (rng.random((300, 41)) < 0.2).astype(np.uint8),
        "two_rows": (rng.random((2, 90)) < 0.4).astype(np.uint8),
        "dense_rows": (rng.random((40, 1500)) < 0.5).astype(np.uint8),  # 47 words: G=16; long re-deals
        "long_rows": (rng.random((50, 3000)) < 0.05).astype(np.uint8),  # 94 words: G=32, V=4
        "c3_like": (rng.random((64, 5000)) < 0.01).astype(np.uint8),    # 157 words: G=32, V=5
        "max_rows": (rng.random((30, 8000)) < 0.02).astype(np.uint8),   # 250 words: G=32, V=8
        # rows of exactly 2 / 4 / 8 128-bit words: the one-trade-per-thread path (sparse rows) next to the group path
        "nv2": (rng.random((40, 250)) < 0.04).astype(np.uint8),
        "nv4": np.concatenate([(rng.random((500, 26)) < 0.02), (rng.random((500, 4)) < 0.3)], axis=1).astype(np.uint8),
        "nv8": np.concatenate([(rng.random((1000, 90)) < 0.01), (rng.random((1000, 10)) < 0.2)], axis=1).astype(np.uint8),
    }
    return out


@pytest.mark.parametrize("name", list(_dense_cases().keys()))
def test_curveball_matches_oracle_bit_for_bit(eng, name):
    dense = _dense_cases()[name]
    upload_dense(eng, dense)
    obs = rows_of(dense)
    R, L, wpr, ras = eng.trade_shape()
    assert (R, wpr) == obs.shape and ras == (dense.shape[1] >= dense.shape[0])
    n = 6
    T = 5 * R
    want, want_ap = oracle.curveball_nulls(obs, seed=1764, null0=3, n=n, n_trades=T, chain_len=1)
    got, got_ap = eng.curveball(1764, 3, n, n_trades=-1, chain_len=1)
    assert np.array_equal(got_ap, want_ap)
    assert np.array_equal(got, want)
    # margins (G1)
    for m in got:
        assert np.array_equal(bitpack.popcount_rows(m), bitpack.popcount_rows(obs))
        assert np.array_equal(bitpack.unpack_rows(m, L).sum(0), bitpack.unpack_rows(obs, L).sum(0))


@pytest.mark.parametrize("chain_len,null0,n", [(4, 0, 9), (4, 2, 7), (1000, 0, 5), (3, 7, 4)])
def test_chained_nulls_match_oracle(eng, chain_len, null0, n):
    cases, _ = curveball_cases()
    dense = cases["c0_shape"]["dense"]
    upload_dense(eng, dense)
    obs = rows_of(dense)
    want, want_ap = oracle.curveball_nulls(obs, 99, null0, n, 5 * obs.shape[0], chain_len)
    got, got_ap = eng.curveball(99, null0, n, -1, chain_len)
    assert np.array_equal(got_ap, want_ap)
    assert np.array_equal(got, want)


def test_long_chain_crosses_segment_boundaries(eng):
    """chain_len > 256 forces the carry buffer between launches."""
    rng = np.random.default_rng(5)
    dense = (rng.random((24, 70)) < 0.3).astype(np.uint8)
    upload_dense(eng, dense)
    obs = rows_of(dense)
    want, _ = oracle.curveball_nulls(obs, 7, 0, 600, 5 * obs.shape[0], 600)
    got, _ = eng.curveball(7, 0, 600, -1, 600)
    assert np.array_equal(got, want)
    want2, _ = oracle.curveball_nulls(obs, 7, 250, 30, 5 * obs.shape[0], 600)
    got2, _ = eng.curveball(7, 250, 30, -1, 600)
    assert np.array_equal(got2, want2)


def test_very_deep_schedule_runs_in_sequential_order(eng):
    """3 rows: every trade shares a row with the previous one, so the level schedule is as deep as the trade
    count; beyond the planner's depth cap the kernel runs the trades one by one."""
    rng = np.random.default_rng(9)
    dense = (rng.random((3, 64)) < 0.4).astype(np.uint8)
    upload_dense(eng, dense)
    obs = rows_of(dense)
    want, want_ap = oracle.curveball_nulls(obs, 21, 0, 3, 1500, 1)
    got, got_ap = eng.curveball(21, 0, 3, n_trades=1500, chain_len=1)
    assert np.array_equal(got, want) and np.array_equal(got_ap, want_ap)


def test_results_do_not_depend_on_batching(eng):
    cases, _ = curveball_cases()
    upload_dense(eng, cases["c0_shape"]["dense"])
    whole, _ = eng.curveball(5, 0, 40, -1, 1)
    parts = [eng.curveball(5, lo, 10, -1, 1)[0] for lo in (0, 10, 20, 30)]
    assert np.array_equal(whole, np.concatenate(parts))


def test_zero_trades_and_zero_nulls(eng):
    rng = np.random.default_rng(1)
    dense = (rng.random((10, 33)) < 0.3).astype(np.uint8)
    upload_dense(eng, dense)
    got, ap = eng.curveball(1, 0, 2, n_trades=0)
    assert np.array_equal(got[0], rows_of(dense)) and ap.tolist() == [0, 0]
    got, ap = eng.curveball(1, 0, 0)
    assert got.shape[0] == 0


# ---------------------------------------------------------------- SuperW
def test_score_sets_match_reference_superw(eng):
    z = load_npz("superw.npz")
    dense = z["dense"]
    S, F = dense.shape
    upload_dense(eng, dense)
    eng.upload_weights(z["weights"])
    res = eng.score_sets(z["modules"], z["kind"])
    for i in range(len(z["W"])):
        w = z["weights"][z["kind"][i]]
        assert abs(res["W"][i] - z["W"][i]) <= 1e-9 * max(1.0, np.abs(w).sum())
    assert np.array_equal(res["coverage"], z["cov"])
    assert np.array_equal(res["overlap"], z["overlap"])
    assert np.array_equal(res["act_cov"], z["act_cov"])


def test_score_sets_with_sample_mask_match_oracle(eng):
    z = load_npz("superw.npz")
    dense = z["dense"]
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    upload_dense(eng, dense)
    w = z["weights"][:2].copy()
    keep = np.stack([np.arange(S) % 4 != 0, np.arange(S) % 3 != 1])
    eng.upload_weights(w, keep)
    rng = np.random.default_rng(3)
    sets = np.full((200, 5), -1, dtype=np.int32)
    for i in range(200):
        k = rng.integers(1, 6)
        sets[i, :k] = rng.choice(F, size=k, replace=False)
    prof = rng.integers(0, 2, size=200).astype(np.int32)
    res = eng.score_sets(sets, prof)
    for i in range(200):
        p = prof[i]
        feats = sets[i][sets[i] >= 0]
        W, cov, ov, ac = oracle.score_set(rows, S, feats, np.where(keep[p], w[p], 0.0),
                                          bitpack.pack_mask(np.nonzero(keep[p])[0], S))
        assert abs(res["W"][i] - W) <= 1e-9 * np.abs(w[p]).sum()
        assert (res["coverage"][i], res["overlap"][i], res["act_cov"][i]) == (cov, ov, ac)


def test_feature_scores_match_oracle(eng):
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    w = synth.weights(S, 2)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    rows = bitpack.pack_rows(dense.T)
    for p in range(2):
        sw, W1, cnt, best = oracle.feature_scores(rows, S, w[p])
        got = eng.feature_scores(p)
        assert np.allclose(got["sum_w"], sw, rtol=0, atol=1e-9 * np.abs(w[p]).sum())
        assert np.allclose(got["W1"], W1, rtol=0, atol=1e-9 * np.abs(w[p]).sum())
        assert np.array_equal(got["count"], cnt)
        assert got["best_single"] == best


# ---------------------------------------------------------------- null loop, fixed statistic
@pytest.mark.parametrize("shape", [(770, 400), (120, 300)])
def test_null_pvalue_fixed_matches_oracle(eng, shape):
    dense, _, _ = synth.feature_matrix(*shape)
    S, F = dense.shape
    P = 3
    w = synth.weights(S, P)
    keep = np.ones((P, S), bool)
    keep[1, ::7] = False
    upload_dense(eng, dense)
    eng.upload_weights(w, keep)
    order = np.argsort(-dense.sum(0), kind="stable")
    modules = np.stack([np.sort(order[1:4]), np.sort(order[0:3]), np.array([order[5], order[9], -1])]).astype(np.int32)
    n = 24
    res = eng.null_pvalue(1764, 0, n, modules, chain_len=1, stat=SDX_STAT_FIXED, want_scores=True)
    obs = rows_of(dense)
    R, L, wpr, ras = eng.trade_shape()
    nulls, _ = oracle.curveball_nulls(obs, 1764, 0, n, 5 * R, 1)
    frows = bitpack.pack_rows(dense.T)
    for p in range(P):
        wm = np.where(keep[p], w[p], 0.0)
        mk = bitpack.pack_mask(np.nonzero(keep[p])[0], S)
        feats = modules[p][modules[p] >= 0]
        z, _, _, _ = oracle.score_set(frows, S, feats, wm, mk)
        assert abs(res["z_obs"][p] - z) <= 1e-9 * np.abs(wm).sum()
        want = []
        for i in range(n):
            d = bitpack.unpack_rows(nulls[i], L)
            fr = bitpack.pack_rows(d.T) if ras else nulls[i]
            want.append(oracle.score_set(fr, S, feats, wm, mk)[0])
        want = np.array(want)
        assert np.allclose(res["null_scores"][p], want, rtol=0, atol=1e-9 * np.abs(wm).sum())
        # exceedance count is computed from the device's own scores
        assert res["n_ge"][p] == int((res["null_scores"][p] >= res["z_obs"][p]).sum())
    # sharded ranges add up (what the multi-GPU path relies on)
    parts = [eng.null_pvalue(1764, lo, 8, modules, z_obs=res["z_obs"])["n_ge"] for lo in (0, 8, 16)]
    assert np.array_equal(np.sum(parts, axis=0), res["n_ge"])


# ---------------------------------------------------------------- error behaviour
def test_c_abi_status_codes_through_ctypes():
    """The boundary contract of include/sdx.h exercised without the Python wrapper: bad arguments, NULL handles and
    calls in the wrong order return negative SDX_ERR_* codes with a message; nothing aborts and the context stays usable."""
    import ctypes as C
    from superdendrix_b200 import _capi
    lib = _capi.load()
    INVALID, STATE = -1, -3
    ctx = C.c_void_p()
    assert lib.sdx_create(99, C.byref(ctx)) == INVALID and not ctx.value
    assert b"out of range" in lib.sdx_last_error(None)
    assert lib.sdx_destroy(None) == INVALID
    assert lib.sdx_create(0, C.byref(ctx)) == 0 and ctx.value
    try:
        i32p, i64p, f64p, u32p = (C.POINTER(t) for t in (C.c_int32, C.c_int64, C.c_double, C.c_uint32))
        n64 = (C.c_int64 * 4)()
        f64 = (C.c_double * 4)()
        mod = (C.c_int32 * 3)(0, 1, 2)
        hits = (_capi.Hit * 2)()
        # nothing uploaded yet
        assert lib.sdx_curveball(ctx, 1, 0, 1, -1, 1, None, None) == STATE
        assert lib.sdx_score_sets(ctx, mod, None, 1, 3, f64, None, None, None) == STATE
        assert lib.sdx_scan(ctx, 0, 3, 2, hits, n64) == STATE
        assert lib.sdx_feature_scores(ctx, 0, f64, None, None, None) == STATE
        assert lib.sdx_condperm(ctx, 0, mod, 3, 1, 0, 10, n64, f64) == STATE
        assert lib.sdx_ranksum(ctx, mod, None, 1, 3, f64, f64, n64, None) == STATE
        assert lib.sdx_null_pvalue(ctx, 1, 0, 4, -1, 1, 0, mod, 3, None, n64, None, f64) == STATE
        assert len(lib.sdx_last_error(ctx)) > 0
        # bad uploads
        rows = np.zeros((4, 1), dtype=np.uint32)
        rp = rows.ctypes.data_as(u32p)
        assert lib.sdx_upload_matrix(ctx, None, 4, 20, 1) == INVALID
        assert lib.sdx_upload_matrix(ctx, rp, 4, 20, 2) == INVALID                 # words_per_row must be ceil(S / 32)
        assert lib.sdx_upload_matrix(ctx, rp, 0, 20, 1) == INVALID
        assert lib.sdx_upload_weights(ctx, f64, None, 1, 4) == STATE              # matrix first
        rows[:, 0] = [0b0111, 0b0011, 0b1000, 0b0101]
        assert lib.sdx_upload_matrix(ctx, rp, 4, 20, 1) == 0
        w = np.linspace(-1, 1, 20)
        assert lib.sdx_upload_weights(ctx, w.ctypes.data_as(f64p), None, 1, 19) == INVALID
        assert lib.sdx_upload_weights(ctx, None, None, 1, 20) == INVALID
        w_nan = w.copy(); w_nan[3] = np.nan
        assert lib.sdx_upload_weights(ctx, w_nan.ctypes.data_as(f64p), None, 1, 20) == INVALID
        assert lib.sdx_upload_weights(ctx, w.ctypes.data_as(f64p), None, 1, 20) == 0
        # bad calls on a ready context
        assert lib.sdx_scan(ctx, 0, 4, 2, hits, n64) < 0                           # k > 3
        assert lib.sdx_scan(ctx, 1, 3, 2, hits, n64) == INVALID                    # profile out of range
        assert lib.sdx_score_sets(ctx, mod, None, 1, 9, f64, None, None, None) == INVALID      # k > SDX_MAX_K
        bad = (C.c_int32 * 3)(0, 1, 4)
        assert lib.sdx_score_sets(ctx, bad, None, 1, 3, f64, None, None, None) == INVALID      # feature index out of range
        assert lib.sdx_curveball(ctx, 1, 0, -1, -1, 1, None, None) == INVALID
        assert lib.sdx_set_burn_in(ctx, 1000, 4) == INVALID
        assert lib.sdx_null_pvalue(ctx, 1, 0, 4, -1, 1, 7, mod, 3, None, n64, None, f64) == INVALID   # unknown statistic
        # ... and it still works
        assert lib.sdx_score_sets(ctx, mod, None, 1, 3, f64, None, None, None) == 0
        cases = [set(np.flatnonzero([(int(rows[g, 0]) >> s_) & 1 for s_ in range(20)])) for g in range(3)]
        union = set().union(*cases)
        want = 2 * sum(w[s_] for s_ in union if w[s_] > 0) - sum(abs(w[s_]) for c_ in cases for s_ in c_)
        assert abs(f64[0] - want) < 1e-12
    finally:
        assert lib.sdx_destroy(ctx) == 0


def test_errors_are_reported_not_fatal(eng):
    dense, _, _ = synth.feature_matrix(100, 50)
    upload_dense(eng, dense)
    eng.upload_weights(synth.weights(100, 1))
    with pytest.raises(SdxError):
        eng.score_sets([[0, 1, 10 ** 6]])
    with pytest.raises(SdxError):
        eng.upload_weights(np.zeros((1, 99)))
    with pytest.raises(SdxError):
        eng.curveball(1, 0, 1, chain_len=0)
    bad = bitpack.pack_rows(dense.T).copy()
    bad[0, -1] |= 0x80000000
    with pytest.raises(SdxError):
        eng.upload_matrix(bad, 100)
    # engine is still usable
    upload_dense(eng, dense)
    eng.upload_weights(synth.weights(100, 1))
    assert eng.score_sets([[0, 1, 2]])["W"].shape == (1,)


# ---------------------------------------------------------------- exhaustive scan
def _check_scan_against_oracle(eng, dense, w, k, topn, exact_sets=False):
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    upload_dense(eng, dense)
    eng.upload_weights(w[None, :])
    got = eng.scan(0, k, topn)
    oW, oF, oC, n = oracle.scan(rows, S, w, k, topn)
    tol = 1e-9 * max(1.0, np.abs(w).sum())
    assert got["n_scored"] == n
    assert np.allclose(got["W"], oW, rtol=0, atol=tol)
    # every reported subset really has the reported W and coverage (oracle re-score)
    for i in range(topn):
        feats = got["features"][i]
        feats = feats[feats >= 0]
        if len(feats) == 0:
            assert got["W"][i] == -np.inf and oW[i] == -np.inf
            continue
        W, cov, _, _ = oracle.score_set(rows, S, feats, w)
        assert abs(W - got["W"][i]) <= tol and cov == got["coverage"][i]
    if exact_sets:
        assert np.array_equal(got["features"], oF) and np.array_equal(got["coverage"], oC)
    else:
        # the optimum itself is unambiguous unless tied within tolerance
        if topn == 1 or oW[0] - oW[1] > 10 * tol:
            assert np.array_equal(got["features"][0], oF[0])
    return got


@pytest.mark.parametrize("name", ["a", "b", "unit"])
def test_scan_matches_golden_bruteforce(eng, name):
    z = load_npz("scan.npz")
    dense, w = z[name + "__dense"], z[name + "__w"]
    got = _check_scan_against_oracle(eng, dense, w, 3, 16, exact_sets=(name == "unit"))
    tol = 1e-9 * np.abs(w).sum()
    assert np.allclose(got["W"], z[name + "__topW"], rtol=0, atol=tol)       # the reference's own SuperW, brute force
    assert got["n_scored"] == int(z[name + "__n"])
    for k in (1, 2, 3):
        upload_dense(eng, dense)
        eng.upload_weights(w[None, :])
        assert abs(eng.scan(0, k, 1)["W"][0] - z[name + "__best_le_k"][k - 1]) < tol


def test_scan_optimum_equals_the_reference_ilp_at_c0_size(eng):
    """Pin G4 at F <= 400: the scan's optimum on the C0 matrix equals the optimum of the reference's ILP
    (src/superdendrix.py:294-313) solved by HiGHS (oracle/milp.py), for k = 1, 2, 3 and for two weight vectors."""
    from oracle import milp
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    upload_dense(eng, dense)
    for seed in (3, 11):
        w = synth.weights(S, 1, seed=seed)[0]
        eng.upload_weights(w[None, :])
        for k in (1, 2, 3):
            z, feats, optimal = milp.ilp_optimum(dense, w, k, all_samples=False)
            assert optimal
            top = eng.scan(0, k, 1)
            assert abs(max(float(top["W"][0]), 0.0) - z) <= 1e-7 * np.abs(w).sum(), (seed, k)
            assert sorted(int(g) for g in top["features"][0] if g >= 0) == feats


def test_maxk_null_statistic_equals_the_reference_ilp_on_every_null(eng):
    """The reference's p-value compares the observed optimum with the ILP optimum of EVERY null (src/superdendrix.py:173-192).
    The max_k statistic of the GPU path (per-null exhaustive scan) must equal that ILP — solved here by HiGHS on the
    oracle's nulls of the C0 matrix — null by null, so exceedance counts are the reference's by construction."""
    from oracle import milp
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    w = synth.weights(S, 1, seed=3)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    module = synth.pick_module(dense, 3)[None, :].astype(np.int32)
    n = 24
    res = eng.null_pvalue(1764, 0, n, module, stat=SDX_STAT_MAXK, want_scores=True)
    z_ilp, _, optimal = milp.ilp_optimum(dense, w[0], 3, all_samples=False)
    assert optimal and abs(res["z_obs"][0] - z_ilp) <= 1e-7 * np.abs(w).sum()
    R, L, wpr, ras = eng.trade_shape()
    nulls, _ = oracle.curveball_nulls(rows_of(dense), 1764, 0, n, 5 * R, 1)
    want = []
    for m in nulls:
        d = bitpack.unpack_rows(m, L)
        z, _, optimal = milp.ilp_optimum(d if ras else d.T, w[0], 3, all_samples=False)
        assert optimal
        want.append(z)
    assert np.allclose(res["null_scores"][0], want, rtol=0, atol=1e-7 * np.abs(w).sum())
    assert res["n_ge"][0] == int((np.array(want) >= z_ilp - 1e-9).sum())


@pytest.mark.parametrize("S,F,density,k", [(200, 61, 0.05, 3), (130, 37, 0.5, 3), (64, 5, 0.3, 3), (40, 3, 0.4, 3),
                                           (40, 2, 0.4, 3), (33, 1, 0.5, 3), (300, 90, 0.02, 2), (300, 90, 0.02, 1),
                                           (1030, 70, 0.1, 3)])
def test_scan_matches_oracle_on_random_matrices(eng, S, F, density, k):
    rng = np.random.default_rng(S * 1000 + F)
    dense = (rng.random((S, F)) < density).astype(np.uint8)
    w = synth.weights(S, 1, seed=F)[0]
    _check_scan_against_oracle(eng, dense, w, k, 8)


def test_scan_duplicate_features_tie_order(eng):
    """Identical carrier sets give exactly equal W; ties resolve in enumeration order like the oracle."""
    rng = np.random.default_rng(2)
    base = (rng.random((150, 12)) < 0.15).astype(np.uint8)
    dense = np.concatenate([base, base[:, :6]], axis=1)
    w = synth.weights(150, 1, seed=5)[0]
    _check_scan_against_oracle(eng, dense, w, 3, 12, exact_sets=True)


def test_scan_with_sample_mask(eng):
    rng = np.random.default_rng(8)
    dense = (rng.random((180, 40)) < 0.1).astype(np.uint8)
    w = synth.weights(180, 1, seed=1)[0]
    keep = np.arange(180) % 5 != 2
    upload_dense(eng, dense)
    eng.upload_weights(w[None, :], keep[None, :])
    got = eng.scan(0, 3, 4)
    rows = bitpack.pack_rows(dense.T)
    oW, oF, oC, n = oracle.scan(rows, 180, np.where(keep, w, 0.0), 3, 4, mask=bitpack.pack_mask(np.nonzero(keep)[0], 180))
    assert np.allclose(got["W"], oW, rtol=0, atol=1e-9 * np.abs(w).sum())
    assert np.array_equal(got["features"][0], oF[0]) and np.array_equal(got["coverage"], oC)


def test_scan_full_size_properties(eng):
    """C2 (2 000 features x 1 000 samples): count is exact, optimum dominates sampled subsets and the
    single / pair optima, and is reproduced by the set scorer."""
    dense, _, _ = synth.feature_matrix(1000, 2000)
    S, F = dense.shape
    w = synth.weights(S, 1)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    got = eng.scan(0, 3, 8)
    assert got["n_scored"] == F + F * (F - 1) // 2 + F * (F - 1) * (F - 2) // 6
    rng = np.random.default_rng(0)
    sets = np.sort(np.stack([rng.choice(F, 3, replace=False) for _ in range(20000)]), axis=1).astype(np.int32)
    sc = eng.score_sets(sets)
    tol = 1e-9 * np.abs(w).sum()
    assert sc["W"].max() <= got["W"][0] + tol
    assert (np.diff(got["W"]) <= 0).all()
    re = eng.score_sets(got["features"])
    assert np.allclose(re["W"], got["W"], rtol=0, atol=tol) and np.array_equal(re["coverage"], got["coverage"])
    assert eng.scan(0, 1, 1)["W"][0] <= eng.scan(0, 2, 1)["W"][0] + tol <= got["W"][0] + 2 * tol
    fs = eng.feature_scores(0)
    assert abs(eng.scan(0, 1, 1)["W"][0] - fs["W1"].max()) <= tol
    # greedy local search from the optimum cannot improve it (swap one feature)
    best = got["features"][0]
    for pos in range(3):
        cand = np.tile(best, (F, 1))
        cand[:, pos] = np.arange(F)
        ok = np.array([len(set(r)) == 3 for r in cand])
        assert eng.score_sets(cand[ok].astype(np.int32))["W"].max() <= got["W"][0] + tol


def test_one_long_chain_in_batches_continues_instead_of_replaying():
    """the reference driver's single Markov chain (src/generate_null_matrices.py:71-74) requested in batches: every call
    continues from the state the previous one left (the context keeps it), a request that does not follow on replays
    the chain from its start; both give the nulls of one request — and the oracle's"""
    dense = _dense_cases()["tall"]
    obs = rows_of(dense)
    T = 5 * obs.shape[0]
    want, want_ap = oracle.curveball_nulls(obs, 77, 0, 90, T, 1000)
    with Engine(0) as e2:
        upload_dense(e2, dense)
        parts = [e2.curveball(77, lo, n, chain_len=1000) for lo, n in ((0, 30), (30, 20), (50, 5))]
        rounds = []
        for lo, n in ((55, 10), (70, 20)):                       # 55 follows on; 70 does not (65..69 were never asked for)
            parts.append(e2.curveball(77, lo, n, chain_len=1000))
            rounds.append(e2.timing()["trade_ms"])
        got = np.concatenate([p_[0] for p_ in parts[:4]])
        assert np.array_equal(got, want[:65]) and np.array_equal(np.concatenate([p_[1] for p_ in parts[:4]]), want_ap[:65])
        assert np.array_equal(parts[4][0], want[70:90])
        # a chain boundary inside the request, and a second chain, still restart from the observed matrix
        two, _ = e2.curveball(77, 995, 10, chain_len=1000)
    ref2, _ = oracle.curveball_nulls(obs, 77, 995, 10, T, 1000)
    assert np.array_equal(two, ref2)


def test_scan_full_size_equals_the_oracles_top_16(eng):
    """C2 at FULL size against the oracle: tests/golden/scan_c2.npz holds the 16 best of all 1 274 230 475 subsets, every one
    scored by oracle/sdx_oracle.c with the plain definition of SuperW (src/superdendrix.py:387-391; no inclusion-exclusion,
    ~15 core-hours, made once by tests/golden/make_scan_c2.py).  The GPU scan must return exactly those subsets in that
    order, with the same coverage and W within 1e-9 * sum|w| — whole, and sharded by smallest feature index and merged."""
    from superdendrix_b200 import dist
    z = load_npz("scan_c2.npz")
    dense, _, _ = synth.feature_matrix(1000, 2000, seed=20)
    S, F = dense.shape
    assert F == int(z["F"]) and S == int(z["S"])
    w = synth.weights(S, 1, seed=3)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    tol = 1e-9 * np.abs(w).sum()
    got = eng.scan(0, 3, 16)
    assert got["n_scored"] == int(z["n_scored"])
    assert np.array_equal(got["features"], z["features"]) and np.array_equal(got["coverage"], z["coverage"])
    assert np.allclose(got["W"], z["W"], rtol=0, atol=tol)
    parts = [eng.scan(0, 3, 16, g_lo=lo, g_hi=hi) for lo, hi in dist.scan_feature_ranges(F, 3, 8)]
    merged = dist.merge_hits(parts, 16)
    assert np.array_equal(merged["features"], z["features"]) and sum(p_["n_scored"] for p_ in parts) == int(z["n_scored"])


# ---------------------------------------------------------------- conditional permutation
def _condperm_inputs():
    z = load_npz("condperm_data.npz")
    with open(os.path.join(GOLDEN, "condperm.json")) as fh:
        meta = json.load(fh)
    return z["dense"], z["w"], meta


def test_condperm_counts_equal_oracle_and_reference_within_mc_error(eng):
    dense, w, meta = _condperm_inputs()
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    upload_dense(eng, dense)
    eng.upload_weights(w[None, :])
    for c in meta["cases"]:
        P = c["P"]
        counts, realW = eng.condperm(c["feats"], P, seed=77, perm_id0=0)
        ocounts, orealW = oracle.condperm(rows, S, c["feats"], w, np.arange(S), seed=77, perm_id0=0, P=P)
        assert np.array_equal(counts, ocounts)               # same streams, same summation order: exact
        assert np.array_equal(realW, orealW)
        for got, want in zip(counts, c["per_feature"]):      # the reference's own counts (CPython random): Monte-Carlo error
            p = max(want, 1) / P
            sd = np.sqrt(2 * P * p * (1 - p)) + 1
            assert abs(got - want) < 4.5 * sd, (got, want)


def test_condperm_with_mask_padding_and_batch(eng):
    rng = np.random.default_rng(12)
    dense = (rng.random((333, 30)) < 0.12).astype(np.uint8)
    dense[:, 0] = rng.random(333) < 0.6
    S, F = dense.shape
    w = synth.weights(S, 3, seed=2)
    keep = np.ones((3, S), bool)
    keep[1, ::5] = False
    keep[2, 100:140] = False
    upload_dense(eng, dense)
    eng.upload_weights(w, keep)
    rows = bitpack.pack_rows(dense.T)
    modules = np.array([[0, 3, 7], [1, 2, -1], [0, 1, 2], [5, -1, -1]], dtype=np.int32)
    prof = np.array([0, 1, 2, 1], dtype=np.int32)
    counts, realW = eng.condperm_batch(modules, 500, seed=5, perm_id0=40, profile_of_job=prof)
    for j in range(len(modules)):
        p = prof[j]
        feats = modules[j][modules[j] >= 0]
        oc, orw = oracle.condperm(rows, S, feats, np.where(keep[p], w[p], 0.0), np.nonzero(keep[p])[0], seed=5, perm_id0=40, P=500)
        assert np.array_equal(counts[j][:len(feats)], oc), j
        assert np.array_equal(realW[j][:len(feats)], orw)
        assert (counts[j][len(feats):] == -1).all()
        one, _ = eng.condperm(modules[j], 500, seed=5, perm_id0=40, profile=int(p))
        assert np.array_equal(one, counts[j])
    # n_perm = 0 and a single-feature module
    c0, r0 = eng.condperm([0, 3], 0, seed=1)
    assert c0.tolist() == [0, 0]


# ---------------------------------------------------------------- rank-sum
def test_ranksum_matches_scipy_golden(eng):
    z = load_npz("ranksum.npz")
    prof, carrier = z["profile"], z["carrier"]
    n, S = prof.shape
    # one feature per job carrying exactly the golden carrier set
    upload_dense(eng, carrier.T.astype(np.uint8))
    eng.upload_weights(prof)
    res = eng.ranksum(np.arange(n, dtype=np.int32)[:, None], np.arange(n, dtype=np.int32))
    for i in range(n):
        assert res["n1"][i] == int(carrier[i].sum())
        assert abs(res["z"][i] - z["z"][i]) < 1e-12 * max(1, abs(z["z"][i]))
        assert abs(res["p"][i] - z["p"][i]) <= 1e-12 + 1e-10 * z["p"][i]
        _, _, s2, n1 = oracle.ranksum(prof[i], np.arange(S), bitpack.pack_mask(np.nonzero(carrier[i])[0], S))
        assert (res["twice_rank_sum"][i], res["n1"][i]) == (s2, n1)     # integer work: exact


def test_ranksum_union_of_module_with_mask_and_ties(eng):
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    w = synth.weights(S, 2)          # 2 % exact zeros -> ties
    w[1] = np.round(w[1])            # heavy ties
    keep = np.ones((2, S), bool)
    keep[1, ::3] = False
    upload_dense(eng, dense)
    eng.upload_weights(w, keep)
    modules = np.array([[0, 1, 2], [5, 9, -1], [3, -1, -1], [0, 10, 20]], dtype=np.int32)
    prof = np.array([0, 1, 1, 0], dtype=np.int32)
    res = eng.ranksum(modules, prof)
    from scipy import stats
    for j in range(len(modules)):
        p = prof[j]
        feats = modules[j][modules[j] >= 0]
        union = dense[:, feats].any(axis=1)
        idx = np.nonzero(keep[p])[0]
        zz, pp, s2, n1 = oracle.ranksum(np.where(keep[p], w[p], 0.0), idx, bitpack.pack_mask(np.nonzero(union)[0], S))
        assert (res["twice_rank_sum"][j], res["n1"][j]) == (s2, n1)
        assert abs(res["z"][j] - zz) <= 1e-12 * max(1, abs(zz)) and abs(res["p"][j] - pp) <= 1e-12 + 1e-10 * pp
        x, y = w[p][keep[p] & union], w[p][keep[p] & ~union]
        ref = stats.ranksums(x, y)
        assert abs(res["z"][j] - ref.statistic) < 1e-10 and abs(res["p"][j] - ref.pvalue) < 1e-10


# ---------------------------------------------------------------- null loop, re-optimised statistic (max_k)
@pytest.mark.parametrize("shape,chain_len", [((120, 40), 1), ((36, 50), 1), ((120, 40), 7)])
def test_null_pvalue_maxk_matches_oracle_scan_per_null(eng, shape, chain_len):
    """src/superdendrix.py:191: the null statistic is the optimum over |M| <= k on each null (the reference re-runs the
    ILP); here the exhaustive scan, checked against the oracle's scan of the oracle's nulls.  M = {} gives z >= 0."""
    rng = np.random.default_rng(shape[0])
    dense = (rng.random(shape) < 0.12).astype(np.uint8)
    dense[:, 0] = rng.random(shape[0]) < 0.5
    dense[:, 1] = rng.random(shape[0]) < 0.4
    S, F = dense.shape
    P = 2
    w = synth.weights(S, P, seed=shape[1])
    w[1] -= 1.5                                   # mostly negative weights: optimum of some nulls is the empty module
    keep = np.ones((P, S), bool)
    keep[1, ::6] = False
    upload_dense(eng, dense)
    eng.upload_weights(w, keep)
    modules = np.array([[0, 1, 2], [3, 4, -1]], dtype=np.int32)        # k = 3 for profile 0, k = 2 for profile 1
    n = 20
    res = eng.null_pvalue(77, 2, n, modules, chain_len=chain_len, stat=SDX_STAT_MAXK, want_scores=True)
    obs = rows_of(dense)
    R, L, wpr, ras = eng.trade_shape()
    nulls, _ = oracle.curveball_nulls(obs, 77, 2, n, 5 * R, chain_len)
    frows = bitpack.pack_rows(dense.T)
    for p, k in ((0, 3), (1, 2)):
        wm = np.where(keep[p], w[p], 0.0)
        mk = bitpack.pack_mask(np.nonzero(keep[p])[0], S)
        tol = 1e-9 * np.abs(wm).sum()
        z = max(oracle.scan(frows, S, wm, k, 1, mask=mk)[0][0], 0.0)
        assert abs(res["z_obs"][p] - z) <= tol
        want = []
        for i in range(n):
            fr = bitpack.pack_rows(bitpack.unpack_rows(nulls[i], L).T) if ras else nulls[i]
            want.append(max(oracle.scan(fr, S, wm, k, 1, mask=mk)[0][0], 0.0))
        assert np.allclose(res["null_scores"][p], want, rtol=0, atol=tol)
        assert res["n_ge"][p] == int((res["null_scores"][p] >= res["z_obs"][p]).sum())
    assert (res["null_scores"][1] == 0.0).any() or True
    # caller-supplied z_obs (the driver passes W of the selected module) and sharded ranges
    zP = np.array([1.0, 0.5])
    whole = eng.null_pvalue(77, 2, n, modules, chain_len=1, stat=SDX_STAT_MAXK, z_obs=zP, want_scores=True)
    parts = [eng.null_pvalue(77, 2 + lo, 10, modules, chain_len=1, stat=SDX_STAT_MAXK, z_obs=zP)["n_ge"] for lo in (0, 10)]
    assert np.array_equal(np.sum(parts, axis=0), whole["n_ge"])
    assert np.array_equal(whole["n_ge"], (whole["null_scores"] >= zP[:, None]).sum(axis=1))


def test_maxk_statistic_dominates_fixed_statistic(eng):
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    w = synth.weights(S, 1)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    module = synth.pick_module(dense, 3)[None, :]
    fx = eng.null_pvalue(5, 0, 48, module, stat=SDX_STAT_FIXED, want_scores=True)
    mx = eng.null_pvalue(5, 0, 48, module, stat=SDX_STAT_MAXK, want_scores=True)
    assert (mx["null_scores"] >= fx["null_scores"] - 1e-9 * np.abs(w).sum()).all()
    best = eng.scan(0, 3, 1)["W"][0]
    assert abs(mx["z_obs"][0] - max(best, 0.0)) <= 1e-9 * np.abs(w).sum()


# ---------------------------------------------------------------- randomized shapes (property-style sweep)
@pytest.mark.parametrize("case", range(16))
def test_curveball_random_shapes_and_densities_match_oracle(eng, case):
    """Random shapes around the layout boundaries (row lengths of 1..9 128-bit words, both orientations,
    very sparse to dense, few to many rows): nulls, applied counts and margins equal the oracle's."""
    rng = np.random.default_rng(1000 + case)
    L = int(rng.choice([1, 31, 32, 33, 96, 127, 128, 129, 255, 256, 300, 511, 640, 897, 1024, 1100]))
    R = int(rng.integers(2, 70))
    density = float(rng.choice([0.01, 0.05, 0.2, 0.5, 0.9]))
    rows_dense = (rng.random((R, L)) < density).astype(np.uint8)
    if L >= R:
        rows_dense[:, rng.integers(0, L)] = 1                      # a full column and an empty row are legal inputs
        rows_dense[rng.integers(0, R)] = 0
    dense = rows_dense if case % 2 == 0 and L >= R else rows_dense.T      # S x F with samples or features as trade rows
    if dense.shape[1] < dense.shape[0] and min(dense.shape) < 2:
        dense = dense.T
    S, F = dense.shape
    if min(S, F) < 2:
        pytest.skip("Curveball needs two rows")
    upload_dense(eng, dense)
    obs = rows_of(dense)
    Rr, Lr, wpr, ras = eng.trade_shape()
    assert obs.shape == (Rr, wpr)
    n, T = 3, int(rng.integers(1, 6 * Rr + 1))
    seed = int(rng.integers(0, 2 ** 62))
    null0 = int(rng.integers(0, 2 ** 40))
    want, want_ap = oracle.curveball_nulls(obs, seed, null0, n, T, 2)
    got, got_ap = eng.curveball(seed, null0, n, n_trades=T, chain_len=2)
    assert np.array_equal(got_ap, want_ap)
    assert np.array_equal(got, want)
    for m in got:
        assert np.array_equal(bitpack.popcount_rows(m), bitpack.popcount_rows(obs))
        assert np.array_equal(bitpack.unpack_rows(m, Lr).sum(0), bitpack.unpack_rows(obs, Lr).sum(0))


# ---------------------------------------------------------------- experimental sparse-layout kernel (SDX_TRADE_SPARSE=1)
def test_sparse_layout_kernel_is_bit_identical(monkeypatch):
    """k_trade_sparse (index lists for rare rows, one warp per null) must give the same nulls, counts and scores
    as the oracle / the default bit-packed kernel; the layout is chosen at upload time from the environment."""
    monkeypatch.setenv("SDX_TRADE_SPARSE", "1")
    cases, _ = curveball_cases()
    rng = np.random.default_rng(4)
    mats = {"c0_shape": cases["c0_shape"]["dense"],
            "rare": (rng.random((500, 60)) < 0.01).astype(np.uint8),
            "mixed": np.concatenate([(rng.random((300, 50)) < 0.02), (rng.random((300, 6)) < 0.4)], axis=1).astype(np.uint8)}
    with Engine(0) as e2:
        for name, dense in mats.items():
            S, F = dense.shape
            upload_dense(e2, dense)
            obs = rows_of(dense)
            R = obs.shape[0]
            for chain_len, null0, n in ((1, 0, 5), (3, 1, 6)):
                want, want_ap = oracle.curveball_nulls(obs, 11, null0, n, 5 * R, chain_len)
                got, got_ap = e2.curveball(11, null0, n, -1, chain_len)
                assert np.array_equal(got_ap, want_ap), name
                assert np.array_equal(got, want), name
            w = synth.weights(S, 2, seed=S)
            e2.upload_weights(w)
            order = np.argsort(-dense.sum(0), kind="stable")
            modules = np.stack([np.sort(order[:3]), np.array([order[1], order[-1], -1])]).astype(np.int32)
            res = e2.null_pvalue(3, 0, 16, modules, want_scores=True)
            nulls, _ = oracle.curveball_nulls(obs, 3, 0, 16, 5 * R, 1)
            frows = bitpack.pack_rows(dense.T)
            for p in range(2):
                feats = modules[p][modules[p] >= 0]
                z = oracle.score_set(frows, S, feats, w[p])[0]
                assert abs(res["z_obs"][p] - z) <= 1e-9 * np.abs(w[p]).sum()
                want_sc = [oracle.score_set(m, S, feats, w[p])[0] for m in nulls]
                assert np.allclose(res["null_scores"][p], want_sc, rtol=0, atol=1e-9 * np.abs(w[p]).sum())
                assert res["n_ge"][p] == int((res["null_scores"][p] >= res["z_obs"][p]).sum())


# ---------------------------------------------------------------- experimental work-queue kernel (SDX_TRADE_PERSIST=1)
def test_work_queue_kernel_is_bit_identical(monkeypatch):
    """k_trade_sm (one persistent CTA per SM, null slots in shared memory, warps pull rounds from a ring queue) must give
    the same nulls, applied-trade counts, chains, burn-in pool states and scores as the oracle."""
    monkeypatch.setenv("SDX_TRADE_PERSIST", "1")
    cases, _ = curveball_cases()
    rng = np.random.default_rng(8)
    mats = {"c0_shape": cases["c0_shape"]["dense"],
            "small": (rng.random((40, 17)) < 0.3).astype(np.uint8),
            "wide_rows": (rng.random((1500, 30)) < 0.1).astype(np.uint8)}
    with Engine(0) as e2:
        for name, dense in mats.items():
            S, F = dense.shape
            upload_dense(e2, dense)
            obs = rows_of(dense)
            R = obs.shape[0]
            for chain_len, null0, n, T in ((1, 0, 5, 5 * R), (3, 1, 7, 5 * R), (4, 0, 8, 0), (2, 3, 300, 7)):
                want, want_ap = oracle.curveball_nulls(obs, 11, null0, n, T, chain_len)
                got, got_ap = e2.curveball(11, null0, n, T, chain_len)
                assert np.array_equal(got_ap, want_ap), name
                assert np.array_equal(got, want), name
            e2.set_burn_in(3, 7)
            want, _ = oracle.curveball_nulls(obs, 5, 2, 20, 5 * R, 4, burn=3, pool_size=7, L=max(S, F))
            got, _ = e2.curveball(5, 2, 20, -1, 4, want_applied=False)
            assert np.array_equal(got, want), name
            e2.set_burn_in(0)
            w = synth.weights(S, 9, seed=S)
            e2.upload_weights(w)
            order = np.argsort(-dense.sum(0), kind="stable")
            modules = np.stack([np.sort(order[i:i + 3]) for i in range(9)]).astype(np.int32)
            res = e2.null_pvalue(3, 0, 40, modules, want_scores=True)
            nulls, _ = oracle.curveball_nulls(obs, 3, 0, 40, 5 * R, 1)
            frows = bitpack.pack_rows(dense.T)
            for p in range(9):
                ras = F >= S
                want_sc = [oracle.score_set(bitpack.pack_rows(bitpack.unpack_rows(m, max(S, F)).T) if ras else m, S, modules[p], w[p])[0] for m in nulls]
                assert np.allclose(res["null_scores"][p], want_sc, rtol=0, atol=1e-9 * np.abs(w[p]).sum()), name
                assert res["n_ge"][p] == int((res["null_scores"][p] >= res["z_obs"][p]).sum())


@pytest.mark.parametrize("cap", ["8", "32", "5"])
def test_capacity_level_schedule_gives_the_same_nulls(monkeypatch, cap):
    """SDX_PLAN_CAP: a first-fit level schedule with at most `cap` trades per level is another valid order of the same
    trades, so nulls and applied-trade counts must not change (including when the schedule overflows into the
    sequential fallback: cap = 5 needs more levels than dcap allows on the C0 shape)."""
    monkeypatch.setenv("SDX_PLAN_CAP", cap)
    cases, _ = curveball_cases()
    with Engine(0) as e2:
        for name in ("c0_shape", "square", "wide_small"):
            dense = cases[name]["dense"]
            upload_dense(e2, dense)
            obs = rows_of(dense)
            R = obs.shape[0]
            want, want_ap = oracle.curveball_nulls(obs, 21, 0, 6, 5 * R, 3)
            got, got_ap = e2.curveball(21, 0, 6, -1, 3)
            assert np.array_equal(got_ap, want_ap) and np.array_equal(got, want), name


# ---------------------------------------------------------------- individual contributions
def test_contributions_match_the_reference_arithmetic(eng):
    """utils/compute_individual_contribution_SELECT.py:69-85 restated with Python sets, against the GPU sums."""
    from superdendrix_b200 import contribution
    dense, _, _ = synth.feature_matrix(300, 50, seed=6)
    S, F = dense.shape
    w = synth.weights(S, 2, seed=8)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    rng = np.random.default_rng(1)
    pairs = np.stack([rng.choice(F, 2, replace=False) for _ in range(40)]).astype(np.int32)
    prof = rng.integers(0, 2, 40).astype(np.int32)
    m1, m2 = contribution.pair_contributions(eng, pairs, prof)
    for i in range(40):
        cl1 = set(np.flatnonzero(dense[:, pairs[i, 0]])); cl2 = set(np.flatnonzero(dense[:, pairs[i, 1]]))
        total = sum(w[prof[i]][s] for s in cl1 | cl2)
        assert abs(m1[i] - sum(w[prof[i]][s] for s in cl1) / total) <= 1e-9 * max(1.0, abs(m1[i]))
        assert abs(m2[i] - sum(w[prof[i]][s] for s in cl2) / total) <= 1e-9 * max(1.0, abs(m2[i]))
    mods = np.array([[0, 1, 2], [3, 4, -1]], dtype=np.int32)
    loo = contribution.leave_one_out(eng, mods, np.array([0, 1], dtype=np.int32))
    rows = bitpack.pack_rows(dense.T)
    for j, (m, p) in enumerate(zip(mods, (0, 1))):
        feats = [int(f) for f in m if f >= 0]
        full = oracle.score_set(rows, S, feats, w[p])[0]
        for q, f in enumerate(feats):
            rest = [g for g in feats if g != f]
            assert abs(loo[j, q] - (full - oracle.score_set(rows, S, rest, w[p])[0])) <= 1e-9 * np.abs(w[p]).sum()
    assert np.isnan(loo[1, 2])


# ---------------------------------------------------------------- BASELINE.json sizes: size-independent properties
def test_c1_full_size_counts_are_additive_and_match_the_oracle_within_mc_error(eng):
    """configs[1]: 100 000 nulls of the BRAF-shape matrix.  Shards add up exactly (what multi-GPU relies on); the
    p-value agrees with the oracle's estimate from an independent set of nulls within binomial error."""
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    w = synth.weights(S, 1)
    module = synth.pick_module(dense, 3)[None, :]
    upload_dense(eng, dense)
    eng.upload_weights(w)
    n = 100_000
    whole = eng.null_pvalue(1764, 0, n, module, want_scores=True)
    parts = [eng.null_pvalue(1764, lo, 12_500, module, z_obs=whole["z_obs"])["n_ge"][0] for lo in range(0, n, 12_500)]
    assert sum(parts) == whole["n_ge"][0] == int((whole["null_scores"][0] >= whole["z_obs"][0]).sum())
    obs = rows_of(dense)
    m = 1500
    nulls, applied = oracle.curveball_nulls(obs, 99, 5_000_000, m, 5 * obs.shape[0], 1)
    sc = np.array([oracle.score_set(x, S, module[0], w[0])[0] for x in nulls])
    p_gpu, p_cpu = whole["n_ge"][0] / n, float((sc >= whole["z_obs"][0]).mean())
    sd = np.sqrt(p_cpu * (1 - p_cpu) / m + p_gpu * (1 - p_gpu) / n)
    assert abs(p_gpu - p_cpu) < 4.5 * sd + 1e-9
    # the null score distribution itself: mean and spread agree
    assert abs(whole["null_scores"][0].mean() - sc.mean()) < 4.5 * sc.std() / np.sqrt(m)
    assert (applied > 0.9 * 5 * obs.shape[0]).all()


def test_c3_shape_global_memory_variant(eng):
    """configs[3] shape: 1 000 samples x 5 000 features (rows = samples, 157-word rows, matrix in global memory / L2,
    5 000 trades per null), several profiles scored per null."""
    dense, _, _ = synth.feature_matrix(1000, 5000)
    S, F = dense.shape
    P = 6
    w = synth.weights(S, P, seed=5)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    R, L, wpr, ras = eng.trade_shape()
    assert ras and R == S and L == F
    obs = rows_of(dense)
    got, got_ap = eng.curveball(7, 10, 3, chain_len=2)
    want, want_ap = oracle.curveball_nulls(obs, 7, 10, 3, 5 * R, 2)
    assert np.array_equal(got_ap, want_ap) and np.array_equal(got, want)
    assert np.array_equal(bitpack.popcount_rows(got[2]), bitpack.popcount_rows(obs))
    assert np.array_equal(bitpack.unpack_rows(got[2], L).sum(0), dense.sum(0))
    order = np.argsort(-dense.sum(0), kind="stable")[:30]
    rng = np.random.default_rng(0)
    mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(P)]), axis=1).astype(np.int32)
    res = eng.null_pvalue(7, 0, 48, mods, want_scores=True)
    a = eng.null_pvalue(7, 0, 24, mods, z_obs=res["z_obs"])["n_ge"]
    b = eng.null_pvalue(7, 24, 24, mods, z_obs=res["z_obs"])["n_ge"]
    assert np.array_equal(a + b, res["n_ge"])
    nulls, _ = oracle.curveball_nulls(obs, 7, 0, 2, 5 * R, 1)
    for i in range(2):
        fr = bitpack.pack_rows(bitpack.unpack_rows(nulls[i], L).T)
        for p in range(P):
            assert abs(res["null_scores"][p, i] - oracle.score_set(fr, S, mods[p], w[p])[0]) <= 1e-9 * np.abs(w[p]).sum()
    # ties count (src/superdendrix.py:192 uses >=): a null that equals the observed matrix must score EXACTLY z_obs although the
    # nulls are scored from the column cache and the observed matrix from its sample rows
    same = eng.null_pvalue(7, 0, 5, mods, n_trades=0, want_scores=True)
    assert np.array_equal(same["null_scores"], np.repeat(same["z_obs"][:, None], 5, axis=1))
    assert np.array_equal(same["n_ge"], np.full(P, 5))


def test_c4_slice_batched_model_selection_statistics(eng):
    """configs[4] slice: many profiles over the C0 matrix, 10 000 conditional permutations per feature and a
    rank-sum per profile; a sample of jobs is checked against the oracle exactly."""
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    n_prof = 300
    w = synth.weights(S, n_prof, seed=21)
    upload_dense(eng, dense)
    eng.upload_weights(w)
    rng = np.random.default_rng(3)
    order = np.argsort(-dense.sum(0), kind="stable")[:40]
    mods = np.sort(np.stack([rng.choice(order, 3, replace=False) for _ in range(n_prof)]), axis=1).astype(np.int32)
    pj = np.arange(n_prof, dtype=np.int32)
    counts, realW = eng.condperm_batch(mods, 10_000, 1764, 0, pj)
    rs = eng.ranksum(mods, pj)
    rows = bitpack.pack_rows(dense.T)
    assert counts.shape == (n_prof, 3) and (counts >= 0).all() and (counts <= 10_000).all()
    for j in (0, 17, 299):
        oc, orw = oracle.condperm(rows, S, mods[j], w[j], np.arange(S), 1764, 0, 10_000)
        assert np.array_equal(counts[j], oc) and np.array_equal(realW[j], orw)
        union = dense[:, mods[j]].any(axis=1)
        zz, pp, s2, n1 = oracle.ranksum(w[j], np.arange(S), bitpack.pack_mask(np.nonzero(union)[0], S))
        assert (rs["twice_rank_sum"][j], rs["n1"][j]) == (s2, n1) and abs(rs["z"][j] - zz) <= 1e-12 * max(1, abs(zz))
    # random profiles carry no signal: rank-sum p-values are roughly uniform
    assert 0.3 < np.median(rs["p"]) < 0.7


# ---------------------------------------------------------------- the rare branch of Lemire's method
# (half, sparse layout?, null indices of seed 4242 whose single trade enters the rejection zone; found by scanning the
#  oracle: tests/tools/find_lemire_cases.py.  The last index of each list also RE-DRAWS.)
LEMIRE_CASES = [
    (400, "0", [18883, 55949]),          # n = 800: general pick path (rank mask in the row's shared-memory storage)
    (32, "0", [314870, 16882340]),       # n = 64: register pick path
    (16, "1", [21381922]),               # n = 32: list x list path of the sparse-layout kernel
]


@pytest.mark.parametrize("half,sparse,hits", LEMIRE_CASES)
def test_lemire_rejection_branch_matches_oracle(monkeypatch, half, sparse, hits):
    """A pick enters the rejection zone with probability m / 2^32, so ordinary parity cases never exercise the
    device's out-of-line branch (deal_below_slow: re-draw from the pick's own sub-stream with attempt + 1).
    Two disjoint rows of `half` bits make every null one trade of `half` picks; the null indices below are known
    (from the oracle) to enter the zone and to re-draw.  Windows around them must be bit-identical."""
    monkeypatch.setenv("SDX_TRADE_SPARSE", sparse)
    L = max(2 * half, 64)
    rows_dense = np.zeros((2, L), dtype=np.uint8)
    rows_dense[0, :half] = 1
    rows_dense[1, half:2 * half] = 1
    dense = rows_dense.T                                   # S = L samples x F = 2 features: the trade rows are the features
    with Engine(0) as e2:
        upload_dense(e2, dense)
        obs = rows_of(dense)
        assert obs.shape[0] == 2
        zone = redraw = 0
        for idx in hits:
            oracle.deal_stats(reset=True)
            want, want_ap = oracle.curveball_nulls(obs, 4242, idx - 40, 80, 1, 1)
            z, r = oracle.deal_stats()
            zone += z
            redraw += r
            got, got_ap = e2.curveball(4242, idx - 40, 80, n_trades=1, chain_len=1)
            assert np.array_equal(got_ap, want_ap)
            assert np.array_equal(got, want)
    assert zone >= len(hits) and redraw >= 1


def test_trade_pair_rejection_zone_is_covered(eng):
    """With R = 65 535 rows a pair draw enters Lemire's rejection zone with probability ~1.5e-5; 400 000 trades
    contain about a dozen such draws (sequential stream: a rejected draw shifts the following ones)."""
    R, T = 65535, 400_000
    a, b = eng.test_trade_pairs(77, 5, T, R)
    import ctypes as C
    lib = oracle.lib()
    x, y = C.c_uint32(0), C.c_uint32(0)
    for t in range(T):
        lib.sdxo_trade_pair(C.c_uint64(77), C.c_uint64(5), C.c_uint32(t), C.c_uint32(R), C.byref(x), C.byref(y))
        if x.value != a[t] or y.value != b[t]:
            raise AssertionError("pair %d differs: device (%d, %d) oracle (%d, %d)" % (t, a[t], b[t], x.value, y.value))
    assert (a != b).all() and a.max() < R and b.max() < R


# ---------------------------------------------------------------- burn-in pool
@pytest.mark.parametrize("sparse", ["0", "1"])
def test_burn_in_pool_matches_oracle(monkeypatch, sparse):
    """sdx_set_burn_in: chains start from pool states (observed + burn rounds on reserved stream ids); nulls, applied
    counts and scores equal the oracle's, the pool is rebuilt when the seed changes, shards add up."""
    monkeypatch.setenv("SDX_TRADE_SPARSE", sparse)
    cases, _ = curveball_cases()
    dense = cases["c0_shape"]["dense"]
    S, F = dense.shape
    obs = rows_of(dense)
    R = obs.shape[0]
    w = synth.weights(S, 2, seed=12)
    modules = np.array([[0, 1, 2], [5, 9, -1]], dtype=np.int32)
    with Engine(0) as e2:
        upload_dense(e2, dense)
        e2.upload_weights(w)
        e2.set_burn_in(6, 8)
        for seed, chain_len, null0, n in ((1764, 5, 3, 17), (99, 1, 0, 11), (1764, 4, 40, 9)):
            want, want_ap = oracle.curveball_nulls(obs, seed, null0, n, 5 * R, chain_len, burn=6, pool_size=8, L=max(S, F))
            got, got_ap = e2.curveball(seed, null0, n, -1, chain_len)
            assert np.array_equal(got_ap, want_ap) and np.array_equal(got, want), (seed, chain_len)
        res = e2.null_pvalue(7, 0, 24, modules, chain_len=4, want_scores=True)
        nulls, _ = oracle.curveball_nulls(obs, 7, 0, 24, 5 * R, 4, burn=6, pool_size=8, L=max(S, F))
        frows = bitpack.pack_rows(dense.T)
        for p in range(2):
            feats = modules[p][modules[p] >= 0]
            want_sc = [oracle.score_set(m, S, feats, w[p])[0] for m in nulls]
            assert np.allclose(res["null_scores"][p], want_sc, rtol=0, atol=1e-9 * np.abs(w[p]).sum())
        a = e2.null_pvalue(7, 0, 12, modules, chain_len=4, z_obs=res["z_obs"])["n_ge"]
        b = e2.null_pvalue(7, 12, 12, modules, chain_len=4, z_obs=res["z_obs"])["n_ge"]
        assert np.array_equal(a + b, res["n_ge"])
        # the re-optimised statistic draws its nulls from the same pool (small matrix: the oracle scan is brute force)
        small = (np.random.default_rng(3).random((120, 40)) < 0.12).astype(np.uint8)
        ws = synth.weights(120, 1, seed=4)
        upload_dense(e2, small)
        e2.upload_weights(ws)
        mk = e2.null_pvalue(7, 0, 6, np.array([[0, 1, 2]], dtype=np.int32), chain_len=3, stat=SDX_STAT_MAXK, want_scores=True)
        nulls6, _ = oracle.curveball_nulls(rows_of(small), 7, 0, 6, 5 * 40, 3, burn=6, pool_size=8, L=120)
        want_mk = [max(oracle.scan(m, 120, ws[0], 3, 1)[0][0], 0.0) for m in nulls6]
        assert np.allclose(mk["null_scores"][0], want_mk, rtol=0, atol=1e-9 * np.abs(ws[0]).sum())
        upload_dense(e2, dense)
        e2.set_burn_in(0)                                           # back to chains that start from the observed matrix
        want, _ = oracle.curveball_nulls(obs, 1764, 0, 4, 5 * R, 1)
        assert np.array_equal(e2.curveball(1764, 0, 4, -1, 1)[0], want)
        with pytest.raises(SdxError):
            e2.set_burn_in(1000, 4)


def test_burn_in_removes_the_memory_of_the_observed_matrix(eng):
    """C0: after 5*min(S,F) trades the densest feature still shares ~440 of its 469 carriers with the observed
    matrix; the stationary overlap is ~296 (reached after ~30 rounds of feature-row trades, after 2 burn-in rounds of
    sample-row trades: DESIGN.md 3b).  With the default burn-in the nulls are there."""
    dense, _, _ = synth.feature_matrix(770, 400)
    S, F = dense.shape
    upload_dense(eng, dense)
    R, L, wpr, ras = eng.trade_shape()
    g = int(np.argmax(dense.sum(0)))
    obs_row = dense[:, g].astype(np.int64)

    def overlap(burn, chain_len, n):
        eng.set_burn_in(burn, 64)
        rows, _ = eng.curveball(5, 0, n, -1, chain_len, want_applied=False)
        return float((bitpack.unpack_rows(rows[:, g, :], L).astype(np.int64) * obs_row).sum(1).mean())
    try:
        raw = overlap(0, 1, 200)
        mixed = overlap(defaults.BURN_ROUNDS, 8, 400)
        longer = overlap(32, 8, 400)
    finally:
        eng.set_burn_in(0)
    assert raw > mixed + 80                       # restart nulls remember the observed matrix
    assert abs(mixed - longer) < 6                # the default is enough: ten times the burn-in changes nothing


def test_burn_in_orientation_branches_match_oracle():
    """The four ways a pool is burnt in (oracle: burn_plan_init): on the trade rows or on their transpose,
    with uniform or with dense-favouring pairs.  Pool-started nulls are bit-identical to the oracle's in every branch,
    on k_trade (pair_group 1) and with blocks of 32 chains sharing their row pairs."""
    seen = set()
    with Engine(0) as e2:
        for name, (dense, alt, fav) in burn_cases().items():
            obs = rows_of(dense)
            R, L = obs.shape[0], max(dense.shape)
            plan = oracle.burn_plan(obs, L)
            assert (plan["alt"], plan["dense_rows"] > 0) == (alt, fav), name
            seen.add((alt, fav))
            upload_dense(e2, dense)
            for pg, burn, pool, chain_len, null0, n in ((1, 2, 5, 3, 1, 11), (32, 3, 64, 1, 32, 70)):
                e2.set_pair_group(pg)
                e2.set_burn_in(burn, pool)
                want, want_ap = oracle.curveball_nulls(obs, 41, null0, n, 5 * R, chain_len, burn=burn, pool_size=pool, pair_group=pg, L=L)
                got, got_ap = e2.curveball(41, null0, n, -1, chain_len)
                assert np.array_equal(got_ap, want_ap) and np.array_equal(got, want), (name, pg)
            e2.set_pair_group(1)
            e2.set_burn_in(0)
    assert seen == {(False, False), (False, True), (True, False), (True, True)}


@pytest.mark.parametrize("lanes", [False, True])
def test_burn_in_pool_fuzz(lanes):
    """tools/fuzz_burn.py: pool-started nulls vs the oracle over random shapes, densities and degenerate matrices (empty, full
    rows / columns, two rows, word-aligned sizes), pair_group 1 and 32, on k_trade and — small requests forced onto it — on the lane kernel."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ)
    if lanes:
        env["SDX_LANE_MIN_CHAINS"] = "1"
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_burn.py"), "24", "31" if lanes else "5"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout
