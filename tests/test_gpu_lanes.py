"""Parity of k_trade_lanes (one thread per null, csrc/sdx_lanes.cu) through the C ABI: bit for bit against the oracle
(oracle.curveball_nulls(pair_group=32), which restates src/curveball.py:17-40 with the canonical streams), against
k_trade running the same pair_group, and the pool built in parts against the pool built in one piece.
Needs a B200: run with -m gpu."""
import os

import numpy as np
import pytest

import oracle
from superdendrix_b200 import bitpack, defaults, synth
from superdendrix_b200.engine import Engine

from conftest import rows_of

pytestmark = pytest.mark.gpu


def _matrix(S, F, dens=None, seed=0):
    if dens is None:
        dense, _, _ = synth.feature_matrix(S, F)
    else:
        dense = (np.random.default_rng(S * 1000 + F + seed).random((S, F)) < dens).astype(np.uint8)
    return dense


CASES = {
    # name: (S, F, density or None for the DepMap-like synth, n, chain_len, null0, burn, pool)
    "lists_and_bits": (300, 120, None, 64, 1, 0, 0, 7),
    "chains_of_3_offset": (300, 120, None, 70, 3, 5, 0, 7),
    "all_lists": (200, 60, 0.03, 96, 2, 64, 0, 7),
    "all_bits": (70, 40, 0.5, 40, 1, 0, 0, 7),
    "pool_compact": (130, 50, 0.25, 64, 2, 0, 2, 7),            # pool size not a multiple of 32: compact start states
    "pool_groups": (130, 50, 0.25, 96, 1, 32, 2, 64),           # whole groups of the pool are block-copied
    "rows_are_samples": (40, 90, 0.15, 64, 1, 0, 0, 7),
    "c0_shape": (770, 400, None, 64, 2, 0, 0, 7),
    "full_and_empty_rows": (150, 48, 0.1, 64, 1, 0, 0, 7),
}


@pytest.mark.parametrize("name", list(CASES))
def test_lane_kernel_matches_oracle_bit_for_bit(name, monkeypatch):
    S, F, dens, n, chain_len, null0, burn, pool = CASES[name]
    monkeypatch.setenv("SDX_LANE_MIN_CHAINS", "1")
    dense = _matrix(S, F, dens)
    if name == "full_and_empty_rows":
        dense[:, 0] = 1
        dense[:, 1] = 0
        dense[0, :] = 1
    obs = rows_of(dense)
    R = obs.shape[0]
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.set_pair_group(32)
        eng.set_burn_in(burn, pool)
        got, ap = eng.curveball(1764, null0, n, chain_len=chain_len)
        assert eng.timing()["plan_ms"] == 0.0, "the request did not run on the lane kernel"
    want, wap = oracle.curveball_nulls(obs, 1764, null0, n, 5 * R, chain_len, burn=burn, pool_size=pool, pair_group=32, L=max(dense.shape))
    assert np.array_equal(ap, wap)
    assert np.array_equal(got, want)
    L = max(dense.shape)
    ob = bitpack.unpack_rows(obs, L)
    for nul in got[:4]:                                          # margins (G1)
        bits = bitpack.unpack_rows(nul, L)
        assert np.array_equal(bits.sum(0), ob.sum(0)) and np.array_equal(bits.sum(1), ob.sum(1))


def test_lane_kernel_fused_scores_and_counts_match_oracle(monkeypatch):
    monkeypatch.setenv("SDX_LANE_MIN_CHAINS", "1")
    dense = _matrix(300, 120)
    S, F = dense.shape
    w = synth.weights(S, 3, seed=4)
    mask = np.ones((3, S), dtype=bool)
    mask[1, ::7] = False                                         # a profile that covers a subset of the samples
    order = np.argsort(-dense.sum(0), kind="stable")
    modules = np.array([[order[1], order[2], order[3]], [order[0], order[40], -1], [order[60], order[5], order[80]]], dtype=np.int32)
    n = 160
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w, mask)
        eng.set_pair_group(32)
        eng.set_burn_in(2, 64)
        res = eng.null_pvalue(11, 32, n, modules, chain_len=1, want_scores=True)
        assert eng.timing()["plan_ms"] == 0.0
    obs = rows_of(dense)
    nulls, _ = oracle.curveball_nulls(obs, 11, 32, n, 5 * obs.shape[0], 1, burn=2, pool_size=64, pair_group=32, L=max(dense.shape))
    frows = bitpack.pack_rows(dense.T)
    for p in range(3):
        wm = np.where(mask[p], w[p], 0.0)
        mod = modules[p][modules[p] >= 0]
        z = oracle.score_set(frows, S, mod, wm)[0]
        assert abs(res["z_obs"][p] - z) <= 1e-9 * np.abs(wm).sum()
        sc = np.array([oracle.score_set(m, S, mod, wm)[0] for m in nulls])
        assert np.allclose(res["null_scores"][p], sc, rtol=0, atol=1e-9 * np.abs(wm).sum())
        assert int(res["n_ge"][p]) == int((res["null_scores"][p] >= res["z_obs"][p]).sum())


def test_both_kernels_generate_the_same_nulls_for_pair_group_32(monkeypatch):
    dense = _matrix(300, 120)
    out = {}
    for lanes in ("1", "0"):
        monkeypatch.setenv("SDX_LANES", lanes)
        monkeypatch.setenv("SDX_LANE_MIN_CHAINS", "1")
        with Engine(0) as eng:
            eng.upload_dense(dense)
            eng.set_pair_group(32)
            eng.set_burn_in(2, 32)
            out[lanes] = eng.curveball(5, 0, 96, chain_len=1)
            assert (eng.timing()["plan_ms"] == 0.0) == (lanes == "1")
    assert np.array_equal(out["1"][0], out["0"][0]) and np.array_equal(out["1"][1], out["0"][1])


def test_a_null_equal_to_the_observed_matrix_scores_exactly_z_obs(monkeypatch):
    """zero trades: every null is its start state; with burn = 0 that is the observed matrix, and ties count (src/superdendrix.py:192)"""
    monkeypatch.setenv("SDX_LANE_MIN_CHAINS", "1")
    dense = _matrix(300, 120)
    w = synth.weights(dense.shape[0], 1, seed=9)
    order = np.argsort(-dense.sum(0), kind="stable")
    modules = np.array([[order[0], order[30], order[90]]], dtype=np.int32)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        eng.set_pair_group(32)
        res = eng.null_pvalue(1, 0, 64, modules, n_trades=0, chain_len=1, want_scores=True)
        assert eng.timing()["plan_ms"] == 0.0
    assert np.all(res["null_scores"][0] == res["z_obs"][0]) and int(res["n_ge"][0]) == 64


def test_pool_built_in_parts_equals_the_pool_built_in_one_piece():
    """sdx_pool_begin / sdx_pool_build / sdx_pool_commit (the multi-GPU pool: every rank burns a slice) give the nulls of
    the pool the library builds on first use"""
    dense = _matrix(200, 60, 0.1)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(3, 64)
        whole, _ = eng.curveball(21, 0, 128, chain_len=1)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(3, 64)
        ptr, state_bytes = eng.pool_begin(21)
        assert ptr and state_bytes > 0
        eng.pool_build(40, 24)
        eng.pool_build(0, 40)
        eng.pool_commit()
        parts, _ = eng.curveball(21, 0, 128, chain_len=1)
    assert np.array_equal(whole, parts)


def test_c1_defaults_p_value_is_independent_of_the_kernel_and_of_batching(monkeypatch):
    """C1 workload at reduced size with the shipped defaults: the counts of one call equal the sum over two calls that
    split the null range at a block boundary (what two ranks do)"""
    from superdendrix_b200 import dist
    dense, _, _ = synth.feature_matrix(770, 400, seed=20)
    w = synth.weights(dense.shape[0], 1, seed=3)
    module = synth.pick_module(dense, 3)[None, :].astype(np.int32)
    n = 4096
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        one = dist.null_pvalue(eng, 1764, n, module)
        lo0, hi0 = dist.shard_chains(n, defaults.CHAIN_LEN, 0, 2, defaults.PAIR_GROUP)
        lo1, hi1 = dist.shard_chains(n, defaults.CHAIN_LEN, 1, 2, defaults.PAIR_GROUP)
        assert (lo0, hi1) == (0, n) and hi0 == lo1 and hi0 % (32 * defaults.CHAIN_LEN) == 0
        a = eng.null_pvalue(1764, lo0, hi0 - lo0, module, chain_len=defaults.CHAIN_LEN, z_obs=one["z_obs"])
        b = eng.null_pvalue(1764, lo1, hi1 - lo1, module, chain_len=defaults.CHAIN_LEN, z_obs=one["z_obs"])
    assert int(one["n_ge"][0]) == int(a["n_ge"][0]) + int(b["n_ge"][0])
    assert 0.3 < one["p"][0] < 0.95


def _autocorr_inflation(x, max_lag=200):
    """integrated autocorrelation time 1 + 2 sum rho_k of a chain (truncated at the first non-positive estimate)"""
    x = np.asarray(x, dtype=np.float64) - np.mean(x)
    v = float(np.dot(x, x)) / len(x)
    tau = 1.0
    for k in range(1, min(max_lag, len(x) // 4)):
        r = float(np.dot(x[:-k], x[k:])) / (len(x) - k) / v
        if r <= 0:
            break
        tau += 2 * r
    return tau


def test_c0_shipped_defaults_sample_the_reference_chains_null_distribution():
    """North star (3) on the bench workload: tests/golden/nulldist_c0.npz holds SuperW of the bench module on 4 000
    CONSECUTIVE nulls of the reference's own chain (src/generate_null_matrices.py:71-74 driving src/curveball.py:17-40 with
    CPython's random; made by tests/golden/make_golden.py).  The shipped defaults — burn-in pool of 1 472 states, 3
    burn-in rounds on the sample rows, every null ONE ordinary round from a pool state, blocks of 32 nulls sharing their row pairs, the
    lane kernel — must sample the same stationary distribution: p-value, mean, spread and exceedance probabilities at three
    reference quantiles within Monte-Carlo error.  The reference series is one Markov chain: its error is inflated by its
    integrated autocorrelation time; ours by the nulls that share a pool state (68 per state at 100 000 nulls)."""
    from conftest import load_npz
    from superdendrix_b200 import dist
    z = load_npz("nulldist_c0.npz")
    ref_all, z_obs = z["scores"], float(z["z_obs"])
    ref = ref_all[200:]                                      # the chain's first nulls still remember the observed matrix
    dense, _, _ = synth.feature_matrix(770, 400, seed=20)
    w = synth.weights(dense.shape[0], 1, seed=3)
    module = synth.pick_module(dense, 3)
    assert np.array_equal(module, z["module"])
    n = 100_000
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
        res = eng.null_pvalue(1764, 0, n, module[None, :].astype(np.int32), chain_len=defaults.CHAIN_LEN, want_scores=True)
        assert eng.timing()["plan_ms"] == 0.0               # the lane kernel ran
    sc = res["null_scores"][0]
    assert abs(res["z_obs"][0] - z_obs) <= 1e-9 * np.abs(w).sum()
    tau = _autocorr_inflation(ref)
    # ours: nulls of one pool state are one round apart from it; treat each pool state's 68 nulls as one cluster
    per_state = sc[: (n // defaults.POOL_SIZE) * defaults.POOL_SIZE].reshape(-1, defaults.POOL_SIZE)        # null i starts from state i % pool
    cluster_means = per_state.mean(axis=0)
    se_ours_mean = cluster_means.std(ddof=1) / np.sqrt(defaults.POOL_SIZE)
    se_mean = np.sqrt(ref.var() * tau / len(ref) + se_ours_mean ** 2)
    assert abs(sc.mean() - ref.mean()) < 4.5 * se_mean
    assert 0.9 < sc.std() / ref.std() < 1.1
    for thr in [z_obs] + [float(np.quantile(ref, q)) for q in (0.1, 0.5, 0.9)]:
        p_ref, p_new = float((ref >= thr).mean()), float((sc >= thr).mean())
        cl = (per_state >= thr).mean(axis=0)
        se = np.sqrt(p_ref * (1 - p_ref) * tau / len(ref) + cl.var(ddof=1) / defaults.POOL_SIZE)
        assert abs(p_ref - p_new) < 4.5 * se, (thr, p_ref, p_new, se)
    # the chain's own relaxation is visible in the fixture: its first nulls are closer to the observed score than to the stationary mean
    assert abs(ref_all[:3].mean() - z_obs) < abs(ref_all[:3].mean() - ref.mean())


def test_mixing_check_flags_a_pool_that_remembers_the_observed_matrix():
    """superdendrix_b200.mixing: from stationary pool states the overlap with the observed matrix does not drift along a chain;
    without burn-in (chains start from the observed matrix itself) it falls round after round"""
    from superdendrix_b200 import mixing
    dense, _, _ = synth.feature_matrix(770, 400, seed=20)
    frows = bitpack.pack_rows(dense.T)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.set_pair_group(defaults.PAIR_GROUP)
        eng.set_burn_in(0, 0)
        raw = mixing.overlap_drift(eng, frows, dense.shape[0])
        eng.set_burn_in(defaults.BURN_ROUNDS, defaults.POOL_SIZE)
        mixed = mixing.overlap_drift(eng, frows, dense.shape[0])
        assert mixing.check(eng, frows, dense.shape[0], log=lambda m: None)
    assert raw["z"] > 20 and raw["first"] > raw["last"] + 300
    assert abs(mixed["z"]) < 4 and abs(mixed["first"] - 558) < 15
