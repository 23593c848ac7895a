"""Pins the oracle (oracle/sdx_oracle.c + oracle/refport.py) against fixtures
made from the reference itself (tests/golden/make_golden.py).  CPU only."""
import json
import os
import platform
import random

import numpy as np
import pytest

import oracle
from oracle import refport
from superdendrix_b200 import bitpack, defaults, synth

from conftest import GOLDEN, curveball_cases, load_npz, rows_of


# ---------------------------------------------------------------- Philox KAT
def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        assert tuple(int(x) for x in oracle.philox4x32_10(ctr, key)) == want


def test_stream_layout():
    d = oracle.stream_draws(0x1234567890abcdef, 1, (5 << 32) | 7, 42, 9)
    key = (0x90abcdef, 0x12345678)
    for blk in range(3):
        want = oracle.philox4x32_10((blk, 42, 7, 5 | (1 << 30)), key)
        n = min(4, 9 - 4 * blk)
        assert np.array_equal(d[4 * blk:4 * blk + n], want[:n])


def test_below_exact_rejection():
    # m = 3: threshold (2^32 - 3) % 3 = 1 -> only lo == 0 is rejected
    v, used = oracle.below_from_draws([0, 0x55555556, 7], 3)
    assert used == 2 and v == (0x55555556 * 3) >> 32
    v, used = oracle.below_from_draws([0xffffffff, 1], 3)
    assert (v, used) == (2, 1)
    rng = np.random.default_rng(0)
    for m in (1, 2, 7, 384, 770, 1 << 20):
        xs = rng.integers(0, 1 << 32, size=200, dtype=np.uint64).astype(np.uint32)
        for x in xs:
            v, _ = oracle.below_from_draws([int(x), int(x) ^ 0x9e3779b9, 12345], m)
            assert 0 <= v < m


def test_trade_pair_uniform_and_distinct():
    R = 5
    counts = np.zeros((R, R), dtype=int)
    for t in range(20000):
        a, b = oracle.trade_pair(99, 3, t, R)
        assert a != b and a < R and b < R
        counts[a, b] += 1
    off = counts[~np.eye(R, dtype=bool)]
    assert abs(off - 1000).max() < 150      # 20 cells, sd ~ 31


# ---------------------------------------------------------------- Curveball
@pytest.mark.parametrize("name", ["tall_small", "wide_small", "square", "short_run", "c0_shape"])
def test_replay_reproduces_reference(name):
    cases, _ = curveball_cases()
    c = cases[name]
    rows = rows_of(c["dense"])
    margins_r = bitpack.popcount_rows(rows)
    n_nulls = c["a"].shape[0]
    skipped = 0
    for i in range(n_nulls):
        rows, applied = oracle.curveball_replay(rows, c["a"][i], c["b"][i], c["to_a"][i])
        assert np.array_equal(rows, c["out"][i]), "null %d differs from the reference output" % i
        assert np.array_equal(bitpack.popcount_rows(rows), margins_r)
        # skipped trades are exactly those with an all-zero mask
        assert applied == int((c["to_a"][i].any(axis=1)).sum())
        skipped += c["a"].shape[1] - applied
    L = max(c["dense"].shape)
    cols0 = bitpack.unpack_rows(rows_of(c["dense"]), L).sum(0)
    cols1 = bitpack.unpack_rows(rows, L).sum(0)
    assert np.array_equal(cols0, cols1)


@pytest.mark.parametrize("name", ["tall_small", "wide_small", "square", "short_run", "c0_shape"])
def test_refport_matches_reference_stream(name):
    cases, pyver = curveball_cases()
    if pyver != platform.python_version():
        pytest.skip("fixtures were generated on Python %s" % pyver)
    c = cases[name]
    dense = c["dense"].astype(np.float64)
    S, F = dense.shape
    n_trades = c["a"].shape[1] if name == "short_run" else -1
    random.seed(int(c["seed"]))
    lists = refport.presences(dense)
    for i in range(c["a"].shape[0]):
        sched = []
        out = refport.curveball(dense, lists, n_trades, recorder=lambda a, b, to_a: sched.append((a, b)))
        rows = out if F >= S else out.T
        assert np.array_equal(bitpack.pack_rows(rows), c["out"][i])
        assert [s[0] for s in sched] == c["a"][i].tolist()
        assert [s[1] for s in sched] == c["b"][i].tolist()


def test_canonical_curveball_margins_and_determinism():
    cases, _ = curveball_cases()
    obs = rows_of(cases["c0_shape"]["dense"])
    R, wpr = obs.shape
    nulls, applied = oracle.curveball_nulls(obs, seed=1764, null0=0, n=3, n_trades=5 * R, chain_len=1)
    again, _ = oracle.curveball_nulls(obs, seed=1764, null0=1, n=2, n_trades=5 * R, chain_len=1)
    assert np.array_equal(nulls[1:], again)              # keyed on the global null index
    L = 770
    for m in nulls:
        assert np.array_equal(bitpack.popcount_rows(m), bitpack.popcount_rows(obs))
        assert np.array_equal(bitpack.unpack_rows(m, L).sum(0), bitpack.unpack_rows(obs, L).sum(0))
        assert not np.array_equal(m, obs)
    assert (applied > 0.95 * 5 * R).all()
    # chained: null 1 continues from null 0; a shard starting at 1 replays the prefix
    ch, _ = oracle.curveball_nulls(obs, 1764, 0, 3, 5 * R, chain_len=3)
    ch1, _ = oracle.curveball_nulls(obs, 1764, 1, 2, 5 * R, chain_len=3)
    assert np.array_equal(ch[0], nulls[0]) and not np.array_equal(ch[1], nulls[1])
    assert np.array_equal(ch[1:], ch1)


def test_canonical_trade_is_uniform_like_the_reference():
    """On a 2-row matrix one trade re-deals the symmetric difference; every
    |a\\b|-subset must be equally likely (src/curveball.py:30-35)."""
    a = np.zeros(40, bool); b = np.zeros(40, bool)
    a[[0, 1, 2, 33]] = True      # a\b = {0,1,33}; common = {2}
    b[[2, 5, 34, 35, 36]] = True  # b\a = {5,34,35,36}
    obs = bitpack.pack_rows(np.stack([a, b]))
    from collections import Counter
    seen = Counter()
    N = 14000
    for n in range(N):
        out, ap = oracle.curveball_nulls(obs, seed=5, null0=n, n=1, n_trades=1)
        assert ap[0] == 1
        bits = tuple(np.nonzero(bitpack.unpack_rows(out[0, 0], 40))[0].tolist())
        assert 2 in bits and len(bits) == 4
        seen[bits] += 1
    assert len(seen) == 35                      # C(7,3)
    exp = N / 35
    chi2 = sum((v - exp) ** 2 / exp for v in seen.values())
    assert chi2 < 75                            # 34 dof, p ~ 1e-4


# ---------------------------------------------------------------- SuperW
def test_score_set_matches_reference_superw():
    z = load_npz("superw.npz")
    dense = z["dense"]
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    for i in range(len(z["W"])):
        feats = z["modules"][i]
        feats = feats[feats >= 0]
        w = z["weights"][z["kind"][i]]
        W, cov, ov, ac = oracle.score_set(rows, S, feats, w)
        assert abs(W - z["W"][i]) <= 1e-9 * max(1.0, np.abs(w).sum())
        assert (cov, ov, ac) == (z["cov"][i], z["overlap"][i], z["act_cov"][i])


def test_score_set_mask_equals_zero_weight_restriction():
    z = load_npz("superw.npz")
    dense = z["dense"]
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    w = z["weights"][0].copy()
    keep = np.arange(S) % 4 != 0
    w_m = np.where(keep, w, 0.0)
    mask = bitpack.pack_mask(np.nonzero(keep)[0], S)
    sub_rows = bitpack.pack_rows(dense[keep].T)
    for feats in ([3], [1, 7, 20], [0, 5]):
        Wm, cm, om, am = oracle.score_set(rows, S, feats, w_m, mask)
        Ws, cs, os_, as_ = oracle.score_set(sub_rows, int(keep.sum()), feats, w[keep])
        assert abs(Wm - Ws) < 1e-9 and (cm, om, am) == (cs, os_, as_)


# ---------------------------------------------------------------- scan
@pytest.mark.parametrize("name", ["a", "b", "unit"])
def test_scan_matches_bruteforce_over_reference_superw(name):
    z = load_npz("scan.npz")
    dense, w = z[name + "__dense"], z[name + "__w"]
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    oW, oF, oC, n = oracle.scan(rows, S, w, 3, topn=16)
    assert n == int(z[name + "__n"])
    assert np.allclose(oW, z[name + "__topW"], rtol=0, atol=1e-9 * np.abs(w).sum())
    if name == "unit":      # exact arithmetic -> ties resolved by enumeration order, sets must match
        assert np.array_equal(oF, z[name + "__topF"])
    else:
        assert np.array_equal(oF[0], z[name + "__topF"][0])
    for k in (1, 2, 3):
        bW, _, _, _ = oracle.scan(rows, S, w, k, topn=1)
        assert abs(bW[0] - z[name + "__best_le_k"][k - 1]) < 1e-9 * np.abs(w).sum()


def test_scan_optimum_equals_the_reference_ilp_solved_by_highs():
    """Pin G4: the integral optimum of the reference's ILP (src/superdendrix.py:294-313, objective :393-401,
    restated for HiGHS in oracle/milp.py because Gurobi is not installed) equals the exhaustive-scan optimum, for every
    cardinality bound, including instances where the empty module wins (z = 0)."""
    from oracle import milp
    for seed, S, F, dens in ((1, 90, 40, 0.12), (2, 120, 25, 0.3), (3, 60, 30, 0.05)):
        rng = np.random.default_rng(seed)
        dense = (rng.random((S, F)) < dens).astype(np.uint8)
        w = synth.weights(S, 1, seed=seed)[0]
        if seed == 3:
            w = -np.abs(w)                                     # no feature can score > 0: the ILP selects nothing
        rows = bitpack.pack_rows(dense.T)
        for k in (1, 2, 3):
            z, feats, optimal = milp.ilp_optimum(dense, w, k)
            assert optimal
            bW, bF, _, _ = oracle.scan(rows, S, w, k, topn=1)
            best = max(float(bW[0]), 0.0)                      # the scan enumerates non-empty modules; M = {} scores 0
            assert abs(z - best) <= 1e-7 * max(1.0, np.abs(w).sum()), (seed, k, z, best)
            if best > 1e-6:
                got = oracle.score_set(rows, S, np.array(feats, dtype=np.int32), w)[0]
                assert abs(got - z) <= 1e-7 * np.abs(w).sum()    # the ILP's module really has that SuperW


# ---------------------------------------------------------------- cond. permutation
def _condperm_inputs():
    z = load_npz("condperm_data.npz")
    with open(os.path.join(GOLDEN, "condperm.json")) as fh:
        meta = json.load(fh)
    return z["dense"], z["w"], meta


def test_condperm_refport_exact_and_oracle_within_mc_error():
    dense, w, meta = _condperm_inputs()
    S, F = dense.shape
    rows = bitpack.pack_rows(dense.T)
    samples = ["s%d" % i for i in range(S)]
    profile = {samples[i]: float(w[i]) for i in range(S)}
    e2c = {"f%d" % g: set(samples[i] for i in np.nonzero(dense[:, g])[0]) for g in range(F)}
    same_py = meta["python"] == platform.python_version()
    for c in meta["cases"]:
        P = c["P"]
        if same_py:
            random.seed(c["seed"])
            ev, cnt = refport.conditional_permutation(["f%d" % g for g in c["feats"]], e2c, profile, samples, P)
            assert (ev, cnt) == (c["worst"], c["worst_count"])
        counts, realW = oracle.condperm(rows, S, c["feats"], w, np.arange(S), seed=77, perm_id0=0, P=P)
        W, _, _, _ = oracle.score_set(rows, S, c["feats"], w)
        assert np.allclose(realW, W, atol=1e-9 * np.abs(w).sum())
        for got, want in zip(counts, c["per_feature"]):
            p = max(want, 1) / P
            sd = np.sqrt(2 * P * p * (1 - p)) + 1
            assert abs(got - want) < 4.5 * sd, (got, want)


# ---------------------------------------------------------------- rank-sum
def test_ranksum_matches_scipy():
    z = load_npz("ranksum.npz")
    S = z["profile"].shape[1]
    for i in range(len(z["z"])):
        ub = bitpack.pack_mask(np.nonzero(z["carrier"][i])[0], S)
        zz, p, s2, n1 = oracle.ranksum(z["profile"][i], np.arange(S), ub)
        assert n1 == int(z["carrier"][i].sum())
        assert abs(zz - z["z"][i]) < 1e-12 * max(1, abs(z["z"][i]))
        assert abs(p - z["p"][i]) <= 1e-12 + 1e-10 * z["p"][i]


# ---------------------------------------------------------------- loader
def test_refport_loader_matches_reference():
    with open(os.path.join(GOLDEN, "loader.json")) as fh:
        cases = json.load(fh)
    d = os.path.join(GOLDEN, "loader")
    for name, c in cases.items():
        kw = dict(c["kwargs"])
        args = dict(patients=kw.get("patients"), gene_file=os.path.join(d, kw["geneFile"]) if "geneFile" in kw else None,
                    mut_only=kw.get("mut_only", False), min_freq=kw.get("min_freq", 0),
                    max_freq=kw.get("max_freq", float("inf")))
        m, n, genes, patients, g2c, p2g = refport.load_mutation_data(os.path.join(d, "features.tsv"), **args)
        assert (m, n) == (c["m"], c["n"]), name
        assert sorted(genes) == c["genes"] and list(patients) == c["patients"]
        assert {g: sorted(v) for g, v in g2c.items()} == c["geneToCases"]
        assert {p: sorted(v) for p, v in p2g.items()} == c["patientToGenes"]


# ---------------------------------------------------------------- null distribution of the reference's own generator
def _autocorr_inflation(x, max_lag=20):
    """Variance inflation of the mean of a (weakly) autocorrelated series: 1 + 2 * sum of positive autocorrelations."""
    x = np.asarray(x, dtype=np.float64) - np.mean(x)
    v = float(np.dot(x, x))
    rho = [float(np.dot(x[:-k], x[k:])) / v for k in range(1, max_lag + 1)]
    return 1.0 + 2.0 * sum(r for r in rho if r > 0)


def _null_sample(obs, S, F, module, w, n, **kw):
    nulls, _ = oracle.curveball_nulls(obs, 31337, 0, n, 5 * obs.shape[0], L=max(S, F), **kw)
    ras = F >= S
    L = max(S, F)
    sc = np.array([oracle.score_set(bitpack.pack_rows(bitpack.unpack_rows(m, L).T) if ras else m, S, module, w)[0] for m in nulls])
    occ = np.mean([bitpack.unpack_rows(m, L) if ras else bitpack.unpack_rows(m, L).T for m in nulls], axis=0)      # S x F
    return sc, occ


def test_canonical_nulls_follow_the_reference_null_distribution():
    """north star (3): with burn-in, the canonical (Philox / Floyd) nulls and the reference's own chain (CPython
    random; null i has seen i+1 rounds) sample the same null distribution: mean, spread and exceedance probabilities
    of W and the per-cell occupancy agree within Monte-Carlo error (the reference series is one Markov chain, so its
    effective sample size is deflated by its autocorrelation)."""
    z = load_npz("nulldist.npz")
    dense, w, module, ref = z["dense"], z["w"], z["module"], z["scores"]
    S, F = dense.shape
    obs = rows_of(dense)
    n = 1500
    sc, occ = _null_sample(obs, S, F, module, w, n, chain_len=8, burn=24, pool_size=64)
    infl = _autocorr_inflation(ref)
    se = np.sqrt(ref.var() * infl / len(ref) + 2 * sc.var() / n)
    assert abs(sc.mean() - ref.mean()) < 4.5 * se
    assert 0.85 < sc.std() / ref.std() < 1.18
    for q in (0.1, 0.5, 0.9):                                  # exceedance probabilities at reference quantiles
        thr = np.quantile(ref, q)
        p_ref, p_new = (ref >= thr).mean(), (sc >= thr).mean()
        sd = np.sqrt(p_ref * (1 - p_ref) * infl / len(ref) + 2 * p_new * (1 - p_new) / n)
        assert abs(p_ref - p_new) < 4.5 * sd + 0.01
    diff = np.abs(occ - z["cell_mean"])
    assert diff.max() < 0.12 and diff.mean() < 0.008


def test_without_burn_in_restart_nulls_are_not_mixed():
    """Why the burn-in exists: 5*min(S,F) trades from the observed matrix leave the dense rows close to where they
    started (the reference only gets away with it because its nulls form one long chain)."""
    z = load_npz("nulldist.npz")
    dense, w, module = z["dense"], z["w"], z["module"]
    S, F = dense.shape
    _, occ = _null_sample(rows_of(dense), S, F, module, w, 400, chain_len=1)
    diff = np.abs(occ - z["cell_mean"])
    assert diff.max() > 0.3                 # e.g. carriers of the densest feature keep it with p ~0.8 instead of ~0.36


def test_burn_plan_branches_keep_the_margins():
    """the pool is burnt in on the trade rows or on their transpose, with uniform or dense-favouring pairs;
    every branch returns states in the trade orientation with both margins of the observed matrix."""
    from conftest import burn_cases
    for name, (dense, alt, fav) in burn_cases().items():
        obs = rows_of(dense)
        L = max(dense.shape)
        plan = oracle.burn_plan(obs, L)
        assert (plan["alt"], plan["dense_rows"] > 0) == (alt, fav), name
        assert plan["rows"] == (L if alt else obs.shape[0]) and plan["trades_per_round"] == 5 * plan["rows"]
        ob = bitpack.unpack_rows(obs, L)
        for j in range(3):
            st = bitpack.unpack_rows(oracle.pool_state(obs, L, 5, 2, j), L)
            assert np.array_equal(st.sum(0), ob.sum(0)) and np.array_equal(st.sum(1), ob.sum(1)), name
            assert not np.array_equal(st, ob)
        # state j is a pure function of (matrix, seed, burn, j) and is what the chains start from
        nulls, _ = oracle.curveball_nulls(obs, 5, 0, 1, 0, 1, burn=2, pool_size=3, L=L)          # 0 trades: the null is its start state
        assert np.array_equal(nulls[0], oracle.pool_state(obs, L, 5, 2, 0))


def test_burn_in_on_the_balanced_orientation_is_stationary_after_two_rounds():
    """C0 (770 x 384): feature rows reach 469 carriers (44 x the mean row), sample rows 13 features (2.5 x), so the
    burn-in trades sample rows.  Ones shared with the observed matrix, one ordinary round after the pool state: 567 after
    one burn-in round, 558 after two, three and eight (stationary; 30 rounds of feature-row trades get there)."""
    dense, _, _ = synth.feature_matrix(770, 400, seed=20)
    S, F = dense.shape
    obs = rows_of(dense)
    assert oracle.burn_plan(obs, S) == {"alt": True, "rows": S, "trades_per_round": 5 * S, "dense_rows": 0}
    n = 500

    def shared(burn):
        nulls, _ = oracle.curveball_nulls(obs, 77, 0, n, 5 * F, 1, burn=burn, pool_size=n, L=S)
        ov = bitpack.popcount_rows(nulls & obs[None]).sum(1)
        return ov.mean(), ov.std() / np.sqrt(n)
    m2, s2 = shared(2)
    m3, s3 = shared(defaults.BURN_ROUNDS)
    m8, s8 = shared(8)
    assert abs(m2 - m8) < 4.5 * np.hypot(s2, s8) and abs(m3 - m8) < 4.5 * np.hypot(s3, s8)
    raw, _ = oracle.curveball_nulls(obs, 77, 0, 100, 5 * F, 1)
    assert bitpack.popcount_rows(raw & obs[None]).sum(1).mean() > m8 + 500        # restarts from the observed matrix: ~1 285 shared ones
