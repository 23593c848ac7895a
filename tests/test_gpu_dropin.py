"""The reference-named Python entry points (drop-in layer) running on the GPU, checked against the oracle,
the golden fixtures made from the reference and the reference's file formats.  Needs a B200: -m gpu."""
import json
import os
import random

import numpy as np
import pytest

import oracle
from oracle import refport
from superdendrix_b200 import bitpack, curveball, dist, generate_null_matrices as gnm, i_o, synth, defaults
from superdendrix_b200 import superdendrix as sd
from superdendrix_b200.engine import Engine

from conftest import GOLDEN, load_npz, rows_of

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- src/curveball.py
@pytest.mark.parametrize("shape", [(40, 17), (17, 40), (25, 25)])
def test_curveball_shim_keeps_reference_semantics(shape):
    rng = np.random.default_rng(shape[0])
    mat = (rng.random(shape) < 0.3).astype(np.float64)
    hp = curveball.find_presences(mat)
    assert [list(map(int, r)) for r in hp] == [list(map(int, r)) for r in refport.presences(mat)]
    curveball.seed(1764)
    obs = rows_of(mat.astype(np.uint8))
    want, _ = oracle.curveball_nulls(obs, 1764, 0, 3, 5 * min(shape), chain_len=3)     # one chain, nulls 0..2
    L = max(shape)
    for i in range(3):
        out = curveball.curve_ball(mat, hp)
        assert out.dtype == np.int8 and out.shape == mat.shape
        assert np.array_equal(out.sum(0), mat.sum(0)) and np.array_equal(out.sum(1), mat.sum(1))
        state = out if shape[1] >= shape[0] else out.T
        assert np.array_equal(bitpack.pack_rows(state), want[i])
        # r_hp holds the new state (the reference mutates it in place, src/curveball.py:34-35)
        assert [sorted(map(int, r)) for r in hp] == [list(np.flatnonzero(state[r])) for r in range(len(hp))]
    out = curveball.curve_ball(mat, hp, num_iterations=0)
    assert np.array_equal(out if shape[1] >= shape[0] else out.T, bitpack.unpack_rows(want[2], L))


# ---------------------------------------------------------------- SuperW / conditional permutation under the reference names
def test_superw_shim_matches_reference_values():
    z = load_npz("superw.npz")
    dense = z["dense"]
    S, F = dense.shape
    samples = ["s%d" % i for i in range(S)]
    for i in range(0, len(z["W"]), 7):
        w = z["weights"][z["kind"][i]]
        profile = {samples[j]: float(w[j]) for j in range(S)}
        feats = z["modules"][i][z["modules"][i] >= 0]
        cases = [set(samples[j] for j in np.flatnonzero(dense[:, g])) for g in feats]
        assert abs(sd.SuperW(cases, profile, samples) - z["W"][i]) <= 1e-9 * max(1.0, np.abs(w).sum())
    assert sd.SuperW([], {"a": 1.0}, ["a"]) == 0.0


def test_conditional_permutation_shim_agrees_with_reference_decisions():
    zd = load_npz("condperm_data.npz")
    with open(os.path.join(GOLDEN, "condperm.json")) as fh:
        meta = json.load(fh)
    dense, w = zd["dense"], zd["w"]
    S, F = dense.shape
    samples = ["s%d" % i for i in range(S)]
    profile = {samples[i]: float(w[i]) for i in range(S)}
    e2c = {"f%d" % g: set(samples[i] for i in np.flatnonzero(dense[:, g])) for g in range(F)}
    sd.seed(11)
    for c in meta["cases"]:
        P = c["P"]
        ev, cnt = sd.conditional_permutation_test(["f%d" % g for g in c["feats"]], e2c, profile, samples, P)
        want_counts = c["per_feature"]
        top = max(want_counts)
        p = max(top, 1) / P
        assert abs(cnt - top) < 4.5 * (np.sqrt(2 * P * p * (1 - p)) + 1)
        srt = sorted(want_counts, reverse=True)
        if top == 0:
            assert ev is None or cnt <= 3
        elif len(srt) == 1 or srt[0] - srt[1] > 6 * np.sqrt(P * p * (1 - p)) + 3:      # decision only checked when unambiguous
            assert ev == c["worst"]


def test_optimize_model_and_mip_start_hook_against_the_reference_ilp():
    """optimize_model (src/superdendrix.py:267-357 return tuple) by exhaustive scan, and the warm-start hook of SURVEY 8f:
    certified optimum for k <= 3 (equals the reference's ILP solved by HiGHS), feasible lower bound for k > 3."""
    import types
    from oracle import milp
    dense, samples, features = synth.feature_matrix(300, 60, seed=9)
    S, F = dense.shape
    wv = synth.weights(S, 1, seed=4)[0]
    w = {samples[i]: float(wv[i]) for i in range(S)}
    e2s = {features[g]: set(samples[i] for i in np.flatnonzero(dense[:, g])) for g in range(F)}
    s2g = {samples[i]: set(features[g] for g in np.flatnonzero(dense[i])) for i in range(S)}
    for k in (1, 2, 3):
        z_ilp, feats_ilp, optimal = milp.ilp_optimum(dense, wv, k)
        assert optimal
        z, module, best_single, cov, act_cov, act_sample, x, y, m = sd.optimize_model(features, samples, w, e2s, s2g, types.SimpleNamespace(cardinality=k))
        assert abs(z - z_ilp) <= 1e-7 * np.abs(wv).sum() and (x, y, m) == (None, None, None)
        assert sorted(module) == sorted(features[g] for g in feats_ilp)
        assert cov == int(dense[:, feats_ilp].any(1).sum()) and act_sample == int((wv > 0).sum())
        assert act_cov == int((dense[:, feats_ilp].any(1) & (wv > 0)).sum())
        hook = sd.mip_start(features, samples, w, e2s, k)
        assert hook["certified"] and abs(hook["lower_bound"] - z_ilp) <= 1e-7 * np.abs(wv).sum()
        assert sum(hook["x_start"].values()) == len(module)
    z3 = sd.mip_start(features, samples, w, e2s, 3)["lower_bound"]
    for k in (4, 6):
        hook = sd.mip_start(features, samples, w, e2s, k)
        z_ilp, _, optimal = milp.ilp_optimum(dense, wv, k)
        assert optimal and not hook["certified"]
        assert z3 - 1e-9 <= hook["lower_bound"] <= z_ilp + 1e-7 * np.abs(wv).sum()      # feasible: never above the optimum
        idx = [features.index(g) for g in hook["module"]]
        rows = bitpack.pack_rows(dense.T)
        assert abs(oracle.score_set(rows, S, np.array(idx, dtype=np.int32), wv)[0] - hook["lower_bound"]) <= 1e-9 * np.abs(wv).sum()
        assert len(idx) <= k and sum(hook["x_start"].values()) == len(idx)


# ---------------------------------------------------------------- generator CLI and file formats
def test_generate_null_matrices_cli_writes_reference_format(tmp_path):
    gdir = os.path.join(GOLDEN, "generator")
    out = tmp_path / "nulls"
    out.mkdir()
    rc = gnm.main(["-m", os.path.join(gdir, "features.tsv"), "-p", "6", "-o", str(out) + "/", "-rs", "5", "--mode", "restart",
                   "--packed", str(tmp_path / "nulls.npz")])
    assert rc == 0
    m, n, genes, samples, g2c, p2g = i_o.load_mutation_data(os.path.join(gdir, "features.tsv"))
    dense = i_o.dense_matrix(genes, samples, p2g)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        R, L, wpr, ras = eng.trade_shape()
        rows, _ = eng.curveball(5, 0, 6, chain_len=1)
    pk = np.load(tmp_path / "nulls.npz")
    for i in range(6):
        path = out / ("%d.tsv" % i)
        assert open(path).readline() == "#Samples\tEvents\n"
        pm, pn, pgenes, psamples, pg2c, pp2g = i_o.load_mutation_data(str(path))
        assert psamples == samples                                    # every sample listed, file order kept
        nd = i_o.dense_matrix(genes, samples, pp2g)
        assert np.array_equal(nd.sum(0), dense.sum(0)) and np.array_equal(nd.sum(1), dense.sum(1))
        bits = bitpack.unpack_rows(rows[i], L)
        assert np.array_equal(nd, bits if ras else bits.T)
        assert np.array_equal(pk["feature_rows"][i], bitpack.pack_rows(nd.T))
        # events in a row follow the order of `genes`, as src/generate_null_matrices.py:78 does
        for line in open(path).read().splitlines()[1:4]:
            ev = line.split("\t")[1:]
            assert ev == [g for g in genes if g in set(ev)]
    # sharding by -pre: nulls 4..5 of a second process equal nulls 4..5 above
    out2 = tmp_path / "shard"
    out2.mkdir()
    gnm.main(["-m", os.path.join(gdir, "features.tsv"), "-p", "2", "-pre", "4", "-o", str(out2) + "/", "-rs", "5", "--mode", "restart"])
    for i in (4, 5):
        assert open(out2 / ("%d.tsv" % i)).read() == open(out / ("%d.tsv" % i)).read()
    # chained mode: one Markov chain, like the reference driver
    out3 = tmp_path / "chain"
    out3.mkdir()
    gnm.main(["-m", os.path.join(gdir, "features.tsv"), "-p", "3", "-o", str(out3) + "/", "-rs", "5"])
    want, _ = oracle.curveball_nulls(rows_of(dense), 5, 0, 3, 5 * R, chain_len=3)
    pp2g = i_o.load_mutation_data(str(out3 / "2.tsv"))[5]
    nd = i_o.dense_matrix(genes, samples, pp2g)
    assert np.array_equal(bitpack.pack_rows(nd if ras else nd.T), want[2])
    # parallel mode: short chains from the burnt-in pool
    out4 = tmp_path / "par"
    out4.mkdir()
    gnm.main(["-m", os.path.join(gdir, "features.tsv"), "-p", "5", "-o", str(out4) + "/", "-rs", "5", "--mode", "parallel", "--burn-in", "4"])
    want, _ = oracle.curveball_nulls(rows_of(dense), 5, 0, 5, 5 * R, defaults.CHAIN_LEN, burn=4, pool_size=defaults.POOL_SIZE,
                                     pair_group=defaults.PAIR_GROUP, L=max(dense.shape))
    nd = i_o.dense_matrix(genes, samples, i_o.load_mutation_data(str(out4 / "4.tsv"))[5])
    assert np.array_equal(bitpack.pack_rows(nd if ras else nd.T), want[4])


# ---------------------------------------------------------------- the driver end to end
def _write_case(tmp_path, S=120, F=40, seed=3):
    dense, samples, features = synth.feature_matrix(S, F, seed=seed)
    w = synth.weights(S, 2, seed=seed + 1)
    w[0, 5] = np.nan
    mut = tmp_path / "features.tsv"
    synth.write_features_tsv(str(mut), dense, samples, features)
    prof = tmp_path / "profiles.tsv"
    synth.write_profile_tsv(str(prof), samples, w, ["GENE1", "GENE2"])
    return dense, samples, features, w, mut, prof


@pytest.mark.parametrize("stat,scope", [("fixed", "restricted"), ("max_k", "restricted"), ("fixed", "full"), ("max_k", "full")])
def test_driver_end_to_end_against_oracle(tmp_path, stat, scope):
    dense, samples, features, w, mut, prof = _write_case(tmp_path)
    out = tmp_path / "res.tsv"
    args = sd.get_parser().parse_args(["-m", str(mut), "-T", str(prof), "-Tc", "GENE1", "-k", "3", "-cp", "200", "-p", "64",
                                       "-o", str(out), "-d", "positive", "--null-stat", stat, "-rs", "9", "--null-scope", scope])
    module = sd.run(args)
    keep = ~np.isnan(w[0])
    kept = {s for s, k_ in zip(samples, keep) if k_}
    # the matrix as the driver sees it: profiled samples only; features without a carrier among them vanish
    # (src/i_o.py:57,62-67); feature order = the loader's
    _, _, names, smp, _, s2g = i_o.load_mutation_data(str(mut), kept)
    d = i_o.dense_matrix(names, smp, s2g)
    assert smp == [s for s in samples if s in kept] and d.shape[1] == int((dense[keep].sum(0) > 0).sum())
    rows = bitpack.pack_rows(d.T)
    wv = w[0][keep]
    oW, oF, oC, _ = oracle.scan(rows, d.shape[0], wv, 3, 1)
    last = sd.run.last
    assert abs(last["z"] - max(oW[0], 0.0)) <= 1e-9 * np.abs(wv).sum()
    # model selection can only remove features from the scan optimum
    assert set(module) <= {names[f] for f in oF[0] if f >= 0}
    feats = sorted(names.index(g) for g in module)
    zP, cov, _, _ = oracle.score_set(rows, d.shape[0], feats, wv)
    assert abs(last["zP"] - zP) <= 1e-9 * np.abs(wv).sum()
    # null loop: same nulls through the oracle.  restricted: Curveball on the matrix of the profiled samples; full (default, what
    # the reference's generator + driver do, src/generate_null_matrices.py:53 + src/superdendrix.py:186): Curveball on the matrix of
    # ALL samples, every null restricted to the profiled samples before it is scored
    if scope == "full":
        _, _, names_f, smp_f, _, s2g_f = i_o.load_mutation_data(str(mut), None)
        d_gen = i_o.dense_matrix(names_f, smp_f, s2g_f)
        col_of = [smp_f.index(s_) for s_ in smp]
        feats = sorted(names_f.index(g) for g in module)           # a null may give a feature carriers among the profiled samples that it
    else:                                                          # did not have: the per-null optimum ranges over ALL features of the file
        d_gen = d
    obs = rows_of(d_gen)
    nulls, _ = oracle.curveball_nulls(obs, 9, 0, 64, 5 * obs.shape[0], defaults.CHAIN_LEN, burn=defaults.BURN_ROUNDS, pool_size=defaults.POOL_SIZE,
                                      pair_group=defaults.PAIR_GROUP, L=max(d_gen.shape))     # --mode parallel (default)
    ras = d_gen.shape[1] >= d_gen.shape[0]
    sc = []
    for m in nulls:
        fr = bitpack.pack_rows(bitpack.unpack_rows(m, max(d_gen.shape)).T) if ras else m
        if scope == "full":                                        # feature rows of the driver's genes over the profiled samples
            fr = bitpack.pack_rows(bitpack.unpack_rows(fr, d_gen.shape[0])[:, col_of])
        if stat == "fixed":
            sc.append(oracle.score_set(fr, d.shape[0], feats, wv)[0])
        else:                                                      # re-optimised with k = len(module), as :173,191
            sc.append(max(oracle.scan(fr, d.shape[0], wv, len(module), 1)[0][0], 0.0))
    assert np.allclose(last["null_scores"], sc, rtol=0, atol=1e-9 * np.abs(wv).sum())
    # result row (src/superdendrix.py:241-257)
    head, row = open(out).read().splitlines()
    assert head == "profile\tfeatures\t#samples\tscores\tW(M)(%)\tcoverage\tP_value\tdirection\tfeature_count"
    f = row.split("\t")
    assert f[0] == "GENE1" and set(f[1].split(",")) == set(module) and f[7] == "positive" and int(f[8]) == len(module)
    assert f[5] == "%d/%d" % (cov, d.shape[0])
    assert float(f[6]) == float((np.array(last["null_scores"]) >= last["zP"]).sum()) / 64
    max_score = wv[wv > 0].sum()
    assert abs(float(f[4]) - round(100 * zP / max_score, 4)) < 1e-3
    counts = [int(x) for x in f[2].split(",")]
    assert sorted(counts) == sorted(int(d[:, names.index(g)].sum()) for g in module)


def test_driver_reads_null_files_like_the_reference(tmp_path):
    dense, samples, features, w, mut, prof = _write_case(tmp_path, S=90, F=30, seed=8)
    nm = tmp_path / "rand_mat"
    nm.mkdir()
    gnm.main(["-m", str(mut), "-p", "12", "-o", str(nm) + "/", "-rs", "4"])
    out = tmp_path / "res.tsv"
    args = sd.get_parser().parse_args(["-m", str(mut), "-nm", str(nm) + "/", "-T", str(prof), "-Tc", "GENE2", "-k", "2", "-cp", "-1",
                                       "-p", "12", "-o", str(out), "-d", "positive", "--null-stat", "fixed"])
    module = sd.run(args)
    last = sd.run.last
    profile = {s: float(w[1][i]) for i, s in enumerate(samples)}
    sc = []
    for i in range(12):
        e2s = i_o.load_mutation_data(str(nm / ("%d.tsv" % i)), profile.keys())[4]
        sc.append(refport.super_w([e2s[g] for g in module if g in e2s], profile))
    assert np.allclose(last["null_scores"], sc, rtol=0, atol=1e-9 * np.abs(w[1]).sum())
    assert float(open(out).read().splitlines()[1].split("\t")[6]) == float((np.array(sc) >= last["zP"]).sum()) / 12


def test_driver_reads_null_files_max_k_in_batches(tmp_path):
    """--null-source files with the reference-faithful statistic: every null file re-optimised over |M| <= len(module)
    (src/superdendrix.py:173,191), the files scored in batches by sdx_null_pvalue_rows"""
    dense, samples, features, w, mut, prof = _write_case(tmp_path, S=90, F=30, seed=8)
    nm = tmp_path / "rand_mat"
    nm.mkdir()
    gnm.main(["-m", str(mut), "-p", "10", "-o", str(nm) + "/", "-rs", "4"])
    out = tmp_path / "res.tsv"
    args = sd.get_parser().parse_args(["-m", str(mut), "-nm", str(nm) + "/", "-T", str(prof), "-Tc", "GENE1", "-k", "3", "-cp", "-1",
                                       "-p", "10", "-o", str(out), "-d", "positive", "--null-stat", "max_k", "--null-batch", "4"])
    module = sd.run(args)
    last = sd.run.last
    keep = ~np.isnan(w[0])
    kept = [s for s, k_ in zip(samples, keep) if k_]
    _, _, names, smp, _, _ = i_o.load_mutation_data(str(mut), set(kept))
    wv = w[0][keep]
    want = []
    for i in range(10):
        _, _, _, _, e2s, _ = i_o.load_mutation_data(str(nm / ("%d.tsv" % i)), set(kept))
        d = np.array([[1 if s in e2s.get(g, ()) else 0 for g in names] for s in smp], dtype=np.uint8)
        want.append(max(float(oracle.scan(bitpack.pack_rows(d.T), len(smp), wv, len(module), 1)[0][0]), 0.0))
    assert np.allclose(last["null_scores"], want, rtol=0, atol=1e-9 * np.abs(wv).sum())


def test_null_pvalue_rows_equals_per_null_scoring():
    """sdx_null_pvalue_rows: a batch of caller-supplied nulls for several profiles in one call == score_sets / scan per null"""
    dense, _, _ = synth.feature_matrix(200, 60, seed=2)
    S, F = dense.shape
    w = synth.weights(S, 3, seed=6)
    mask = np.ones((3, S), dtype=bool)
    mask[2, ::5] = False
    modules = np.array([[0, 1, 2], [3, 4, -1], [7, -1, -1]], dtype=np.int32)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w, mask)
        nulls, _ = eng.curveball(5, 0, 40, chain_len=40)                    # feature rows (F < S)
        res = eng.null_pvalue_rows(nulls, modules, want_scores=True)
        mk = eng.null_pvalue_rows(nulls[:12], modules, stat=1, want_scores=True)
        z_fixed, z_maxk = res["z_obs"].copy(), mk["z_obs"].copy()
        per_null_fixed = np.zeros((3, 40)); per_null_maxk = np.zeros((3, 12))
        for i, rows in enumerate(nulls):
            eng.upload_matrix(rows, S)
            eng.upload_weights(w, mask)
            per_null_fixed[:, i] = eng.score_sets(modules, np.arange(3, dtype=np.int32))["W"]
            if i < 12:
                for p in range(3):
                    per_null_maxk[p, i] = max(float(eng.scan(p, int((modules[p] >= 0).sum()), 1)["W"][0]), 0.0)
    assert np.array_equal(res["null_scores"], per_null_fixed)
    assert np.allclose(mk["null_scores"], per_null_maxk, rtol=0, atol=1e-9 * np.abs(w).sum())
    for p in range(3):
        assert int(res["n_ge"][p]) == int((per_null_fixed[p] >= z_fixed[p]).sum())
        assert int(mk["n_ge"][p]) == int((mk["null_scores"][p] >= z_maxk[p]).sum())


def test_optimize_model_k_override_matches_the_drivers_null_statistic(tmp_path):
    """the reference's null loop calls optimize_model with the global k = len(module)
    (src/superdendrix.py:173,191); the replacement takes k=len(module) and then returns the driver's max_k statistic"""
    import argparse
    dense, samples, features, w, mut, prof = _write_case(tmp_path, S=90, F=30, seed=8)
    keep = ~np.isnan(w[0])
    kept = {s for s, k_ in zip(samples, keep) if k_}
    profile = {s: float(w[0][i]) for i, s in enumerate(samples) if keep[i]}
    _, _, genes, smp, e2s, s2g = i_o.load_mutation_data(str(mut), kept)
    args = argparse.Namespace(cardinality=3)
    z3 = sd.optimize_model(genes, smp, profile, e2s, s2g, args)[0]
    z2 = sd.optimize_model(genes, smp, profile, e2s, s2g, args, k=2)[0]
    rows = bitpack.pack_rows(i_o.dense_matrix(genes, smp, s2g).T)
    wv = np.array([profile[s] for s in smp])
    assert abs(z3 - max(oracle.scan(rows, len(smp), wv, 3, 1)[0][0], 0.0)) <= 1e-9 * np.abs(wv).sum()
    assert abs(z2 - max(oracle.scan(rows, len(smp), wv, 2, 1)[0][0], 0.0)) <= 1e-9 * np.abs(wv).sum()
    assert z2 <= z3 + 1e-12
    with pytest.raises(NotImplementedError):
        sd.optimize_model(genes, smp, profile, e2s, s2g, argparse.Namespace(cardinality=-1))


def test_curveball_features_transposes_on_the_device():
    """sdx_curveball_features: the nulls as feature rows whatever the trade orientation"""
    for S, F in ((40, 90), (90, 40)):
        dense = (np.random.default_rng(S).random((S, F)) < 0.15).astype(np.uint8)
        with Engine(0) as eng:
            eng.upload_dense(dense)
            rows, _ = eng.curveball(3, 0, 9, chain_len=4)
            feat = eng.curveball_features(3, 0, 9, chain_len=4)
        want = rows if F < S else np.stack([bitpack.pack_rows(bitpack.unpack_rows(r, F).T) for r in rows])
        assert feat.shape == (9, F, bitpack.words_for(S)) and np.array_equal(feat, want)


def test_fixed_query_cli_like_utils_compute_p_value(tmp_path):
    """utils/compute_p_value.py -q a,b,c: W of the query, Curveball p-value of the FIXED set (configs[0]), rank-sum p"""
    import scipy.stats

    from superdendrix_b200 import compute_p_value as cpv
    dense, samples, features, w, mut, prof = _write_case(tmp_path, S=120, F=40, seed=3)
    order = np.argsort(-dense.sum(0), kind="stable")
    query = [features[order[1]], features[order[2]], features[order[5]]]
    out = tmp_path / "q.tsv"
    args = cpv.get_parser().parse_args(["-m", str(mut), "-T", str(prof), "-Tc", "GENE2", "-q", ",".join(query), "-cp", "-1", "-p", "96",
                                        "-o", str(out), "-d", "positive", "--null-stat", "fixed", "--null-scope", "restricted", "-rs", "5"])
    module = cpv.run(args)
    last = cpv.run.last
    assert sorted(module) == sorted(query)
    # W of the query = the reference's SuperW (oracle/refport.py restates src/superdendrix.py:387-391)
    profile = {s: float(w[1][i]) for i, s in enumerate(samples)}
    cases = [set(samples[i] for i in np.flatnonzero(dense[:, features.index(g)])) for g in query]
    assert abs(last["zP"] - refport.super_w(cases, profile)) <= 1e-9 * np.abs(w[1]).sum()
    # the nulls of the fixed set through the oracle (defaults of --mode parallel)
    _, _, names, smp, _, s2g = i_o.load_mutation_data(str(mut), set(samples))
    d = i_o.dense_matrix(names, smp, s2g)
    obs = rows_of(d)
    nulls, _ = oracle.curveball_nulls(obs, 5, 0, 96, 5 * obs.shape[0], defaults.CHAIN_LEN, burn=defaults.BURN_ROUNDS, pool_size=defaults.POOL_SIZE,
                                      pair_group=defaults.PAIR_GROUP, L=max(d.shape))
    feats = sorted(names.index(g) for g in query)
    ras = d.shape[1] >= d.shape[0]
    sc = [oracle.score_set(bitpack.pack_rows(bitpack.unpack_rows(m_, max(d.shape)).T) if ras else m_, d.shape[0], feats, w[1])[0] for m_ in nulls]
    assert np.allclose(last["null_scores"], sc, rtol=0, atol=1e-9 * np.abs(w[1]).sum())
    assert float(last["p_value"]) == float((np.array(sc) >= last["zP"]).sum()) / 96
    union = dense[:, [features.index(g) for g in query]].any(axis=1)
    assert abs(last["ranksum_pval"] - scipy.stats.ranksums(w[1][union], w[1][~union])[1]) < 1e-12
    head, row = open(out).read().splitlines()
    assert head.split("\t")[:5] == ["target", "aberrations", "#sample", "scores", "p-value"] and row.split("\t")[0] == "GENE2"


def test_ranksum_table_cli_like_utils_compute_ranksum_pval_iter(tmp_path):
    """utils/compute_ranksum_pval_iter.py: rank-sum and Welch t p-values for every row of a results table"""
    import scipy.stats

    from superdendrix_b200 import compute_ranksum_pval_iter as cri
    dense, samples, features, w, mut, prof = _write_case(tmp_path, S=120, F=40, seed=3)
    w = np.nan_to_num(w, nan=0.3)
    synth.write_profile_tsv(str(prof), samples, w, ["GENE1", "GENE2"])
    order = np.argsort(-dense.sum(0), kind="stable")
    table = tmp_path / "results.tsv"
    rows = [("GENE1", [features[order[0]], features[order[3]]], "positive"), ("GENE2_x", [features[order[2]]], "negative"),
            ("GENE9", [features[order[1]]], "positive")]
    with open(table, "w") as fh:
        fh.write("\t".join("c%d" % i for i in range(13)) + "\n")
        for t, mod, dr in rows:
            f = ["."] * 13
            f[1], f[2], f[12] = t, ",".join(mod), dr
            fh.write("\t".join(f) + "\n")
    out = tmp_path / "rs.tsv"
    res = cri.run(cri.get_parser().parse_args(["-m", str(mut), "-q", str(table), "-T", str(prof), "-o", str(out)]))
    assert res[2][3] == "FALSE" and res[2][-1] == "NA"
    for (t, mod, dr), line in zip(rows[:2], res[:2]):
        col = 0 if t.startswith("GENE1") else 1
        vals = (-1.0 if dr == "negative" else 1.0) * w[col]
        union = dense[:, [features.index(g) for g in mod]].any(axis=1)
        assert abs(float(line[5]) - scipy.stats.ranksums(vals[union], vals[~union])[1]) < 1e-12
        assert abs(float(line[6]) - scipy.stats.ttest_ind(vals[union], vals[~union], equal_var=False, nan_policy="omit")[1]) < 1e-10
    assert open(out).read().splitlines()[0] == "profile\tfeatures\tdirection\tprofile_is_in_data\tfeature_is_in_data\tranksum_pval\tt_pval"


def test_too_many_bad_profile_values_aborts_like_the_reference(tmp_path):
    dense, samples, features, w, mut, prof = _write_case(tmp_path, S=50, F=12, seed=1)
    w[1, :10] = np.nan
    synth.write_profile_tsv(str(prof), samples, w, ["GENE1", "GENE2"])
    out = tmp_path / "res.tsv"
    args = sd.get_parser().parse_args(["-m", str(mut), "-T", str(prof), "-Tc", "GENE2", "-k", "2", "-o", str(out)])
    assert sd.run(args) == 1
    assert "incomplete due to high Nan, Inf" in open(out).read()


# ---------------------------------------------------------------- dist helpers over the real engine (world = 1)
def test_dist_functions_single_rank():
    dense, _, _ = synth.feature_matrix(200, 60, seed=2)
    S, F = dense.shape
    w = synth.weights(S, 2, seed=6)
    modules = np.array([[0, 1, 2], [3, 4, -1]], dtype=np.int32)
    with Engine(0) as eng:
        eng.upload_dense(dense)
        eng.upload_weights(w)
        pv = dist.null_pvalue(eng, 3, 50, modules)                   # the defaults: one round from the burnt-in pool, blocks of 32 chains
        direct = eng.null_pvalue(3, 0, 50, modules, chain_len=defaults.CHAIN_LEN)
        assert np.array_equal(pv["n_ge"], direct["n_ge"]) and np.allclose(pv["p"], direct["n_ge"] / 50)
        whole = eng.scan(0, 3, 8)
        # two "ranks" emulated in one process: ranges merge to the whole
        parts = [eng.scan(0, 3, 8, g_lo=lo, g_hi=hi) for lo, hi in dist.scan_feature_ranges(F, 3, 3)]
        merged = dist.merge_hits(parts, 8)
        assert np.array_equal(merged["features"], whole["features"]) and np.array_equal(merged["W"], whole["W"])
        assert sum(p["n_scored"] for p in parts) == whole["n_scored"]
        sc = dist.scan(eng, 0, 3, 8)
        assert np.array_equal(sc["features"], whole["features"]) and sc["n_scored"] == whole["n_scored"]
        cp = dist.condperm_batch(eng, modules, 100, seed=1, profile_of_job=np.array([0, 1], dtype=np.int32))
        assert np.array_equal(cp[0], eng.condperm_batch(modules, 100, 1, 0, np.array([0, 1], dtype=np.int32))[0])
        rs = dist.ranksum(eng, modules, np.array([0, 1], dtype=np.int32))
        assert np.array_equal(rs["twice_rank_sum"], eng.ranksum(modules, np.array([0, 1], dtype=np.int32))["twice_rank_sum"])


# ---------------------------------------------------------------- batched screen (many profiles, one matrix, shared nulls)
def test_batched_screen_equals_per_profile_pipeline(tmp_path):
    from superdendrix_b200 import screen as scr
    dense, samples, features = synth.feature_matrix(150, 40, seed=12)
    S, F = dense.shape
    P = 6
    w = synth.weights(S, P, seed=4)
    mask = np.ones((P, S), bool)
    mask[2, ::9] = False
    w[5] = -np.abs(w[5]) - 0.1                    # nothing to find: empty module
    rows = bitpack.pack_rows(dense.T)
    cond_it, n_nulls, seed = 300, 40, 31
    with Engine(0) as eng:
        res = scr.screen(eng, rows, S, w, mask, k=3, cond_it=cond_it, n_nulls=n_nulls, seed=seed, stat="fixed")
    obs = rows_of(dense)
    ras = F >= S
    nulls, _ = oracle.curveball_nulls(obs, seed, 0, n_nulls, 5 * obs.shape[0], defaults.CHAIN_LEN, burn=defaults.BURN_ROUNDS,
                                      pool_size=defaults.POOL_SIZE, pair_group=defaults.PAIR_GROUP, L=max(dense.shape))
    for p in range(P):
        wm = np.where(mask[p], w[p], 0.0)
        mk = bitpack.pack_mask(np.nonzero(mask[p])[0], S)
        idx = np.nonzero(mask[p])[0]
        tol = 1e-9 * np.abs(wm).sum()
        oW, oF, _, _ = oracle.scan(rows, S, wm, 3, 1, mask=mk)
        if oW[0] <= 0:
            assert (res["modules"][p] == -1).all() and np.isnan(res["p_value"][p])
            continue
        assert abs(res["z"][p] - oW[0]) <= tol
        mod = [int(f) for f in oF[0] if f >= 0]
        rnd = 0
        while len(mod) > 1:                        # src/superdendrix.py:151-158 with the oracle's counts
            counts, _ = oracle.condperm(rows, S, mod, wm, idx, seed, 16 * rnd, cond_it)
            worst, wc = -1, 0
            for i, c in enumerate(counts):
                if c > wc:
                    worst, wc = i, int(c)
            if worst < 0 or wc / float(cond_it) <= 1.0 / cond_it:
                break
            mod.pop(worst)
            rnd += 1
        assert [int(f) for f in res["modules"][p] if f >= 0] == mod
        zP, cov, _, _ = oracle.score_set(rows, S, mod, wm, mk)
        assert abs(res["zP"][p] - zP) <= tol and res["coverage"][p] == cov
        ge = 0
        for m in nulls:
            fr = bitpack.pack_rows(bitpack.unpack_rows(m, max(S, F)).T) if ras else m
            ge += oracle.score_set(fr, S, mod, wm, mk)[0] >= res["zP"][p]
        assert res["n_ge"][p] == ge and res["p_value"][p] == ge / n_nulls
        union = dense[:, mod].any(axis=1)
        zz, pp, _, _ = oracle.ranksum(wm, idx, bitpack.pack_mask(np.nonzero(union)[0], S))
        assert abs(res["ranksum_z"][p] - zz) <= 1e-12 * max(1, abs(zz)) and abs(res["ranksum_p"][p] - pp) <= 1e-12 + 1e-10 * pp
    # the CLI writes one reference-format row per profile
    mut, prof, out = tmp_path / "f.tsv", tmp_path / "p.tsv", tmp_path / "screen.tsv"
    synth.write_features_tsv(str(mut), dense, samples, features)
    synth.write_profile_tsv(str(prof), samples, np.where(mask, w, np.nan), ["P%d" % i for i in range(P)])
    assert scr.main(["-m", str(mut), "-T", str(prof), "-k", "3", "-cp", "100", "-p", "20", "-o", str(out), "-d", "positive"]) == 0
    lines = open(out).read().splitlines()
    assert lines[0] + "\n" == scr.HEADER and len(lines) == P + 1
    # P2 misses 17 of 150 values: more than 10 % NaN, the reference gives up on such a profile (src/superdendrix.py:115-123)
    assert lines[-1].split("\t")[:2] == ["P2", "incomplete due to high Nan, Inf"] and [ln.split("\t")[0] for ln in lines[1:-1]] == ["P0", "P1", "P3", "P4", "P5"]
    assert lines[-2].split("\t")[1:4] == ["nan", "nan", "nan"] and lines[-2].split("\t")[6] == "na"
    f0 = lines[1].split("\t")
    assert f0[0] == "P0" and f0[7] == "positive" and int(f0[8]) == len(f0[1].split(","))
