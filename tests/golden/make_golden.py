#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by IMPORTING THE REFERENCE.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so the
pins are made here from the reference's own functions:

  curveball_replay.npz  src/curveball.py:17-40 run with its module-level
                        ``sample``/``shuffle`` wrapped by a recorder; holds the
                        schedule (a, b, elements dealt to a) and the matrices
                        the reference returned.
  superw.npz            src/superdendrix.py:387-391 on random modules/weights.
  scan.npz              brute force itertools.combinations + SuperW.
  condperm.json         src/superdendrix.py:361-384 with random.seed fixed.
  ranksum.npz           scipy.stats.ranksums (utils/compute_p_value.py:306-318).
  loader.json           src/i_o.py:18-88 on a small TSV with every filter.
  generator/            src/generate_null_matrices.py run as a subprocess.
  nulldist.npz          null distribution of SuperW of a fixed module over 2000 nulls of the reference's own chain.
  nulldist_c0.npz       the same on the bench workload (770 x ~400, bench module): 4000 nulls of the reference's chain.

Nothing from the reference is copied into the repo: only inputs and outputs.
"""
import itertools
import json
import os
import platform
import random
import subprocess
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from superdendrix_b200 import bitpack, synth  # noqa: E402

import curveball as ref_cb  # noqa: E402  (reference)
import i_o as ref_io  # noqa: E402  (reference)

_g = types.ModuleType("gurobipy")
_g.__all__ = []
sys.modules["gurobipy"] = _g
import superdendrix as ref_sd  # noqa: E402  (reference; Gurobi stubbed, ILP unused)


def small_matrix(S, F, seed, density=0.25):
    rng = np.random.default_rng(seed)
    freq = np.clip(density * rng.random(F) * 2, 0.02, 0.9)
    m = (rng.random((S, F)) < freq).astype(np.uint8)
    m[:, m.sum(0) == 0] = 0
    return m


# ------------------------------------------------------------------ curveball
def record_curveball(dense, n_nulls, seed, n_trades=-1):
    """Run the reference curve_ball n_nulls times on one chained presence list."""
    mat = dense.astype(np.float64)
    S, F = mat.shape
    rows_are_samples = F >= S
    L = F if rows_are_samples else S
    wpr = bitpack.words_for(L)
    presence = ref_cb.find_presences(mat)
    R = len(presence)
    sched = {"a": [], "b": [], "to_a": []}
    state = {}
    real_sample, real_shuffle = ref_cb.sample, ref_cb.shuffle

    def rec_sample(pop, k):
        ab = real_sample(pop, k)
        state["a"], state["b"] = ab[0], ab[1]
        sched["a"].append(ab[0]); sched["b"].append(ab[1])
        sched["to_a"].append(np.zeros(wpr, dtype=np.uint32))
        return ab

    def rec_shuffle(tot):
        real_shuffle(tot)
        a, b = state["a"], state["b"]
        L_a = len(presence[a]) - len(set(presence[a]) & set(presence[b]))
        mask = np.zeros(wpr * 32, dtype=bool)
        mask[np.asarray(tot[:L_a], dtype=np.int64)] = True
        sched["to_a"][-1] = bitpack.pack_rows(mask[None, :])[0]

    ref_cb.sample, ref_cb.shuffle = rec_sample, rec_shuffle
    random.seed(seed)
    np.random.seed(seed)
    outs = []
    try:
        for _ in range(n_nulls):
            out = ref_cb.curve_ball(mat, presence, n_trades)
            assert out.shape == dense.shape
            rows = out if rows_are_samples else out.T
            outs.append(bitpack.pack_rows(rows))
    finally:
        ref_cb.sample, ref_cb.shuffle = real_sample, real_shuffle
    T = len(sched["a"]) // n_nulls
    return {
        "dense": dense.astype(np.uint8),
        "a": np.asarray(sched["a"], dtype=np.int32).reshape(n_nulls, T),
        "b": np.asarray(sched["b"], dtype=np.int32).reshape(n_nulls, T),
        "to_a": np.stack(sched["to_a"]).reshape(n_nulls, T, wpr),
        "out": np.stack(outs),
        "seed": np.int64(seed),
    }


def make_curveball():
    cases = {
        "tall_small": record_curveball(small_matrix(12, 9, 1), 3, 11),
        "wide_small": record_curveball(small_matrix(30, 45, 2), 3, 12),
        "square": record_curveball(small_matrix(33, 33, 3), 2, 13),
        "short_run": record_curveball(small_matrix(64, 20, 4, 0.4), 2, 14, n_trades=7),
        "c0_shape": record_curveball(synth.feature_matrix(770, 400, 20)[0], 2, 1764),
    }
    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat["%s__%s" % (name, k)] = v
    flat["python"] = np.array(platform.python_version())
    np.savez_compressed(os.path.join(HERE, "curveball_replay.npz"), **flat)
    print("curveball_replay.npz", {n: c["a"].shape for n, c in cases.items()})


# ------------------------------------------------------------------ SuperW
def weight_variants(S, rng):
    w = synth.weights(S, 1, seed=int(rng.integers(1 << 30)))[0]
    out = {"cont": w.copy()}
    w2 = w.copy(); w2[rng.integers(S, size=5)] = 100.0; w2[rng.integers(S, size=5)] = -100.0
    out["inf100"] = w2                              # +-inf -> +-100 (src/superdendrix.py:96)
    out["unit"] = np.where(w > 0, 1.0, -1.0)       # -u (src/superdendrix.py:107-109)
    w3 = w.copy(); w3[::3] = 0.0
    out["zeros"] = w3
    return out


def make_superw():
    rng = np.random.default_rng(5)
    S, F = 203, 61
    dense = small_matrix(S, F, 6, 0.12)
    samples = ["s%d" % i for i in range(S)]
    cases = {g: set(samples[i] for i in np.nonzero(dense[:, g])[0]) for g in range(F)}
    variants = weight_variants(S, rng)
    mods, kinds, W, cov, ov, ac = [], [], [], [], [], []
    names = list(variants)
    for n in range(1200):
        k = int(rng.integers(1, 6))
        feats = sorted(rng.choice(F, size=k, replace=False).tolist())
        kind = n % len(names)
        w = variants[names[kind]]
        profile = {samples[i]: float(w[i]) for i in range(S)}
        cbe = [cases[g] for g in feats]
        W.append(ref_sd.SuperW(cbe, profile, samples))
        union = set().union(*cbe)
        cov.append(len(union)); ov.append(sum(len(c) for c in cbe) - len(union))
        ac.append(sum(1 for s in union if profile[s] > 0))
        mods.append(feats + [-1] * (5 - k)); kinds.append(kind)
    np.savez_compressed(os.path.join(HERE, "superw.npz"), dense=dense,
                        weights=np.stack([variants[n] for n in names]), weight_names=np.array(names),
                        modules=np.asarray(mods, dtype=np.int32), kind=np.asarray(kinds, dtype=np.int32),
                        W=np.asarray(W), cov=np.asarray(cov, dtype=np.int32),
                        overlap=np.asarray(ov, dtype=np.int32), act_cov=np.asarray(ac, dtype=np.int32))
    print("superw.npz", len(W))


# ------------------------------------------------------------------ scan
def make_scan():
    rng = np.random.default_rng(7)
    out = {}
    for name, (S, F, dens) in {"a": (64, 40, 0.15), "b": (100, 30, 0.3), "unit": (70, 36, 0.2)}.items():
        dense = small_matrix(S, F, int(rng.integers(1 << 30)), dens)
        w = synth.weights(S, 1, seed=int(rng.integers(1 << 30)))[0]
        if name == "unit":
            w = np.where(w > 0, 1.0, -1.0)
        samples = list(range(S))
        profile = {i: float(w[i]) for i in range(S)}
        cases = [set(np.nonzero(dense[:, g])[0].tolist()) for g in range(F)]
        allW = []
        for size in (1, 2, 3):
            for comb in itertools.combinations(range(F), size):
                allW.append((ref_sd.SuperW([cases[g] for g in comb], profile, samples), comb))
        order = sorted(range(len(allW)), key=lambda i: (-allW[i][0], i))[:16]
        out[name + "__dense"] = dense
        out[name + "__w"] = w
        out[name + "__topW"] = np.asarray([allW[i][0] for i in order])
        out[name + "__topF"] = np.asarray([list(allW[i][1]) + [-1] * (3 - len(allW[i][1])) for i in order], dtype=np.int32)
        out[name + "__n"] = np.int64(len(allW))
        # best per size (k = 1, 2, 3 optimum of the ILP with sum x <= k)
        best = []
        for k in (1, 2, 3):
            best.append(max(W for W, c in allW if len(c) <= k))
        out[name + "__best_le_k"] = np.asarray(best)
    np.savez_compressed(os.path.join(HERE, "scan.npz"), **out)
    print("scan.npz")


# ------------------------------------------------------------------ cond perm
def make_condperm():
    rng = np.random.default_rng(9)
    S, F = 150, 30
    dense = small_matrix(S, F, 10, 0.15)
    w = synth.weights(S, 1, seed=11)[0]
    samples = ["s%d" % i for i in range(S)]
    profile = {samples[i]: float(w[i]) for i in range(S)}
    e2c = {"f%d" % g: set(samples[i] for i in np.nonzero(dense[:, g])[0]) for g in range(F)}
    cases = []
    for n in range(6):
        k = int(rng.integers(2, 4))
        feats = sorted(rng.choice(F, size=k, replace=False).tolist())
        module = ["f%d" % g for g in feats]
        random.seed(100 + n)
        ev, cnt = ref_sd.conditional_permutation_test(module, e2c, profile, samples, 3000)
        # per-feature counts, same generator state sequence
        random.seed(100 + n)
        per = []
        cbe = [e2c[e] for e in module]
        real = ref_sd.SuperW(cbe, profile, samples)
        for i, e in enumerate(module):
            c = 0
            for _ in range(3000):
                cbe[i] = random.sample(samples, len(e2c[e]))
                c += ref_sd.SuperW(cbe, profile, samples) >= real
            cbe[i] = e2c[e]
            per.append(int(c))
        cases.append({"feats": feats, "worst": ev, "worst_count": int(cnt), "per_feature": per,
                      "seed": 100 + n, "P": 3000})
    with open(os.path.join(HERE, "condperm.json"), "w") as fh:
        json.dump({"python": platform.python_version(), "cases": cases}, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "condperm_data.npz"), dense=dense, w=w)
    print("condperm.json", [(c["worst"], c["worst_count"]) for c in cases])


# ------------------------------------------------------------------ ranksum
def make_ranksum():
    from scipy.stats import ranksums
    rng = np.random.default_rng(13)
    S = 311
    prof, masks, zs, ps = [], [], [], []
    for n in range(40):
        p = rng.standard_normal(S)
        if n % 3 == 1:
            p = np.round(p, 1)            # ties
        if n % 3 == 2:
            p = np.where(p > 0, 1.0, -1.0)
        m = rng.random(S) < rng.uniform(0.02, 0.6)
        if m.sum() == 0:
            m[0] = True
        z, pv = ranksums(p[m], p[~m])
        prof.append(p); masks.append(m); zs.append(z); ps.append(pv)
    np.savez_compressed(os.path.join(HERE, "ranksum.npz"), profile=np.stack(prof), carrier=np.stack(masks),
                        z=np.asarray(zs), p=np.asarray(ps))
    print("ranksum.npz")


# ------------------------------------------------------------------ loader
def make_loader():
    d = os.path.join(HERE, "loader")
    os.makedirs(d, exist_ok=True)
    feat = os.path.join(d, "features.tsv")
    with open(feat, "w") as fh:
        fh.write("#Samples\tEvents\n")
        fh.write("s1\tA_MUT\tB_MUT\tC_AMP\n")
        fh.write("s2\tA_MUT\n")
        fh.write("s3\n")
        fh.write("s4\tB_MUT\tC_AMP\tD_MUT\n")
        fh.write("s5\tA_MUT\tD_MUT\n")
        fh.write("s6\tE_MUT\n")
    with open(os.path.join(d, "genes.txt"), "w") as fh:
        fh.write("#genes\nA_MUT\nB_MUT\nC_AMP\n")
    calls = {
        "plain": dict(),
        "whitelist": dict(patients=["s1", "s2", "s3", "s4"]),
        "genefile": dict(geneFile=os.path.join(d, "genes.txt")),
        "mut_only": dict(mut_only=True),
        "freq": dict(min_freq=2, max_freq=2),
        "all": dict(patients=["s1", "s2", "s4", "s5"], mut_only=True, min_freq=2),
    }
    res = {}
    for name, kw in calls.items():
        m, n, genes, patients, g2c, p2g = ref_io.load_mutation_data(feat, **kw)
        kwj = {k: (os.path.relpath(v, d) if k == "geneFile" else v) for k, v in kw.items()}
        res[name] = {"kwargs": kwj, "m": m, "n": n, "genes": sorted(genes), "patients": list(patients),
                     "geneToCases": {g: sorted(c) for g, c in g2c.items()},
                     "patientToGenes": {p: sorted(g) for p, g in p2g.items()}}
    with open(os.path.join(HERE, "loader.json"), "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print("loader.json")


# ------------------------------------------------------------------ generator CLI
def make_generator():
    d = os.path.join(HERE, "generator")
    os.makedirs(d, exist_ok=True)
    dense, samples, features = synth.feature_matrix(40, 25, 21)
    dense[3, :] = 0  # a sample without events
    keep = dense.sum(0) > 0
    dense = dense[:, keep]; features = [f for f, k in zip(features, keep) if k]
    feat = os.path.join(d, "features.tsv")
    synth.write_features_tsv(feat, dense, samples, features)
    env = dict(os.environ, PYTHONHASHSEED="0", PYTHONDONTWRITEBYTECODE="1")
    subprocess.check_call([sys.executable, os.path.join(REF, "generate_null_matrices.py"), "-m", feat, "-p", "3",
                           "-o", os.path.join(d, "ref_null_"), "-pre", "5", "-rs", "7"], cwd=REF, env=env,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    print("generator/", sorted(os.listdir(d)))


def make_nulldist():
    """Null distribution of the set objective under the REFERENCE's own generator: curve_ball driven as
    src/generate_null_matrices.py:71-74 drives it (one chain), SuperW (src/superdendrix.py:387-391) of a fixed module on
    every null.  Pins criterion (3) of the north star — p-values within Monte-Carlo error — against the reference
    itself rather than against the oracle."""
    dense, samples, features = synth.feature_matrix(200, 60, 33)
    S, F = dense.shape
    w = synth.weights(S, 1, seed=34)[0]
    module = synth.pick_module(dense, 3)
    profile = {samples[i]: float(w[i]) for i in range(S)}
    random.seed(2024)
    np.random.seed(2024)
    mat = dense.astype(np.float64)
    presence = ref_cb.find_presences(mat)
    n = 2000
    scores = np.zeros(n)
    cell = np.zeros(dense.shape)
    for i in range(n):
        null = ref_cb.curve_ball(mat, presence)
        cases = [set(samples[r] for r in np.flatnonzero(null[:, g])) for g in module]
        scores[i] = ref_sd.SuperW(cases, profile, samples)
        cell += null
    obs_cases = [set(samples[r] for r in np.flatnonzero(dense[:, g])) for g in module]
    np.savez_compressed(os.path.join(HERE, "nulldist.npz"), dense=dense, w=w, module=module, scores=scores,
                        z_obs=ref_sd.SuperW(obs_cases, profile, samples), cell_mean=cell / n,
                        python=platform.python_version())
    print("nulldist.npz", n, "nulls, mean W %.4f sd %.4f" % (scores.mean(), scores.std()))


def make_nulldist_c0():
    """The same on the BENCH workload (BASELINE configs[0] / [1]: 770 x ~400 BRAF-shape matrix, fixed k=3 module): 4 000
    consecutive nulls of the reference's own chain (src/generate_null_matrices.py:71-74), SuperW of the bench module on each.
    The GPU test compares the shipped defaults (burn-in pool, one round per null, blocks of 32 nulls sharing their pairs)
    with the stationary part of this chain; the first nulls of the chain are kept so the test can also see the relaxation."""
    dense, samples, features = synth.feature_matrix(770, 400, seed=20)
    S, F = dense.shape
    w = synth.weights(S, 1, seed=3)[0]
    module = synth.pick_module(dense, 3)
    profile = {samples[i]: float(w[i]) for i in range(S)}
    random.seed(1764)
    np.random.seed(1764)
    mat = dense.astype(np.float64)
    presence = ref_cb.find_presences(mat)
    n = 4000
    scores = np.zeros(n)
    shared = np.zeros(n, dtype=np.int32)
    feat_shared = np.zeros((n, 3), dtype=np.int32)
    for i in range(n):
        null = ref_cb.curve_ball(mat, presence)
        cases = [set(samples[r] for r in np.flatnonzero(null[:, g])) for g in module]
        scores[i] = ref_sd.SuperW(cases, profile, samples)
        both = (null != 0) & (dense != 0)
        shared[i] = int(both.sum())
        feat_shared[i] = [int(both[:, g].sum()) for g in module]
    obs_cases = [set(samples[r] for r in np.flatnonzero(dense[:, g])) for g in module]
    np.savez_compressed(os.path.join(HERE, "nulldist_c0.npz"), module=module, scores=scores, shared=shared, feat_shared=feat_shared,
                        z_obs=ref_sd.SuperW(obs_cases, profile, samples), python=platform.python_version())
    print("nulldist_c0.npz", n, "nulls, mean W %.4f sd %.4f, p %.4f" % (scores[200:].mean(), scores[200:].std(),
                                                                        (scores[200:] >= ref_sd.SuperW(obs_cases, profile, samples)).mean()))


if __name__ == "__main__":
    if "--nulldist-only" in sys.argv:
        make_nulldist()
        sys.exit(0)
    if "--nulldist-c0-only" in sys.argv:
        make_nulldist_c0()
        sys.exit(0)
    make_curveball()
    make_superw()
    make_scan()
    make_condperm()
    make_ranksum()
    make_loader()
    make_generator()
    make_nulldist()
    make_nulldist_c0()
