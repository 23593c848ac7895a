"""Batched screen: many dependency profiles against ONE resident feature matrix and ONE set of nulls.

The reference runs one process per profile (``src/superdendrix.py:84-89``; its multi-profile driver
``src/multi-superdendrix.py`` is listed in ``.gitignore:196`` but not shipped) and re-reads the same
``rand_mat/`` files for each (``demo/modules_2_3.sh:14``).  Here the matrix is uploaded once, every profile's
module is found by the exhaustive scan (k <= 3), model selection runs as batched conditional permutations,
and each generated null is scored for all profiles in the same kernel (BASELINE.json configs[3], [4]).

    python -m superdendrix_b200.screen -m FEATURES.tsv -T scores.tsv -k 3 -cp 10000 -p 10000 -o screen.tsv

Output: one row per profile with the columns of the reference's result file (``src/superdendrix.py:243``)
plus rank-sum z / p (``utils/compute_ranksum_pval_iter.py:124-137``).
"""
from __future__ import annotations

import argparse
import sys
import time

import numpy as np

from . import dist
from .engine import SDX_STAT_FIXED, SDX_STAT_MAXK, Engine
from .i_o import load_mutation_data, load_profiles, pack_features

HEADER = "profile\tfeatures\t#samples\tscores\tW(M)(%)\tcoverage\tP_value\tdirection\tfeature_count\tranksum_z\tranksum_p\tFDR\n"


def benjamini_hochberg(p):
    """BH-adjusted p-values (what utils/compute_FDR.py:35 obtains from R's p.adjust(method="BH")); NaN stays NaN."""
    p = np.asarray(p, dtype=np.float64)
    out = np.full(p.shape, np.nan)
    ok = np.flatnonzero(~np.isnan(p))
    if len(ok):
        order = ok[np.argsort(p[ok], kind="stable")]
        n = len(order)
        adj = p[order] * n / np.arange(1, n + 1)
        adj = np.minimum(np.minimum.accumulate(adj[::-1])[::-1], 1.0)
        out[order] = adj
    return out


def select_modules(eng, modules, cond_it, seed):
    """The model-selection loop of src/superdendrix.py:151-158 for every profile at once: while a module has more
    than one feature, drop the feature whose conditional-permutation count is largest, unless that count is <= 1
    (count / cond_it <= 1 / cond_it).  modules: (P, k) int32, -1 padded; returns the pruned copy."""
    mods = np.array(modules, dtype=np.int32, copy=True)
    P, k = mods.shape
    active = np.array([(m >= 0).sum() > 1 for m in mods])
    rnd = 0
    while cond_it > 0 and active.any():
        jobs = np.flatnonzero(active)
        counts, _ = dist.condperm_batch(eng, mods[jobs], cond_it, seed, perm_id0=16 * rnd, profile_of_job=jobs.astype(np.int32))
        for j, c in zip(jobs, counts):
            worst, worst_count = -1, 0
            for slot in range(k):                                   # strict '>' : first feature wins ties (:380-382)
                if mods[j, slot] >= 0 and c[slot] > worst_count:
                    worst, worst_count = slot, int(c[slot])
            if worst < 0 or worst_count / float(cond_it) <= 1.0 / float(cond_it):
                active[j] = False
                continue
            mods[j, worst] = -1
            keep = mods[j][mods[j] >= 0]
            mods[j] = -1
            mods[j, :len(keep)] = keep
            if len(keep) <= 1:
                active[j] = False
        rnd += 1
    return mods


def screen(eng: Engine, feature_rows, n_samples, weights, sample_mask=None, k=3, cond_it=10000, n_nulls=10000, seed=1764,
           stat="fixed", log=None):
    """Returns a dict of per-profile arrays: modules (P, 3), z (scan optimum), zP, n_ge, p_value, coverage,
    feature counts / scores, rank-sum z and p, and timings."""
    t0 = time.time()
    w = np.atleast_2d(np.asarray(weights, dtype=np.float64))
    P = len(w)
    eng.upload_matrix(feature_rows, n_samples)
    eng.upload_weights(w, sample_mask)
    tm = {}
    # 1. search: exhaustive scan per profile (the integral optimum of the ILP, src/superdendrix.py:294-313)
    mods = np.full((P, 3), -1, dtype=np.int32)
    z = np.zeros(P)
    for p in range(P):
        hit = dist.scan(eng, p, k, 1)
        if hit["W"][0] > 0 and hit["features"][0][0] >= 0:
            mods[p] = hit["features"][0]
            z[p] = hit["W"][0]
    tm["scan_s"] = time.time() - t0
    # 2. model selection
    t1 = time.time()
    mods = select_modules(eng, mods, cond_it, seed)
    tm["selection_s"] = time.time() - t1
    # 3. observed statistic of the selected modules, per-feature outputs
    nonempty = mods[:, 0] >= 0
    safe = np.where(nonempty[:, None], mods, np.array([0, -1, -1], dtype=np.int32))
    sc = eng.score_sets(safe, np.arange(P, dtype=np.int32))
    zP = np.where(nonempty, sc["W"], 0.0)
    singles = eng.score_sets(np.where(mods.reshape(-1, 1) >= 0, mods.reshape(-1, 1), 0).astype(np.int32),
                             np.repeat(np.arange(P, dtype=np.int32), 3))
    single_W = np.where(mods >= 0, singles["W"].reshape(P, 3), np.nan)
    single_n = np.where(mods >= 0, singles["coverage"].reshape(P, 3), -1)
    # 4. permutation p-values: every null is scored for all profiles
    t2 = time.time()
    n_ge = np.zeros(P, dtype=np.int64)
    if n_nulls > 0:
        st = SDX_STAT_MAXK if stat == "max_k" else SDX_STAT_FIXED
        n_ge = dist.null_pvalue(eng, seed, n_nulls, safe, stat=st, z_obs=zP)["n_ge"]
    tm["nulls_s"] = time.time() - t2
    # 5. rank-sum of the union's carriers against the rest
    rs = dist.ranksum(eng, safe, np.arange(P, dtype=np.int32))
    tm["total_s"] = time.time() - t0
    if log:
        log("screen: %d profiles, scan %.2fs, selection %.2fs, nulls %.2fs" % (P, tm["scan_s"], tm["selection_s"], tm["nulls_s"]))
    max_score = np.where(w > 0, w, 0.0).sum(axis=1) if sample_mask is None else np.where((w > 0) & np.asarray(sample_mask, bool), w, 0.0).sum(axis=1)
    return {"modules": mods, "z": z, "zP": zP, "n_ge": np.where(nonempty, n_ge, 0), "n_nulls": int(n_nulls),
            "p_value": np.where(nonempty & (n_nulls > 0), n_ge / float(max(n_nulls, 1)), np.nan),
            "coverage": np.where(nonempty, sc["coverage"], 0), "single_W": single_W, "single_n": single_n,
            "ranksum_z": np.where(nonempty, rs["z"], np.nan), "ranksum_p": np.where(nonempty, rs["p"], np.nan),
            "max_score": np.where(max_score == 0, -0.1, max_score), "timing": tm}


def write_results(path, names, genes, res, n_samples_of_profile, direction):
    """One row per profile in the column order of src/superdendrix.py:243-257 (+ rank-sum columns)."""
    fdr = benjamini_hochberg(res["p_value"])
    with open(path, "w") as fh:
        fh.write(HEADER)
        for p, name in enumerate(names):
            feats = [int(f) for f in res["modules"][p] if f >= 0]
            # the reference orders features by per-feature score (descending, rounded to 2 dp), :221-227
            score = [float(round(float(res["single_W"][p][i]), 2)) for i in range(len(feats))]
            # two stable sorts as the reference's: by carrier count, then by rounded score (ties keep the carrier-count order, :202,227)
            order = sorted(range(len(feats)), key=lambda i: -int(res["single_n"][p][i]))
            order = sorted(order, key=lambda i: -score[i])
            cols = [name]
            if feats:
                cols += [",".join(genes[feats[i]] for i in order), ",".join(str(int(res["single_n"][p][i])) for i in order),
                         ",".join(str(score[i]) for i in order)]
            else:
                cols += ["nan", "nan", "nan"]
            pv = res["p_value"][p]
            cols += [str(round(100 * float(res["zP"][p]) / float(res["max_score"][p]), 4)),
                     "%d/%d" % (int(res["coverage"][p]), int(n_samples_of_profile[p])),
                     "na" if np.isnan(pv) else str(float(pv)), direction, str(len(feats)),
                     "nan" if not feats else repr(float(res["ranksum_z"][p])), "nan" if not feats else repr(float(res["ranksum_p"][p])),
                     "na" if np.isnan(fdr[p]) else repr(float(fdr[p]))]
            fh.write("\t".join(cols) + "\n")


def get_parser():
    ap = argparse.ArgumentParser(description="Screen every profile of a score table against one feature matrix on a B200",
                                 fromfile_prefix_chars="@")
    ap.add_argument("-m", "--mutation_matrix", required=True)
    ap.add_argument("-T", "--target", required=True, help="profile table (.tsv / .csv), one column per profile")
    ap.add_argument("-Tc", "--target_columns", default=None, help="comma separated subset of columns (default: all)")
    ap.add_argument("-gf", "--gene_file", default=None)
    ap.add_argument("-M", "--mutations_only", default=False, action="store_true")
    ap.add_argument("-min_f", "--min_freq", type=int, default=0)
    ap.add_argument("-max_f", "--max_freq", type=int, default=float("inf"))
    ap.add_argument("-k", "--cardinality", type=int, default=3)
    ap.add_argument("-cp", "--cond_it", type=int, default=10000)
    ap.add_argument("-p", "--p_val_it", type=int, default=10000)
    ap.add_argument("-d", "--direction", default="increased_dependency",
                    choices=["increased_dependency", "decreased_dependency", "positive", "negative"])
    ap.add_argument("-o", "--output_file", required=True)
    ap.add_argument("-rs", "--random_seed", type=int, default=1764)
    ap.add_argument("-u", "--unit_weights", default=False, action="store_true", help="Use unit weights and simple binarization [False]")
    ap.add_argument("-O", "--offset", type=float, default=0, help="manual weight offset [0]")
    ap.add_argument("--null-stat", default="fixed", choices=["fixed", "max_k"])
    ap.add_argument("--device", type=int, default=0)
    return ap


def main(argv=None):
    args = get_parser().parse_args(argv)
    profiles, psamples = load_profiles(args.target)
    names = list(profiles) if not args.target_columns else args.target_columns.split(",")
    m, N, genes, samples, g2c, p2g = load_mutation_data(args.mutation_matrix, set(psamples), args.gene_file, args.mutations_only,
                                                         args.min_freq, args.max_freq)
    row = {s: i for i, s in enumerate(psamples)}
    idx = np.array([row[s] for s in samples])
    flip = -1.0 if args.direction in ("negative", "increased_dependency") else 1.0            # src/superdendrix.py:97-98,105-106
    raw = np.stack([profiles[nm][idx] for nm in names])
    inf = np.isinf(raw)
    mask = ~np.isnan(raw)                                                                      # NaN / blank: sample not in the profile (:92-94)
    # per profile, exactly as parse_profile: +-inf -> +-100 then the sign flip, no offset / unit weights for those (:96-99); others:
    # offset, flip, then unit weights (:103-109)
    val = np.where(inf, np.sign(raw) * 100.0, np.where(mask, raw, 0.0) + args.offset) * flip
    if args.unit_weights:
        val = np.where(inf, val, np.where(val <= 0, -1.0, 1.0))
    W = np.where(mask, val, 0.0)
    # the reference gives up on a profile with more than 10 % NaN / Inf entries (:115-123): those profiles are reported, not screened
    all_rows = np.stack([profiles[nm] for nm in names])
    bad_frac = (np.isnan(all_rows) | np.isinf(all_rows)).mean(axis=1) if all_rows.shape[1] else np.zeros(len(names))
    skipped = [(nm, float(bf)) for nm, bf in zip(names, bad_frac) if bf > 0.1]
    keep = [i for i, bf in enumerate(bad_frac) if bf <= 0.1]
    names = [names[i] for i in keep]
    W, mask = W[keep], mask[keep]
    if not names:
        with open(args.output_file, "w") as fh:
            fh.write(HEADER)
            for nm, bf in skipped:
                fh.write("%s\tincomplete due to high Nan, Inf\tNan, Inf fraction: %s\n" % (nm, bf))
        return 1
    with Engine(args.device) as eng:
        res = screen(eng, pack_features(genes, samples, g2c), len(samples), W, mask, args.cardinality, args.cond_it, args.p_val_it,
                     args.random_seed, args.null_stat, log=print)
    write_results(args.output_file, names, genes, res, mask.sum(axis=1), args.direction)
    if skipped:
        with open(args.output_file, "a") as fh:
            for nm, bf in skipped:
                fh.write("%s\tincomplete due to high Nan, Inf\tNan, Inf fraction: %s\n" % (nm, bf))
    return 0


if __name__ == "__main__":
    sys.exit(main())
