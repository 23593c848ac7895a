#!/usr/bin/env python
"""Drop-in for ``utils/compute_ranksum_pval_iter.py``: Wilcoxon rank-sum and Welch t-test p-values for every row of a
SuperDendrix results table (profile, feature set, direction).

Same inputs and output columns as the reference script (utils/compute_ranksum_pval_iter.py:32-45, :144-151): ``-q`` is the
results table (tab separated; profile in column 1, comma separated features in column 2, direction in column 12, header
line first, :49-57), ``-T`` the profile table, ``-m`` the feature file.  For each row the carriers of the union of the
features are compared with the other profiled samples (:124-137): ``scipy.stats.ranksums`` (no tie / continuity correction)
runs on the GPU for all rows in ONE call (sdx_ranksum: exact integer rank sums, z, two-sided normal p); Welch's unequal
variance t-test (``ttest_ind(equal_var=False, nan_policy='omit')``, :137) is a closed form over two means and variances
and is evaluated on the host with the Student-t survival function."""
from __future__ import annotations

import argparse
import logging
import sys

import numpy as np

from . import bitpack
from .engine import Engine
from .i_o import getLogger, load_events, load_mutation_data, load_profiles, pack_features

IND_PROFILE, IND_FEATURE, IND_DIR = 1, 2, 12          # utils/compute_ranksum_pval_iter.py:49-51


def get_parser():
    p = argparse.ArgumentParser(description="rank-sum / Welch t p-values for a SuperDendrix results table (B200)", fromfile_prefix_chars="@")
    p.add_argument("-m", "--mutation_matrix", required=True, help="File name for mutation data")
    p.add_argument("-gf", "--gene_file", default=None, help="File of events to be included (optional).")
    p.add_argument("-q", "--query", required=True, help="query results")
    p.add_argument("-T", "--target", required=True, help="File name for target profile")
    p.add_argument("-Tf", "--target_format", default="achilles", choices=["achilles", "revealer"])
    p.add_argument("-Tc", "--target_column", required=False)
    p.add_argument("-o", "--output_file", help="Name of output file (tsv)")
    p.add_argument("-d", "--direction", default="positive", choices=["positive", "negative"])
    p.add_argument("-v", "--verbosity", default=logging.INFO, type=int)
    p.add_argument("-rs", "--random_seed", default=1764, type=int)
    p.add_argument("-t", "--threads", type=int, default=-1)
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--columns", default=None, help="profile,features,direction column indices of the query table [1,2,12]")
    return p


def welch_t_pvalue(x, y):
    """scipy.stats.ttest_ind(x, y, equal_var=False, nan_policy='omit')[1] (utils/compute_ranksum_pval_iter.py:137)."""
    from scipy import special
    x = np.asarray(x, dtype=np.float64); y = np.asarray(y, dtype=np.float64)
    x = x[~np.isnan(x)]; y = y[~np.isnan(y)]
    n1, n2 = len(x), len(y)
    if n1 < 2 or n2 < 2:
        return float("nan")
    v1, v2 = x.var(ddof=1), y.var(ddof=1)
    se2 = v1 / n1 + v2 / n2
    if se2 == 0:
        return float("nan")
    t = (x.mean() - y.mean()) / np.sqrt(se2)
    df = se2 ** 2 / ((v1 / n1) ** 2 / (n1 - 1) + (v2 / n2) ** 2 / (n2 - 1))
    return float(special.stdtr(df, -abs(t)) * 2)


def run(args, engine=None):
    getLogger(args.verbosity)
    cols = (IND_PROFILE, IND_FEATURE, IND_DIR) if not args.columns else tuple(int(c) for c in args.columns.split(","))
    with open(args.query) as f:
        f.readline()
        arrs = [line.rstrip("\n").split("\t") for line in f if line.strip() and not line.startswith("target")]
    _, mutation_samples = load_events(args.mutation_matrix, verbose=0)
    profiles, profile_samples = load_profiles(args.target, sample_whitelist=mutation_samples, verbose=0)
    m, N, features, samples, events_to_samples, _ = load_mutation_data(args.mutation_matrix, profile_samples, args.gene_file)
    newprofiles = {k.split("_")[0]: v for k, v in profiles.items()}                   # :76-80
    spos = {s: i for i, s in enumerate(profile_samples)}
    fidx = {g: i for i, g in enumerate(features)}
    results, jobs = [], []
    for arr in arrs:
        target, module, direction = arr[cols[0]], [g for g in arr[cols[1]].split(",") if g], arr[cols[2]]
        gene = target.split("_")[0] if "_" in target else target
        line = [target, str(module), direction, "TRUE" if gene in newprofiles else "FALSE",
                str([g + ":" + str(g in features) for g in module])]
        if gene not in newprofiles:
            line.append("NA")                                                         # :116-118
            results.append(line)
            continue
        coef = -1.0 if direction == "negative" else 1.0
        vals = np.array([coef * newprofiles[gene][spos[s]] for s in samples], dtype=np.float64)
        jobs.append((len(results), vals, [fidx[g] for g in module if g in fidx], len(module)))
        results.append(line)
    if jobs:
        eng = engine or Engine(args.device)
        try:
            kmax = max(1, max(len(j[2]) for j in jobs))
            if kmax > 8:
                raise SystemExit("feature sets of more than 8 features are not supported")
            eng.upload_matrix(pack_features(features, samples, events_to_samples), len(samples))
            W = np.stack([j[1] for j in jobs])
            finite = np.isfinite(W)
            eng.upload_weights(np.where(finite, W, 0.0), finite)
            mods = np.full((len(jobs), kmax), -1, dtype=np.int32)
            for r, j in enumerate(jobs):
                mods[r, :len(j[2])] = j[2]
            rs = eng.ranksum(mods, np.arange(len(jobs), dtype=np.int32))
            dense = bitpack.unpack_rows(pack_features(features, samples, events_to_samples), len(samples)).astype(bool)
        finally:
            if engine is None:
                eng.close()
        for r, (idx, vals, feats, n_mod) in enumerate(jobs):
            if n_mod == 0:
                results[idx] += ["1", "nan"]                                          # :126-127 (t_pval is unset there; reported as nan)
                continue
            union = dense[feats].any(axis=0) if feats else np.zeros(len(samples), dtype=bool)
            ok = np.isfinite(vals)
            results[idx] += [str(float(rs["p"][r])), str(welch_t_pvalue(vals[union & ok], vals[~union & ok]))]
    if args.output_file:
        with open(args.output_file, "w") as of:
            of.write("profile\tfeatures\tdirection\tprofile_is_in_data\tfeature_is_in_data\tranksum_pval\tt_pval\n")
            for cur in results:
                of.write("\t".join(cur) + "\n")
    run.last = results
    return results


def main(argv=None):
    run(get_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
