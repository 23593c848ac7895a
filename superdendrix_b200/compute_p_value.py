#!/usr/bin/env python
"""Drop-in for the fixed-query flow of the reference's ``utils/compute_p_value.py``: the significance of a GIVEN feature
set (``-q a,b,c``) for one profile — the "fixed k=3 set, 1 000 Curveball nulls + permutation p-value" of BASELINE configs[0].

Same steps as utils/compute_p_value.py:185-330: W of the query (:194-196), model selection by conditional permutation
(:204-212), the Curveball permutation test (:215-250; ``-curve`` with ``-nm`` files, or nulls generated on the GPU), the
per-feature scores / coverage (:272-291), the rank-sum p-value (:306-318) and the one-row result table (:321-...).  The
objective is the one of src/superdendrix.py (mutex on, negative part = mutations), computed by libsdx.so.

Null statistic: the reference re-optimises every null over |M| <= len(module) (:239, the per-null ILP) — ``--null-stat max_k``
(default for k <= 3); ``--null-stat fixed`` scores the query itself on every null (configs[0])."""
from __future__ import annotations

import argparse
import logging
import sys
import time
from urllib.parse import quote, urlencode

import numpy as np

from . import defaults, superdendrix as sd
from .i_o import getLogger, load_events, load_mutation_data


def get_parser():
    p = argparse.ArgumentParser(description="Significance of a given feature set for one dependency profile (B200)", fromfile_prefix_chars="@")
    p.add_argument("-m", "--mutation_matrix", required=True, help="File name for mutation data")
    p.add_argument("-nm", "--null_matrices", required=False, help="directory for curveball matrices")
    p.add_argument("-min_f", "--min_freq", type=int, default=0, help="Minimum gene mutation frequency [0].")
    p.add_argument("-max_f", "--max_freq", type=int, default=float("inf"), help="Maximum gene mutation frequency [infty].")
    p.add_argument("-gf", "--gene_file", default=None, help="File of events to be included (optional).")
    p.add_argument("-q", "--query", required=True, help="query alteration(s), comma separated")
    p.add_argument("-M", "--mutations_only", default=False, action="store_true", help="Consider only mutations (events ending with '_MUT')")
    p.add_argument("-T", "--target", required=True, help="File name for target profile")
    p.add_argument("-Tf", "--target_format", default="achilles", choices=["achilles", "revealer"], help="format of target profile [achilles]")
    p.add_argument("-Tc", "--target_column", required=True, help="Name of column to use in target profile")
    p.add_argument("-o", "--output_file", help="Name of output file (tsv)")
    p.add_argument("-cp", "--cond_it", type=int, default=1000, help="Number of iterations to compute conditional p-value ([-1] = don't)")
    p.add_argument("-p", "--p_val_it", type=int, default=-1, help="Number of iterations to compute empirical p-value ([-1] = don't)")
    p.add_argument("-curve", "--curveball", default=False, action="store_true", help="use curveball for p-val")
    p.add_argument("-d", "--direction", default="positive", choices=["positive", "negative", "increased_dependency", "decreased_dependency"])
    p.add_argument("-u", "--unit_weights", default=False, action="store_true")
    p.add_argument("-O", "--offset", type=float, default=0)
    p.add_argument("-v", "--verbosity", default=logging.INFO, type=int)
    p.add_argument("-rs", "--random_seed", default=1764, type=int)
    p.add_argument("--null-stat", default=None, choices=["fixed", "max_k"])
    p.add_argument("--null-source", default=None, choices=["generate", "files"])
    p.add_argument("--null-scope", default="full", choices=["full", "restricted"])
    p.add_argument("--null-batch", type=int, default=1024)
    p.add_argument("--mode", default="parallel", choices=["parallel", "chained", "restart"])
    p.add_argument("--burn-in", type=int, default=defaults.BURN_ROUNDS)
    p.add_argument("--pool-size", type=int, default=defaults.POOL_SIZE)
    p.add_argument("--check-mixing", action="store_true")
    p.add_argument("--device", type=int, default=0)
    return p


def run(args):
    logger = getLogger(args.verbosity)
    sd.set_device(args.device)
    sd.seed(args.random_seed)
    t_start = time.time()
    w, bad, n_rows = sd.parse_profile(args.target, args.target_column, args.direction, args.unit_weights, args.offset, logger)
    load_events(args.mutation_matrix, verbose=0)
    m, N, genes, samples, events_to_samples, samples_to_genes = load_mutation_data(
        args.mutation_matrix, w.keys(), args.gene_file, args.mutations_only, args.min_freq, args.max_freq)
    profile = {s: w[s] for s in samples}
    module = [g for g in args.query.split(",") if g]
    unknown = [g for g in module if g not in events_to_samples]
    if unknown:
        raise SystemExit("query features not in the mutation data: %s" % ", ".join(unknown))
    zP = sd.SuperW([events_to_samples[e] for e in module], profile, samples)          # utils/compute_p_value.py:194-196
    max_score = sum(profile[p] for p in samples if profile[p] > 0)
    if max_score == 0:
        max_score = -0.1
    if args.cond_it != -1:                                                              # :204-212
        while len(module) > 1:
            worst_event, worst_count = sd.conditional_permutation_test(module, events_to_samples, profile, samples, args.cond_it)
            logger.info("worst= " + str(worst_event) + str(worst_count / float(args.cond_it)))
            if worst_count / float(args.cond_it) <= (1.0 / float(args.cond_it)):
                break
            module.remove(worst_event)
        zP = sd.SuperW([events_to_samples[e] for e in module], profile, samples)
    p_val_str, avg_p_z, null_scores = "NA", "NA", []
    if len(module) > 0 and args.p_val_it > 0 and (args.curveball or args.null_source == "generate" or not args.null_matrices):   # :215-250
        no_better, n, null_scores = sd.null_loop(genes, samples, profile, events_to_samples, module, zP, args, logger)
        p_val_str = str(float(no_better) / n)
        avg_p_z = str(float(np.sum(null_scores)) / max(len(null_scores), 1))
    scores = {g: float(round(sd.SuperW([events_to_samples[g]], profile, samples), 2)) for g in module}       # :272-283
    scores_ind = [str(v) for v in sorted(scores.values(), reverse=True)]
    module.sort(key=lambda g: scores[g], reverse=True)
    cov_ind = [str(len(events_to_samples[g])) for g in module]
    cover = set()
    for g in module:
        cover |= events_to_samples[g]
    ranksum_pval = 1.0                                                                  # :306-318
    if module:
        eng = sd._engine()
        gidx = {g: i for i, g in enumerate(genes)}
        from .i_o import pack_features
        eng.upload_matrix(pack_features(genes, samples, events_to_samples), len(samples))
        eng.upload_weights(np.array([profile[s] for s in samples], dtype=np.float64))
        mod = np.array([[gidx[g] for g in module]], dtype=np.int32)
        ranksum_pval = float(eng.ranksum(mod)["p"][0])
    dataset = "Project Achilles" if args.target_format == "achilles" else "Project Revealer"
    queries = urlencode([("dataset", dataset), ("profile", args.target_column), ("sample_lists", "CERES"),
                         ("events", " ".join(module).replace("_MUT", ""))], quote_via=quote)
    t_total = time.time() - t_start
    if args.output_file:
        with open(args.output_file, "w") as of:
            of.write("target\taberrations\t#sample\tscores\tp-value\tz\tmax_score\tz/max_score(%)\tcoverage\tranksum_pval\tbrowser_link\tt_total\tavg_random_z\n")
            of.write(str(args.target_column) + "\t")
            if module:
                of.write(",".join(module) + "\t" + ",".join(cov_ind) + "\t" + ",".join(scores_ind))
            else:
                of.write("nan\tnan\tnan")
            of.write("\t%s\t%s\t%s\t%s\t%d/%d\t%s\t%s\t%s\t%s\n" % (p_val_str, zP, max_score, 100.0 * float(zP) / max_score, len(cover), N, ranksum_pval,
                                                              "https://superdendrix-data-explorer.lrgr.io/#" + queries, t_total, avg_p_z))
    run.last = {"zP": zP, "module": list(module), "p_value": p_val_str, "null_scores": null_scores, "ranksum_pval": ranksum_pval,
                "max_score": max_score}
    return module


def main(argv=None):
    run(get_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
