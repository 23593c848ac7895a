"""Drop-in for the significance-and-search part of ``src/superdendrix.py``.

Same command line (src/superdendrix.py:34-62), same result file
(:241-257), same Python call signatures for the functions on the hot path:

    SuperW(cases_by_event, profile, samples) -> float                                  :387-391
    conditional_permutation_test(module, eventToCases, profile, samples, n) -> (event|None, int)   :361-384
    optimize_model(genes, samples, w, events_to_samples, samples_to_genes, args, logger)            :267-357

What changed underneath: the set objective, the conditional permutations, the
null loop (:166-197) and — for k <= 3 — the search itself run in the sm_100a
kernels of libsdx.so.  ``optimize_model`` here is the exhaustive scan (the
integral optimum of the ILP of :294-313, 393-401); the Gurobi ILP of the
reference is untouched and still what you call for k > 3.

Added, non-breaking flags:
  --null-stat fixed|max_k   statistic of the null loop: W of the selected module on each null
                            (fixed) or the re-optimised optimum over |M| <= k (max_k; what the
                            reference's per-null ILP computes, :191).  Default max_k for k <= 3.
  --null-source generate|files   generate: nulls are produced and scored on the GPU, no files;
                            files: read ``<-nm><i>.tsv`` as the reference does (:186).
  --null-scope full|restricted   generate: Curveball on the full feature matrix, scored on the profiled samples (the
                            reference's generator + driver, default) or on the matrix restricted to the profiled samples
  --null-batch N            files: null files parsed and scored per library call (sdx_null_pvalue_rows)
  --mode parallel|chained|restart   generated nulls: parallel (default) = every null one ordinary round away from a burnt-in
                            pool state (3 burn-in rounds on the balanced orientation), blocks of 32 nulls sharing their row pairs — stationary like
                            the reference's long chain and enough independent units to fill the GPUs; chained = ONE chain from the observed matrix, exactly the reference's
                            semantics (src/generate_null_matrices.py:71-74); restart = every null 5*min(S,F) trades from
                            the observed matrix (src/curveball_main.py:136; NOT mixed for dense features)
  --device N
"""
from __future__ import annotations

import argparse
import logging
import math
import socket
import sys
import time

import numpy as np

from . import bitpack, defaults
from .engine import SDX_STAT_FIXED, SDX_STAT_MAXK, Engine
from .i_o import getLogger, load_events, load_mutation_data, pack_features, read_nulls_tsv

_state = {"engine": None, "device": 0, "seed": 1764, "perm_calls": 0}


def set_device(device: int) -> None:
    if _state["engine"] is not None:
        _state["engine"].close()
        _state["engine"] = None
    _state["device"] = int(device)


def seed(value: int) -> None:
    """Plays the role of ``random.seed(args.random_seed)`` (src/superdendrix.py:76)."""
    _state["seed"] = int(value)
    _state["perm_calls"] = 0


def _engine() -> Engine:
    if _state["engine"] is None:
        _state["engine"] = Engine(_state["device"])        # raises without a GPU: there is no CPU fallback
    return _state["engine"]


def get_parser():
    description = ("Given mutation profiles and a continous target profiles (and integer k), find the set of genes "
                   "(of size k) that maximizes the supervised Dendrix score; significance on a B200")
    parser = argparse.ArgumentParser(description=description, fromfile_prefix_chars="@")
    parser.add_argument("-m", "--mutation_matrix", required=True, help="File name for mutation data")
    parser.add_argument("-nm", "--null_matrices", required=False, help="directory for curveball matrices")
    parser.add_argument("-min_f", "--min_freq", type=int, default=0, help="Minimum gene mutation frequency [0].")
    parser.add_argument("-max_f", "--max_freq", type=int, default=float("inf"), help="Maximum gene mutation frequency [infty].")
    parser.add_argument("-pf", "--sample_file", default=None, help="File of samples to be included (optional).")
    parser.add_argument("-gf", "--gene_file", default=None, help="File of events to be included (optional).")
    parser.add_argument("-k", "--cardinality", default=-1, type=int, help="Fixed cardinality of event subset [-1]")
    parser.add_argument("-M", "--mutations_only", default=False, action="store_true",
                        help="Consider only mutations (events ending with '_MUT')")
    parser.add_argument("-T", "--target", required=True, help="File name for target profile")
    parser.add_argument("-Tc", "--target_column", required=False, type=str, help="Name of column to use in target profile")
    parser.add_argument("-o", "--output_file", help="Name of output file (csv)")
    parser.add_argument("-r", "--reoptimize", default=False, action="store_true", help="Reoptimize for number of genes if k=-1")
    parser.add_argument("-cp", "--cond_it", type=int, default=10000,
                        help="Number of iterations to compute conditional p-value ([-1] = don't)")
    parser.add_argument("-p", "--p_val_it", type=int, default=10000,
                        help="Number of iterations to compute empirical p-value ([-1] = don't)")
    parser.add_argument("-d", "--direction", default="increased_dependency",
                        choices=["increased_dependency", "decreased_dependency", "positive", "negative"],
                        help="direction: match [positive] or negative values")
    parser.add_argument("-u", "--unit_weights", default=False, action="store_true",
                        help="Use unit weights and simple binarization [False]")
    parser.add_argument("-O", "--offset", type=float, default=0, help="manual weight offset [0]")
    parser.add_argument("-v", "--verbosity", default=logging.INFO, type=int, help="Flag verbose output")
    parser.add_argument("-rs", "--random_seed", default=1764, type=int, help="Seed for random number generator")
    parser.add_argument("-t", "--threads", type=int, default=-1, help="Number of threads to use ([-1] = all).")
    parser.add_argument("--null-stat", default=None, choices=["fixed", "max_k"])
    parser.add_argument("--null-source", default=None, choices=["generate", "files"])
    parser.add_argument("--null-scope", default="full", choices=["full", "restricted"],
                        help="generate mode: randomise the full feature matrix and restrict to the profiled samples when scoring, as the "
                             "reference's generator + driver do [full], or randomise the matrix already restricted to the profiled samples")
    parser.add_argument("--mode", default="parallel", choices=["parallel", "chained", "restart"])
    parser.add_argument("--burn-in", type=int, default=defaults.BURN_ROUNDS, help="burn-in rounds of the pool states (parallel mode)")
    parser.add_argument("--pool-size", type=int, default=defaults.POOL_SIZE, help="pool states the chains start from (parallel mode; a multiple of 32).  Nulls that share a state are correlated: a larger pool raises the effective sample size of large -p runs at ~1.4 us per state (DESIGN.md 3b)")
    parser.add_argument("--check-mixing", action="store_true",
                        help="parallel mode: probe whether the burn-in pool is stationary for this matrix (a few hundred extra nulls) and say so in the log")
    parser.add_argument("--device", type=int, default=0)
    parser.add_argument("--null-batch", type=int, default=1024, help="null files parsed and scored per library call (--null-source files)")
    return parser


# ---------------------------------------------------------------------------
# profile parsing (src/superdendrix.py:79-123)
# ---------------------------------------------------------------------------
def parse_profile(path, column, direction="increased_dependency", unit_weights=False, offset=0.0, logger=None):
    """-> (w: {sample: fp64}, n_bad, n_rows).  Sign flip for negative / increased_dependency (:97-98, 105-106),
    +-inf -> +-100 before the flip and counted as bad (:96-99), NaN / blank dropped (:92-94, 100-102),
    unit weights (:107-109)."""
    flip = direction in ("negative", "increased_dependency")
    delim = "\t" if path.endswith(".tsv") else ","
    w, bad, rows = {}, 0, 0
    with open(path) as fh:
        col = fh.readline().rstrip().split(delim).index(column)
        for line in fh:
            if line.startswith("#"):
                continue
            f = line.rstrip().split(delim)
            rows += 1
            try:
                v = float(f[col])
            except (ValueError, IndexError):
                if logger:
                    logger.warning("Warning: %s %s" % (f[0], f[col] if col < len(f) else ""))
                bad += 1
                continue
            if math.isnan(v):
                bad += 1
                continue
            if math.isinf(v):
                v = 100.0 if v > 0 else -100.0
                bad += 1
                w[f[0]] = -v if flip else v
                continue
            v += offset
            if flip:
                v = -v
            if unit_weights:
                v = -1 if v <= 0 else +1
            w[f[0]] = v
    return w, bad, rows


# ---------------------------------------------------------------------------
# hot-path functions under the reference's names
# ---------------------------------------------------------------------------
def _upload_sets(eng, cases_by_event, profile, samples=None):
    """Uploads the k carrier sets as a k x S matrix over the samples of ``profile``."""
    order = list(samples) if samples else list(profile.keys())
    order = [s for s in order if s in profile]
    pos = {s: i for i, s in enumerate(order)}
    dense = np.zeros((max(len(cases_by_event), 1), len(order)), dtype=bool)
    for g, cases in enumerate(cases_by_event):
        dense[g, [pos[s] for s in cases if s in pos]] = True
    eng.upload_matrix(bitpack.pack_rows(dense), len(order))
    eng.upload_weights(np.array([profile[s] for s in order], dtype=np.float64))
    return order


def SuperW(cases_by_event, profile, samples):
    """W(M) = 2 * sum_{s in union, w_s > 0} w_s - sum_g sum_{s in g} |w_s|   (src/superdendrix.py:387-391)."""
    cases_by_event = list(cases_by_event)
    if not cases_by_event:
        return 0.0
    eng = _engine()
    _upload_sets(eng, cases_by_event, profile, samples)
    return float(eng.score_sets(np.arange(len(cases_by_event), dtype=np.int32)[None, :])["W"][0])


def conditional_permutation_test(module, eventToCases, profile, samples, num_permutations):
    """src/superdendrix.py:361-384: for each feature of the module, how often a uniformly random carrier set of
    the same size scores at least as well; returns the feature with the LARGEST count (first wins ties, and a
    module whose counts are all zero returns (None, 0), as the strict '>' of :380 does)."""
    module = list(module)
    eng = _engine()
    cases = [eventToCases[e] for e in module]
    _upload_sets(eng, cases, profile, samples)
    base_id = _state["perm_calls"] * 16
    _state["perm_calls"] += 1
    counts, _ = eng.condperm(np.arange(len(module), dtype=np.int32), int(num_permutations), _state["seed"], perm_id0=base_id)
    worst_count, worst_event = 0, None
    for e, c in zip(module, counts):
        if c > worst_count:
            worst_count, worst_event = int(c), e
    return worst_event, worst_count


def optimize_model(genes, samples, w, events_to_samples, samples_to_genes, args, logger=None, k=None):
    """Same return tuple as src/superdendrix.py:267-357 — (z, module, best_single, cov, act_cov, act_sample, x, y, m) —
    computed by exhaustive scan over all subsets of size <= k (k <= 3).  x, y, m (Gurobi objects) are None.
    M = {} is feasible for the ILP (z = 0), so a negative scan optimum yields the empty module.

    k: the cardinality bound.  The reference's optimize_model reads the MODULE-LEVEL global k (src/superdendrix.py:306),
    which run() sets to args.cardinality (:73) and resets to len(module) before the null loop (:173) so that every null is
    re-optimised over |M| <= len(module) (:191).  A caller that swaps this function into the reference must therefore pass
    k=len(module) inside the null loop; k=None falls back to args.cardinality, the value of the first
    call (:131)."""
    if k is None:
        k = getattr(args, "cardinality", -1)
    if not 1 <= k <= 3:
        raise NotImplementedError("the exhaustive scan certifies k = 1..3 (got k = %r): for the reference's default -k -1 (no cardinality "
                                  "bound) or k > 3 call the reference's own Gurobi optimize_model, optionally warm-started by mip_start()" % (k,))
    eng = _engine()
    genes = list(genes)
    samples = list(samples)
    eng.upload_matrix(pack_features(genes, samples, events_to_samples), len(samples))
    eng.upload_weights(np.array([w[s] for s in samples], dtype=np.float64))
    fs = eng.feature_scores(0)
    best_single = genes[fs["best_single"]] if fs["best_single"] >= 0 else None
    hit = eng.scan(0, k, 1)
    z, feats = float(hit["W"][0]), [int(f) for f in hit["features"][0] if f >= 0]
    if not feats or z < 0:
        z, feats = 0.0, []
    module = [genes[f] for f in feats]
    act_sample = sum(1 for s in samples if w[s] > 0)
    cov = act_cov = 0
    if feats:
        sc = eng.score_sets(np.array([feats + [-1] * (3 - len(feats))], dtype=np.int32))
        cov, act_cov = int(sc["coverage"][0]), int(sc["act_cov"][0])
    return z, module, best_single, cov, act_cov, act_sample, None, None, None


def mip_start(genes, samples, w, events_to_samples, k, logger=None):
    """Certificate / warm start for the reference's Gurobi model (src/superdendrix.py:289-324).

    k <= 3: the exhaustive scan IS the integral optimum, ``certified`` is True and the solve can be skipped.
    k  > 3 (up to SDX_MAX_K = 8): the k=3 optimum is extended greedily, one feature at a time, each step scoring all F
    candidate extensions with the set scorer; the result is a feasible module, i.e. a MIP start and a valid lower
    bound (Gurobi ``Cutoff``) — the ILP itself stays the reference's.

    Returns dict(module=[gene...], x_start={gene: 0/1}, lower_bound=z, certified=bool).  A maintainer adds, after
    ``m.update()`` (src/superdendrix.py:315):  ``for g in genes: x[g].Start = hook["x_start"][g]`` and
    ``m.params.Cutoff = hook["lower_bound"]``."""
    if not 1 <= k <= 8:
        raise ValueError("k must be in 1..8")
    eng = _engine()
    genes = list(genes)
    samples = list(samples)
    eng.upload_matrix(pack_features(genes, samples, events_to_samples), len(samples))
    eng.upload_weights(np.array([w[s] for s in samples], dtype=np.float64))
    hit = eng.scan(0, min(k, 3), 1)
    z, feats = float(hit["W"][0]), [int(f) for f in hit["features"][0] if f >= 0]
    if not feats or z < 0:
        z, feats = 0.0, []
    F = len(genes)
    while feats and len(feats) < k:
        cand = np.full((F, len(feats) + 1), -1, dtype=np.int32)
        cand[:, :len(feats)] = feats
        cand[:, -1] = np.arange(F)
        W = eng.score_sets(cand)["W"]
        W[feats] = -np.inf                                   # a repeated feature is not an extension
        best = int(np.argmax(W))
        if W[best] <= z:
            break
        z, feats = float(W[best]), feats + [best]
    chosen = set(feats)
    return {"module": [genes[f] for f in sorted(feats)], "x_start": {g: int(i in chosen) for i, g in enumerate(genes)},
            "lower_bound": z, "certified": k <= 3}


# ---------------------------------------------------------------------------
def null_loop(genes, samples, profile, events_to_samples, module, zP, args, logger):
    """The permutation test of src/superdendrix.py:166-197 -> (no_better, n, null statistics)."""
    eng = _engine()
    k = len(module)
    stat_name = args.null_stat or ("max_k" if k <= 3 else "fixed")
    stat = SDX_STAT_MAXK if stat_name == "max_k" else SDX_STAT_FIXED
    source = args.null_source or ("files" if args.null_matrices else "generate")
    n = args.p_val_it
    gidx = {g: i for i, g in enumerate(genes)}
    w_vec = np.array([profile[s] for s in samples], dtype=np.float64)
    if source == "generate":
        chain, burn = {"parallel": (defaults.CHAIN_LEN, args.burn_in), "chained": (max(n, 1), 0), "restart": (1, 0)}[args.mode]
        if args.null_scope == "full":
            # What the reference does: its generator randomises the FULL feature matrix — every sample of the file,
            # no profile in sight (src/generate_null_matrices.py:53,62) — and the driver restricts each null to the profiled
            # samples when it loads it (src/superdendrix.py:186, src/i_o.py:57).  Same here: the full matrix is uploaded, the
            # profile is a weight vector with a sample mask, so the nulls keep the margins of the full matrix.
            _, _, genes_f, samples_f, e2s_f, _ = load_mutation_data(args.mutation_matrix, None, args.gene_file, args.mutations_only,
                                                                     args.min_freq, args.max_freq)
            fidx = {g: i for i, g in enumerate(genes_f)}
            missing = [g for g in module if g not in fidx]
            if missing:                                     # frequency window applied to the full matrix dropped a module feature
                raise ValueError("features %s of the module are not in the unrestricted matrix (-min_f / -max_f): use --null-scope restricted" % missing)
            in_profile = np.array([s in profile for s in samples_f], dtype=bool)
            w_full = np.array([profile.get(s, 0.0) for s in samples_f], dtype=np.float64)
            eng.upload_matrix(pack_features(genes_f, samples_f, e2s_f), len(samples_f))
            eng.upload_weights(w_full, in_profile[None, :])
            mod = np.array([[fidx[g] for g in module] + [-1] * (max(k, 1) - k)], dtype=np.int32)
        else:
            # nulls of the matrix already restricted to the profiled samples: keeps the margins of the matrix that is scored
            eng.upload_matrix(pack_features(genes, samples, events_to_samples), len(samples))
            eng.upload_weights(w_vec)
            mod = np.array([[gidx[g] for g in module] + [-1] * (max(k, 1) - k)], dtype=np.int32)
        eng.set_burn_in(burn, args.pool_size)
        eng.set_pair_group(defaults.PAIR_GROUP if args.mode == "parallel" else 1)
        if args.check_mixing and args.mode == "parallel":
            from . import mixing
            fr = pack_features(genes_f, samples_f, e2s_f) if args.null_scope == "full" else pack_features(genes, samples, events_to_samples)
            mixing.check(eng, fr, len(samples_f) if args.null_scope == "full" else len(samples), args.random_seed, log=logger.info if logger else None)
        res = eng.null_pvalue(args.random_seed, 0, n, mod, chain_len=chain, stat=stat, z_obs=np.array([zP]), want_scores=True)
        return int(res["n_ge"][0]), n, res["null_scores"][0]
    # files: <-nm><i>.tsv, restricted to the profiled samples as src/superdendrix.py:186 does.  Read in batches by the
    # library's multi-threaded parser straight into packed rows over this run's sample / event tables (events with no
    # carrier among the profiled samples are all-zero rows, which is what "missing from p_genes" means for W).
    # every batch of parsed nulls is scored in ONE library call (sdx_null_pvalue_rows): fixed = the module on the null,
    # max_k = the per-null optimum over |M| <= len(module) that the reference's ILP returns (src/superdendrix.py:173,191)
    eng.upload_matrix(pack_features(genes, samples, events_to_samples), len(samples))
    eng.upload_weights(w_vec)
    mod = np.array([[gidx[g] for g in module] + [-1] * (max(k, 1) - k)], dtype=np.int32)
    scores = []
    no_better = 0
    for i0 in range(0, n, args.null_batch):
        rows, _ = read_nulls_tsv(args.null_matrices, i0, min(args.null_batch, n - i0), samples, genes)
        res = eng.null_pvalue_rows(rows, mod, stat=stat, z_obs=np.array([zP]), want_scores=True)
        scores.append(res["null_scores"][0])
        no_better += int(res["n_ge"][0])
    return int(no_better), n, np.concatenate(scores) if scores else np.zeros(0)


def run(args):
    logger = getLogger(args.verbosity)
    logger.info("# calling %s" % " ".join(sys.argv))
    logger.info("# at time %s" % time.strftime("%H:%M:%S on %a, %d %b %Y "))
    logger.info("# on machine %s" % socket.gethostname())
    set_device(args.device)
    seed(args.random_seed)
    assert args.target_column
    w, bad, n_rows = parse_profile(args.target, args.target_column, args.direction, args.unit_weights, args.offset, logger)
    if n_rows and float(bad) / n_rows > 0.1:                              # src/superdendrix.py:115-123
        print("too many Nan or Inf, terminating")
        logger.info("Nan, Inf fraction: %s " % (float(bad) / n_rows))
        if args.output_file:
            with open(args.output_file, "w") as of:
                of.write(str(args.target_column) + "\t")
                of.write("incomplete due to high Nan, Inf\t")
                of.write("Nan, Inf fraction: %s" % (float(bad) / n_rows))
        return 1
    load_events(args.mutation_matrix, verbose=1)
    m, N, genes, samples, events_to_samples, samples_to_genes = load_mutation_data(
        args.mutation_matrix, w.keys(), args.gene_file, args.mutations_only, args.min_freq, args.max_freq)
    profile = {s: w[s] for s in samples}
    logger.info("* Mutation data")
    logger.info("\t- Alterations: %s" % m)
    logger.info("\t- Samples: %s" % N)

    t_start = time.time()
    z, module, best_single, cov, act_cov, act_sample, _, _, _ = optimize_model(
        genes, samples, profile, events_to_samples, samples_to_genes, args, logger)
    logger.info("Before model selection, module: " + str(module))
    t_pruning = time.time()
    if args.cond_it != -1:                                                # :151-158
        while len(module) > 1:
            logger.info("Number of features in module:" + str(len(module)))
            worst_event, worst_count = conditional_permutation_test(module, events_to_samples, profile, samples, args.cond_it)
            logger.info("Least significant feature: " + str(worst_event) + str(worst_count / float(args.cond_it)))
            if worst_count / float(args.cond_it) <= (1.0 / float(args.cond_it)):
                break
            module.remove(worst_event)
    logger.info("time for model selection(sec): " + str(time.time() - t_pruning))
    logger.info("module after module selection: " + str(module))

    zP = SuperW([events_to_samples[e] for e in module], profile, samples)
    p_val_str_curve = "na"
    null_scores = []
    if len(module) > 0 and args.p_val_it > 0:                             # :173-197
        t_s1 = time.time()
        no_better, n, null_scores = null_loop(genes, samples, profile, events_to_samples, module, zP, args, logger)
        p_val_str_curve = str(float(no_better) / n)
        print("Time for permutation test:", str((time.time() - t_s1) / 60), "min")

    logger.info("coverage: %d/%d" % (cov, N))
    max_score = sum(profile[p] for p in samples if profile[p] > 0)
    if max_score == 0:
        max_score = -0.1
    scores = {g: float(round(SuperW([events_to_samples[g]], profile, samples), 2)) for g in module}      # :219-227
    scores_ind = [str(v) for v in sorted(scores.values(), reverse=True)]
    module.sort(key=lambda g: len(events_to_samples[g]), reverse=True)
    module.sort(key=lambda g: scores[g], reverse=True)
    cov_ind = [str(len(events_to_samples[g])) for g in module]
    cover = set()
    for g in module:
        cover |= events_to_samples[g]
    if args.output_file:                                                  # :241-257
        with open(args.output_file, "w") as of:
            of.write("profile\tfeatures\t#samples\tscores\tW(M)(%)\tcoverage\tP_value\tdirection\tfeature_count\n")
            of.write(str(args.target_column) + "\t")
            if len(module) > 0:
                of.write(",".join(module))
                of.write("\t" + ",".join(cov_ind))
                of.write("\t" + ",".join(scores_ind))
            else:
                of.write("nan\tnan\tnan")
            of.write("\t" + str(round(100 * zP / max_score, 4)))
            of.write("\t" + str(len(cover)) + "/" + str(N))
            of.write("\t" + str(p_val_str_curve))
            of.write("\t" + str(args.direction))
            of.write("\t" + str(len(module)) + "\n")
    logger.info("Profile: " + args.target_column)
    logger.info("FINAL MODULE: " + ", ".join(module))
    for g in module:
        logger.info("%s %s" % (g, len(events_to_samples[g])))
    run.last = {"z": z, "zP": zP, "module": list(module), "p_value": p_val_str_curve, "null_scores": null_scores,
                "best_single": best_single, "time_s": time.time() - t_start}
    return module


def main(argv=None):
    run(get_parser().parse_args(argv))
    return 0


if __name__ == "__main__":
    sys.exit(main())
