"""Drop-in for ``src/generate_null_matrices.py``: same flags, same output files
(``<out><i + prefix>.tsv`` in the sample -> events format), nulls generated on the GPU.

    python -m superdendrix_b200.generate_null_matrices -m FEATURES.tsv -p N -o DIR/ [-pre OFFSET] [-rs SEED]

Reference flags kept: -m -min_f -max_f -pf -gf -M -o -p -pre -v -rs
(src/generate_null_matrices.py:16-33).  Added, non-breaking:
  --mode chained|parallel|restart
                           chained (default) = the reference's single Markov chain per process
                           (src/generate_null_matrices.py:71-74); parallel = chains of 16 nulls from burnt-in pool
                           states (--burn-in rounds; stationary, many chains); restart = every null from the observed
                           matrix (not mixed for dense features)
  --device N               CUDA device
  --packed FILE.npz        also dump the packed nulls (no TSV parsing needed downstream)
  --no-tsv                 skip the TSV files (with --packed)
Null i of this process has the global index i + prefix: processes (or GPUs) given
disjoint prefixes produce disjoint, reproducible nulls, as the reference's own
sharding by -pre does.
"""
from __future__ import annotations

import argparse
import logging
import socket
import sys
import time

import numpy as np

from . import bitpack, defaults
from .engine import Engine
from .i_o import dense_matrix, getLogger, load_mutation_data, write_nulls_tsv


def get_parser():
    description = "Generate Curveball null mutation matrices (row and column margins preserved) on a B200"
    parser = argparse.ArgumentParser(description=description, fromfile_prefix_chars="@")
    parser.add_argument("-m", "--mutation_matrix", required=True, help="File name for mutation data")
    parser.add_argument("-min_f", "--min_freq", type=int, default=0, help="Minimum gene mutation frequency [0].")
    parser.add_argument("-max_f", "--max_freq", type=int, default=float("inf"), help="Maximum gene mutation frequency [infty].")
    parser.add_argument("-pf", "--sample_file", default=None, help="File of samples to be included (optional).")
    parser.add_argument("-gf", "--gene_file", default=None, help="File of events to be included (optional).")
    parser.add_argument("-M", "--mutations_only", default=False, action="store_true",
                        help="Consider only mutations (events ending with '_MUT')")
    parser.add_argument("-o", "--output_file", help="Prefix (directory) of the output files")
    parser.add_argument("-p", "--p_val_it", type=int, default=-1, help="Number of null matrices ([-1] = none)")
    parser.add_argument("-pre", "--prefix", type=int, default=0, help="index offset of the output files")
    parser.add_argument("-v", "--verbosity", default=logging.INFO, type=int, help="Flag verbose output")
    parser.add_argument("-rs", "--random_seed", default=1764, type=int, help="Seed for random number generator")
    parser.add_argument("--mode", default="chained", choices=["chained", "parallel", "restart"])
    parser.add_argument("--burn-in", type=int, default=defaults.BURN_ROUNDS, help="burn-in rounds of the pool states (parallel mode)")
    parser.add_argument("--pool-size", type=int, default=defaults.POOL_SIZE, help="pool states the chains start from (parallel mode; a multiple of 32).  Nulls that share a state are correlated: a larger pool raises the effective sample size of large -p runs at ~1.4 us per state (DESIGN.md 3b)")
    parser.add_argument("--device", type=int, default=0)
    parser.add_argument("--packed", default=None, help="write the packed nulls to this .npz")
    parser.add_argument("--no-tsv", action="store_true")
    parser.add_argument("--batch", type=int, default=1024, help="nulls per device call")
    return parser


def generate_packed(dense, n_nulls, seed, prefix=0, mode="chained", device=0, batch=1024, engine=None, burn_in=defaults.BURN_ROUNDS,
                    pool_size=defaults.POOL_SIZE):
    """Yields (first global null index, packed feature rows (n, F, ceil(S/32)) uint32) per device batch."""
    S, F = dense.shape
    eng = engine or Engine(device)
    try:
        eng.upload_dense(dense)
        R, L, wpr, rows_are_samples = eng.trade_shape()
        chain_len = {"chained": max(int(n_nulls), 1), "parallel": defaults.CHAIN_LEN, "restart": 1}[mode]
        eng.set_burn_in(burn_in if mode == "parallel" else 0, pool_size)
        eng.set_pair_group(defaults.PAIR_GROUP if mode == "parallel" else 1)
        # a chain starts at a multiple of chain_len: with prefix a multiple of n_nulls every process owns one chain; the
        # library keeps the chain's state between the batches below, so a batch continues where the previous one stopped
        if mode == "chained" and n_nulls > 0 and prefix % max(int(n_nulls), 1):
            import warnings
            warnings.warn("--mode chained: -pre %d is not a multiple of -p %d, so this run covers the end of one chain and the start of the "
                          "next (null i restarts from the observed matrix when i %% p == 0)" % (prefix, n_nulls))
        done = 0
        while done < n_nulls:
            n = min(batch, n_nulls - done)
            rows = eng.curveball_features(seed, prefix + done, n, n_trades=-1, chain_len=chain_len)      # feature rows; transposed on the device if needed
            yield prefix + done, rows
            done += n
    finally:
        if engine is None:
            eng.close()


def generate(dense, n_nulls, seed, prefix=0, mode="chained", device=0, batch=1024, engine=None, burn_in=defaults.BURN_ROUNDS):
    """Yields (global null index, S x F uint8 matrix) for n_nulls nulls of ``dense`` (S x F)."""
    S = dense.shape[0]
    for first, rows in generate_packed(dense, n_nulls, seed, prefix, mode, device, batch, engine, burn_in):
        bits = bitpack.unpack_rows(rows, S)           # (n, F, S)
        for i in range(len(rows)):
            yield first + i, bits[i].T


def main(argv=None):
    args = get_parser().parse_args(argv)
    logger = getLogger(args.verbosity)
    logger.info("# calling %s" % " ".join(sys.argv))
    logger.info("# at time %s" % time.strftime("%H:%M:%S on %a, %d %b %Y "))
    logger.info("# on machine %s" % socket.gethostname())
    m, N, genes, samples, events_to_samples, samples_to_genes = load_mutation_data(
        args.mutation_matrix, None, args.gene_file, args.mutations_only, args.min_freq, args.max_freq)
    logger.info("* Mutation data")
    logger.info("\t- Alterations: %s" % m)
    logger.info("\t- Samples: %s" % N)
    t0 = time.time()
    dense = dense_matrix(genes, samples, samples_to_genes)
    print(len(samples), len(genes))
    packed, ids = [], []
    for first, rows in generate_packed(dense, max(args.p_val_it, 0), args.random_seed, args.prefix, args.mode, args.device, args.batch, burn_in=args.burn_in,
                                       pool_size=args.pool_size):
        if not args.no_tsv:
            write_nulls_tsv(args.output_file, first, rows, samples, genes)      # multi-threaded C++ writer, reference format
        if args.packed:
            packed.append(rows)
            ids.extend(range(first, first + len(rows)))
    if args.packed:
        rows_all = np.concatenate(packed) if packed else np.zeros((0, len(genes), bitpack.words_for(len(samples))), dtype=np.uint32)
        np.savez_compressed(args.packed, feature_rows=rows_all, null_index=np.array(ids, dtype=np.int64),
                            genes=np.array(genes), samples=np.array(samples))
    print(str((time.time() - t0) / 60), "min")
    return 0


if __name__ == "__main__":
    sys.exit(main())
