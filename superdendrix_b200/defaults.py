"""Defaults of the high-level null paths (dist.null_pvalue, superdendrix.py --mode parallel, screen.py, bench.py).

Every null is ONE ordinary Curveball round (5*min(S,F) uniformly drawn trades, src/curveball.py:20-22) away from a
burnt-in pool state, and aligned blocks of 32 chains share their schedule of row pairs (include/sdx.h:
sdx_set_pair_group) — the layout of the one-thread-per-null kernel k_trade_lanes.  The pool size is a multiple of 32
so that the start states of a block are one contiguous group of the pool."""
PAIR_GROUP = 32
CHAIN_LEN = 1
BURN_ROUNDS = 3          # burn-in rounds in the balanced orientation of the matrix: stationary after 2 at C0
POOL_SIZE = 1472         # 46 groups of 32 states
