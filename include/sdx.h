/*
 * include/sdx.h — C ABI of libsdx.so, the B200 (sm_100a) implementation of the
 * SuperDendrix significance-and-search path.
 *
 * The reference (raphael-group/superdendrix) has no FFI: its boundary is a set
 * of Python call signatures.  Each entry point below names the
 * reference function it replaces (path:line relative to the reference root);
 * the Python shims in superdendrix_b200/ keep the reference's names and return
 * types and call these through ctypes.
 *
 * Conventions
 *   - every function returns an int status: 0 = SDX_OK, negative = error;
 *     sdx_last_error(ctx) gives the message.  Nothing aborts the process.
 *   - all pointers are HOST pointers owned by the caller, C-contiguous,
 *     little-endian.  The library owns device memory inside the ctx and copies
 *     in/out itself; functions are synchronous (stream-synchronised on return).
 *   - packed rows: uint32 words, bit j of word i of a row = element 32*i+j;
 *     padding bits are zero and stay zero.
 *   - "feature rows": F rows over S samples (words_per_row = ceil(S/32)).
 *     "trade rows": the orientation Curveball works in — rows over the SHORTER
 *     matrix dimension, as find_presences picks it (src/curveball.py:10-11):
 *     samples are the rows when F >= S, features otherwise.
 *   - one ctx per (process, GPU); calls on one ctx are serialised by the caller.
 *   - there is no CPU fallback: without a CUDA device sdx_create fails.
 */
#ifndef SDX_H
#define SDX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDX_OK 0
#define SDX_ERR_INVALID (-1)     /* bad argument */
#define SDX_ERR_CUDA (-2)        /* CUDA runtime error (message has the detail) */
#define SDX_ERR_STATE (-3)       /* call order: matrix / weights not uploaded */
#define SDX_ERR_UNSUPPORTED (-4) /* shape outside what the kernels cover */

#define SDX_MAX_K 8              /* largest module size the scorer accepts */

#define SDX_STAT_FIXED 0         /* null statistic: W_null(observed module) */
#define SDX_STAT_MAXK 1          /* null statistic: max_{|M|<=k} W_null(M)  (src/superdendrix.py:191) */

typedef struct sdx_ctx sdx_ctx;

/* One scanned subset.  features[] are ascending feature indices, -1 padded. */
typedef struct sdx_hit {
    double W;
    int32_t features[3];
    int32_t coverage;
} sdx_hit;

/* Timings (ms, CUDA events on the ctx stream) of the kernels of the last call. */
typedef struct sdx_timing {
    float plan_ms;     /* trade-schedule kernels (pair draw + level sort)   */
    float trade_ms;    /* Curveball trade (+ fused scoring) kernel          */
    float score_ms;    /* stand-alone scorer / scanner / perm / rank kernels */
    float total_ms;    /* first launch to last launch of the call           */
    int32_t launches;  /* kernels launched by the call                      */
    int32_t trade_launches; /* launches of the trade kernel among them (nulls per launch = n / trade_launches) */
} sdx_timing;

int sdx_version(void);
int sdx_device_count(int *n);

/* lifecycle */
int sdx_create(int device, sdx_ctx **out);
int sdx_destroy(sdx_ctx *ctx);
const char *sdx_last_error(const sdx_ctx *ctx); /* ctx may be NULL: error of a failed sdx_create */
int sdx_set_stream(sdx_ctx *ctx, void *cuda_stream); /* cudaStream_t; NULL = the ctx's own stream */
int sdx_get_timing(const sdx_ctx *ctx, sdx_timing *out);
int sdx_total_launches(const sdx_ctx *ctx, int64_t *n);

/* Binary sample x feature matrix as F packed feature rows over S samples.
 * Replaces the densify step of src/generate_null_matrices.py:62-68 and the
 * geneToCases sets of src/i_o.py:18-88 as the kernels' input. */
int sdx_upload_matrix(sdx_ctx *ctx, const uint32_t *feature_rows, int n_features, int n_samples,
                      int words_per_row);

/* n_profiles weight vectors w[p][s] (fp64, already sign-flipped as
 * src/superdendrix.py:97-98,105-106) and, optionally, per-profile sample masks
 * (samples with a usable profile value, src/superdendrix.py:125,186 via
 * src/i_o.py:57).  Weights of masked-out samples are ignored.  mask == NULL
 * means every sample belongs to every profile. */
int sdx_upload_weights(sdx_ctx *ctx, const double *w, const uint32_t *sample_mask, int n_profiles,
                       int n_samples);

/* SuperW (src/superdendrix.py:387-391) for n_sets modules of up to k features
 * (rows of `sets`, -1 padded) on the uploaded matrix; profile_of_set may be
 * NULL (all profile 0).  Also the integer outputs of src/superdendrix.py:229-236
 * and :345-350: coverage |U|, overlap sum|g|-|U|, act_cov |U ∩ {w>0}|.
 * Any output pointer may be NULL. */
int sdx_score_sets(sdx_ctx *ctx, const int32_t *sets, const int32_t *profile_of_set, int64_t n_sets,
                   int k, double *W, int32_t *coverage, int32_t *overlap, int32_t *act_cov);

/* sum_w[i] = sum of w over the union of set i — the total_score of
 * utils/compute_individual_contribution_SELECT.py:69-85 (m_i_score is the same for the single feature); the
 * leave-one-out contribution W(M) - W(M \ {g}) comes from sdx_score_sets. */
int sdx_union_sums(sdx_ctx *ctx, const int32_t *sets, const int32_t *profile_of_set, int64_t n_sets, int k,
                   double *sum_w);

/* Per-feature outputs for one profile: sum of w over carriers and the argmax
 * (first maximum wins, src/superdendrix.py:281-286), W of the singleton
 * (:219-227, unrounded), carrier count (:229-232). */
int sdx_feature_scores(sdx_ctx *ctx, int profile, double *sum_w, double *W1, int32_t *count,
                       int32_t *best_single);

/* Dimensions of the trade-row orientation of the uploaded matrix. */
int sdx_trade_shape(const sdx_ctx *ctx, int *n_rows, int *n_bits, int *words_per_row,
                    int *rows_are_samples);

/* curve_ball (src/curveball.py:17-40) driven as src/generate_null_matrices.py:71-74
 * drives it: nulls null0 .. null0+n_nulls-1, n_trades each (-1 = 5*min(S,F),
 * src/curveball.py:20).  Null i restarts from the observed matrix when
 * i % chain_len == 0 and continues from null i-1 otherwise (chain_len >= n
 * = the reference's single Markov chain; 1 = independent restarts).  Random
 * draws come from the counter-based stream keyed on (seed, global null index,
 * trade index), so results do not depend on batching or on the GPU count.
 * out_trade_rows: n_nulls x R x words_per_row packed trade rows, or NULL.
 * applied: n_nulls counts of non-skipped trades (src/curveball.py:29), or NULL. */
int sdx_curveball(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                  int64_t chain_len, uint32_t *out_trade_rows, int64_t *applied);

/* The same nulls as FEATURE rows over samples (n_nulls x F x ceil(S/32) words) whatever the trade orientation: the
 * layout of one null file, one line per sample with its events (src/generate_null_matrices.py:75-86).  When the trade
 * rows are samples (F >= S) the batch is transposed on the device. */
int sdx_curveball_features(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                           int64_t chain_len, uint32_t *out_feature_rows);

/* Burn-in.  5*min(S,F) trades do not mix the dense rows: the reference's single chain is
 * stationary only because null i has already seen (i+1) rounds (src/generate_null_matrices.py:73-74).  With
 * burn_rounds > 0 every chain starts, instead of from the observed matrix, from one of pool_size start states:
 * state j = the observed matrix after burn_rounds burn-in rounds drawn from reserved stream ids
 * ((B + j) * burn_rounds + r, r < burn_rounds, B = 2^61 / burn_rounds rounded up to a multiple of 32; the states of an
 * aligned block of 32 share their schedule of row pairs — round r of state j draws its pairs from the stream of round r
 * of state j - j % 32 and deals the elements from its own, cf. sdx_set_pair_group); chain c uses state c % pool_size.
 * The burn-in rounds trade the rows of the orientation of the matrix whose largest row is the smaller multiple of
 * its mean row — the transpose of the trade orientation iff L <= 12000 and max_col * L < max_row * R (R x L = the
 * trade orientation): both orientations keep both margins and have the uniform stationary distribution, but a row
 * far denser than its partners mixes slowly (C0: feature rows need ~30 rounds, sample rows 2).  A burn-in round
 * is 5 * R_b trades (R_b rows of that orientation; at most 60000), independent of n_trades; its pairs are uniform,
 * unless even that orientation has a row above 8 x the mean row: then a trade takes both rows from the
 * max(2, ceil(R_b / 10)) densest rows with probability 1/4.  The high-level paths use burn_rounds = 3.
 * Null indices must stay below 2^61 - 256.  Null i keeps the stream id i, so results still depend only on
 * (seed, i, chain_len, burn_rounds, pool_size) and not on batching or on the GPU count.  The pool is built on first
 * use and cached until the matrix or the seed change.
 * burn_rounds = 0 (default) starts every chain from the observed matrix. */
int sdx_set_burn_in(sdx_ctx *ctx, int64_t burn_rounds, int pool_size);

/* The burn-in pool built in parts — the multi-GPU form of the pool (the reference's counterpart is one burn-in per
 * `-pre` process, src/generate_null_matrices.py:29,71-74): rank r burns a slice of the pool_size states, the slices are
 * exchanged with an all-gather over NVLink directly on the pool's device memory, and every rank commits the same pool.
 *   sdx_pool_begin   allocates the pool for (seed, n_trades; -1 = 5*min(S,F)); *d_pool = its DEVICE address (pool_size
 *                    states of *state_bytes bytes, state j at offset j * state_bytes); the pool is invalid until commit
 *   sdx_pool_build   burns states [first_state, first_state + n_states) into their slots (synchronous)
 *   sdx_pool_commit  declares every slot written (by sdx_pool_build or by the caller's exchange)
 *   sdx_pool_valid   1 when a committed pool for (matrix, seed, n_trades, burn-in) is cached, else 0 (never an error code)
 * A pool built this way is bit-identical to the one sdx_curveball / sdx_null_pvalue build on first use. */
int sdx_pool_begin(sdx_ctx *ctx, uint64_t seed, int64_t n_trades, void **d_pool, size_t *state_bytes);
int sdx_pool_build(sdx_ctx *ctx, int first_state, int n_states);
int sdx_pool_commit(sdx_ctx *ctx);
int sdx_pool_valid(sdx_ctx *ctx, uint64_t seed, int64_t n_trades);

/* How the PAIR stream (which two rows trade t exchanges) is shared between nulls.  pair_group = 1 (default): every
 * null draws its own pairs, stream id = its null index.  pair_group = 32: the chains of an aligned block of 32
 * consecutive chains (chain c = null index / chain_len) walk the SAME schedule of row pairs — round j of chain c draws
 * its pairs from the stream of round j of chain c - c % 32 — while every null keeps its own DEAL stream (who gets
 * which element).  Each null on its own is still curve_ball with uniformly random pairs (src/curveball.py:22); given the
 * pairs the 32 chains are independent, and the uniform distribution is stationary for every fixed pair sequence, so
 * nulls that start from independent stationary states stay independent.  What it buys: the two rows of a trade have
 * the same sizes in all 32 nulls (row sums are invariants), so one warp runs 32 nulls in lock step without
 * divergence — the layout of k_trade_lanes, ~5x fewer instructions per null.  Results depend on
 * (seed, null index, chain_len, burn-in, pair_group) only — not on batching, kernel choice or the GPU count. */
int sdx_set_pair_group(sdx_ctx *ctx, int pair_group);

/* Replay of a recorded schedule (parity with the reference under its own
 * random draws): trade t exchanges rows a[t], b[t]; to_a[t] (words_per_row
 * words) holds the symmetric-difference elements given to row a, i.e.
 * tot[:L] of src/curveball.py:32-34.  start_trade_rows == NULL starts from the
 * uploaded matrix.  Skips are detected on the device (src/curveball.py:29). */
int sdx_curveball_replay(sdx_ctx *ctx, const uint32_t *start_trade_rows, const int32_t *a,
                         const int32_t *b, const uint32_t *to_a, int64_t n_trades,
                         uint32_t *out_trade_rows, int64_t *applied);

/* The null loop of src/superdendrix.py:166-197 without files: generate each
 * null as sdx_curveball does, score it for every uploaded profile and count
 * exceedances n_ge[p] = #{i : stat_i(p) >= z_obs[p]} (:192); p = n_ge/n (:195).
 * modules: n_profiles x k feature indices (-1 padded).  stat = SDX_STAT_FIXED
 * scores the module itself on the null; SDX_STAT_MAXK re-optimises by
 * exhaustive scan over |M| <= k (k <= 3) as the reference's per-null ILP does.
 * z_obs == NULL: the library scores the observed matrix with the same kernel
 * code and returns the values in z_obs_out (may be NULL).
 * null_scores: n_profiles x n_nulls, or NULL. */
int sdx_null_pvalue(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                    int64_t chain_len, int stat, const int32_t *modules, int k, const double *z_obs,
                    int64_t *n_ge, double *null_scores, double *z_obs_out);

/* The same loop over nulls the CALLER supplies — the reference's own form of the loop: it reads null i from
 * <nm><i>.tsv restricted to the profiled samples (src/superdendrix.py:175-192, src/i_o.py:57).  feature_rows:
 * n_nulls x F x words_per_row packed feature rows over the uploaded matrix's samples (host or device memory; e.g. the
 * output of sdx_read_nulls_tsv).  Every null is scored for every uploaded profile in one call: stat = SDX_STAT_FIXED
 * scores module p on the null, SDX_STAT_MAXK re-optimises over |M| <= k_p (k_p = size of module p, :173).
 * z_obs == NULL: the statistic of the uploaded (observed) matrix.  null_scores: n_profiles x n_nulls, or NULL. */
int sdx_null_pvalue_rows(sdx_ctx *ctx, const uint32_t *feature_rows, int64_t n_nulls, int stat, const int32_t *modules,
                         int k, const double *z_obs, int64_t *n_ge, double *null_scores, double *z_obs_out);

/* Exhaustive scan of every feature subset of size 1..k (k <= 3) under W for
 * one profile: the integral optimum of the ILP of src/superdendrix.py:294-313,
 * 393-401.  hits: the topn best subsets, W descending (ties: enumeration
 * order, size-major then lexicographic).  n_scored: number of subsets scored. */
int sdx_scan(sdx_ctx *ctx, int profile, int k, int topn, sdx_hit *hits, int64_t *n_scored);

/* The same restricted to subsets whose smallest feature index lies in
 * [g_lo, g_hi): the unit of multi-GPU sharding (ranks take disjoint ranges and
 * merge their top lists; the counts add up to sdx_scan's). */
int sdx_scan_range(sdx_ctx *ctx, int profile, int k, int topn, int g_lo, int g_hi, sdx_hit *hits,
                   int64_t *n_scored);

/* conditional_permutation_test (src/superdendrix.py:361-384) for one module of
 * one profile: counts[i] = #{t < n_perm : W(M with g_i := random |g_i|-subset
 * of the profile's samples) >= W(M)} (:372-374); the caller applies the
 * argmax / removal rule of :380-382 and :151-158.  real_W[i] (k values, may be
 * NULL): W(M) as used in the comparison of slot i.  Empty slots (-1) get
 * counts[i] = -1.  Draws: stream (PERM, id = perm_id0 + i, t). */
int sdx_condperm(sdx_ctx *ctx, int profile, const int32_t *module, int k, uint64_t seed,
                 uint64_t perm_id0, int64_t n_perm, int64_t *counts, double *real_W);

/* The same for n_jobs (profile, module) pairs in one call (modules: n_jobs x k,
 * counts / real_W: n_jobs x k) — the genome-wide screen of BASELINE configs[4].
 * Every job uses streams perm_id0 .. perm_id0 + k - 1, exactly as n_jobs
 * separate sdx_condperm calls would. */
int sdx_condperm_batch(sdx_ctx *ctx, const int32_t *profile_of_job, const int32_t *modules, int64_t n_jobs,
                       int k, uint64_t seed, uint64_t perm_id0, int64_t n_perm, int64_t *counts,
                       double *real_W);

/* scipy.stats.ranksums as called at utils/compute_p_value.py:306-318 and
 * utils/compute_ranksum_pval.py:116-128 for n_jobs (profile, module) pairs:
 * x = values over carriers of the union, y = over the profile's other samples.
 * `values` are the uploaded weights of that profile.  twice_rank_sum is the
 * exact integer 2*sum(rank(x)) with mid-ranks. */
int sdx_ranksum(sdx_ctx *ctx, const int32_t *profile_of_job, const int32_t *modules, int64_t n_jobs,
                int k, double *z, double *p, int64_t *twice_rank_sum, int32_t *n1);

/* ---- host-side format path (no GPU, no ctx): the sample -> events TSV files ---------------------------
 * The files that carry null matrices from the reference's generator to its driver
 * (writer: src/generate_null_matrices.py:75-86; reader: src/i_o.py:18-88 as called at src/superdendrix.py:186).
 * feature_rows: n_nulls x n_features x words_per_row packed feature rows over samples.  File i is
 * <path_prefix><first_index + i>.tsv.  n_threads <= 0 uses all host threads.  Errors: sdx_io_last_error(). */
int sdx_write_nulls_tsv(const uint32_t *feature_rows, int64_t n_nulls, int n_features, int n_samples,
                        int words_per_row, const char *const *sample_names, const char *const *event_names,
                        const char *path_prefix, int64_t first_index, int n_threads);
/* Reads files written in that format into packed feature rows over the GIVEN sample and event tables
 * (samples not in the table are skipped like the whitelist of src/i_o.py:57; events not in the table are
 * counted in n_unknown_events and skipped like the gene whitelist of :58). */
int sdx_read_nulls_tsv(const char *path_prefix, int64_t first_index, int64_t n_nulls, int n_features, int n_samples,
                       int words_per_row, const char *const *sample_names, const char *const *event_names,
                       uint32_t *feature_rows, int64_t *n_unknown_events, int n_threads);
const char *sdx_io_last_error(void);

/* Measurement hook for bench.py ("popc/INT-pipe fraction against a peak measured on the box"): runs two
 * register-only micro-kernels on the context's device and reports the sustained 32-bit POPC and LOP3 rates (ops/s). */
int sdx_measure_int_peaks(sdx_ctx *ctx, double *popc_per_s, double *lop3_per_s);

/* Self-test hooks used by tests/ (known-answer vectors of the random streams). */
int sdx_test_philox(sdx_ctx *ctx, const uint32_t *ctr4, const uint32_t *key2, int n, uint32_t *out4);
int sdx_test_trade_pairs(sdx_ctx *ctx, uint64_t seed, uint64_t null_id, int n_trades, int n_rows,
                         int32_t *a, int32_t *b);

#ifdef __cplusplus
}
#endif
#endif /* SDX_H */
