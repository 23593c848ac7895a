// sdx_nulls.cuh — what the two null-generation paths (k_trade in sdx_curveball.cu, k_trade_lanes in sdx_lanes.cu) share:
// the fused-scoring request and the entry points they offer each other.
#pragma once
#include "sdx_common.cuh"

struct ScoreArgs {
    int n_profiles, k;
    const int32_t *modules;           // P x k
    const double *w;                  // P x S
    const uint32_t *mask;             // P x wprS
    const double *z_obs;              // P
    unsigned long long *n_ge;         // P
    double *null_scores;              // P x n_total, or null
    int64_t n_total;                  // row length of null_scores
    int S, wprS;
    int rows_are_samples;
    // column cache (k_trade, global-memory variant, rows are samples): the distinct features of all modules, and the
    // modules as indices into that list.  n_ufeat = 0: not used.
    const int32_t *ufeat;             // n_ufeat
    const int32_t *umod;              // P x k
    int n_ufeat;
};

struct HostScore {               // optional fused scoring request
    int k = 0;
    const int32_t *modules = nullptr;   // host, P x k
    const double *z_obs = nullptr;      // host, P (or null)
    int64_t *n_ge = nullptr;            // host out, P
    double *null_scores = nullptr;      // host out, P x n (or null)
    double *z_obs_out = nullptr;        // host out, P (or null)
};

// scratch slots of the null paths (sdx_ctx::scratch); 10, 11, 14, 15 belong to sdx_scan.cu
enum { SL_LVL = 0, SL_PLAN, SL_AUX, SL_MOD, SL_Z, SL_SCORES, SL_CARRY, SL_OUT, SL_APPLIED, SL_WORK, SL_PAIRS = 13, SL_RT = 16, SL_LN_STATE = 17, SL_LN_TMP = 18, SL_BURN = 19 };

// sdx_curveball.cu
int sdx_run_nulls(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                  uint32_t *out_rows, int64_t *applied, const HostScore *hs);
int sdx_ensure_pool(sdx_ctx *c, uint64_t seed, int64_t T);      // burn-in pool as bit-packed states in c->d_pool
// sdx_api.cu
int sdx_row_sizes(sdx_ctx *c, const uint32_t *d_rows, int R, int RS, int *d_size);
int sdx_transpose_batch(sdx_ctx *c, const uint32_t *d_in, int n_rows, int wpr_in, int n_cols, uint32_t *d_out, int wpr_out, int nb);
// sdx_lanes.cu
bool sdx_lanes_applicable(sdx_ctx *c, int64_t n_chains, int64_t T, bool fused_scoring);
int sdx_run_nulls_lanes(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t T, int64_t chain_len,
                        uint32_t *out_rows, int64_t *applied, const HostScore *hs);
void sdx_lanes_invalidate(sdx_ctx *c, bool matrix_changed);
