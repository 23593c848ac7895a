// sdx_stats.cu — null-statistic dispatch (fixed / max_k), scan, conditional permutation, rank-sum.
#include "sdx_common.cuh"

int sdx_null_pvalue_fixed(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                          int64_t chain_len, const int32_t *modules, int k, const double *z_obs,
                          int64_t *n_ge, double *null_scores, double *z_obs_out);

extern "C" int sdx_null_pvalue(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                               int64_t chain_len, int stat, const int32_t *modules, int k, const double *z_obs,
                               int64_t *n_ge, double *null_scores, double *z_obs_out) {
    if (!ctx) return SDX_ERR_INVALID;
    if (stat == SDX_STAT_FIXED)
        return sdx_null_pvalue_fixed(ctx, seed, null0, n_nulls, n_trades, chain_len, modules, k, z_obs, n_ge, null_scores, z_obs_out);
    return sdx_fail(ctx, SDX_ERR_UNSUPPORTED, "stat = max_k is not built yet");
}

extern "C" int sdx_scan(sdx_ctx *ctx, int, int, int, sdx_hit *, int64_t *) {
    return ctx ? sdx_fail(ctx, SDX_ERR_UNSUPPORTED, "sdx_scan is not built yet") : SDX_ERR_INVALID;
}
extern "C" int sdx_condperm(sdx_ctx *ctx, int, const int32_t *, int, uint64_t, uint64_t, int64_t, int64_t *, double *) {
    return ctx ? sdx_fail(ctx, SDX_ERR_UNSUPPORTED, "sdx_condperm is not built yet") : SDX_ERR_INVALID;
}
extern "C" int sdx_ranksum(sdx_ctx *ctx, const int32_t *, const int32_t *, int64_t, int, double *, double *, int64_t *, int32_t *) {
    return ctx ? sdx_fail(ctx, SDX_ERR_UNSUPPORTED, "sdx_ranksum is not built yet") : SDX_ERR_INVALID;
}
