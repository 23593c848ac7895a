// sdx_trade_kernel.cuh — k_trade: one CTA per chain, matrix in shared memory (or L2), trades level by level, fused scoring.
// Included by sdx_curveball.cu only.
#pragma once
#include "sdx_nulls.cuh"
#include "sdx_plan.cuh"

// ---------------------------------------------------------------------------
// trade kernel
// ---------------------------------------------------------------------------
struct TradeArgs {
    PhiloxKeys keys;
    NullIndexing ix;
    const uint32_t *observed;         // R x RS
    uint32_t *work;                   // gridDim.x x R x RS (global-resident variant only)
    uint32_t *carry;                  // n_units x R x RS: chain state between launches (or null)
    const uint2 *plan;                // n_local x T: (a | b << 16, t)
    const uint32_t *rt;               // n_local x rt_stride round table (k_trade), see k_plan_sort
    const int *nrounds;               // n_local
    int rt_stride;
    int dcap;                         // schedules at least this deep run in sequential order
    const int *depth;                 // n_local
    int n_units;                      // chain segments in this launch
    int R, RS, T, wprL;
    int save_carry;
    uint32_t *out_rows;               // staging: rows of launch-local null l go to out_rows + l * R * wprL
    unsigned long long *applied;      // (i - ix.null_lo)
    ScoreArgs sc;
    const uint32_t *pool;             // burn-in pool: pool_size start states in the kernel's state layout, or null
    int pool_size;
    int plan_group;                   // 0: one plan per null; g: the chains of an aligned block of g share their plans (pair_group), stored once per block
};

// out-of-line copy of the scorer for the trade kernel: keeps its registers out of the trade loop's budget
static __device__ __noinline__ SetScore score_module_call(const uint32_t *rows, int RS, bool rows_are_samples, const int32_t *mod, int k,
                                                          const double *w, const uint32_t *mask, int S, int lane) {
    return score_module(rows, RS, rows_are_samples, mod, k, w, mask, S, lane);
}

// LB selects the launch bounds: 0 = (512, 1), 1 = (128, 5), 2 = (256, 5), 3 = (256, 4), 4 = (192, 5), 5 = (256, 3), 6 = (64, 5), 7 = (96, 5), 8 = (192, 4), 9 = (160, 4), 10 = (512, 2; global-memory variant) — all
// but 0 bound the registers so that five CTAs (five nulls) stay resident per SM when the matrix fits five times in shared memory.
constexpr int lb_threads(int lb) { return (lb == 0 || lb == 10) ? 512 : (lb == 1 ? 128 : ((lb == 4 || lb == 8) ? 192 : (lb == 6 ? 64 : (lb == 7 ? 96 : (lb == 9 ? 160 : 256))))); }
constexpr int lb_blocks(int lb) { return lb == 0 ? 1 : (lb == 10 ? 2 : ((lb == 3 || lb == 8 || lb == 9) ? 4 : (lb == 5 ? 3 : 5))); }
template <int G, int V, bool SMEM, int LB>
__global__ void __launch_bounds__(lb_threads(LB), lb_blocks(LB)) k_trade(const __grid_constant__ TradeArgs p) {
    extern __shared__ __align__(16) uint32_t smem_rows[];
    __shared__ unsigned int s_applied;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NW = blockDim.x >> 5;
    constexpr int GPW = 32 / G;                 // groups per warp
    const int NG = NW * GPW;                    // groups per CTA
    const int sub = lane & (G - 1);
    const int myg = wid * GPW + (lane / G);
    const int R = p.R, RS = p.RS, T = p.T;
    const int nvec = R * RS / 4;
    uint32_t *rows = SMEM ? smem_rows : p.work + (size_t)blockIdx.x * R * RS;

    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        for (int j = 0; j < p.ix.seg_len; ++j) {
            const int64_t local = (int64_t)u * p.ix.seg_len + j;
            const uint64_t null_id = p.ix.id(local);
            if (null_id >= p.ix.null_hi) break;
            const bool restart = (null_id % (uint64_t)p.ix.chain_len) == 0;
            if (restart || j == 0) {
                const uint32_t *src = p.carry + (size_t)u * R * RS;
                if (restart) src = p.pool ? p.pool + (size_t)((null_id / (uint64_t)p.ix.chain_len) % (uint64_t)p.pool_size) * R * RS : p.observed;
                for (int i = tid; i < nvec; i += blockDim.x)
                    reinterpret_cast<uint4 *>(rows)[i] = reinterpret_cast<const uint4 *>(src)[i];
            }
            if (tid == 0) s_applied = 0;
            __syncthreads();

            // ---- trades, level by level -------------------------------------
            unsigned int my_applied = 0;
            if (T > 0) {
                // One table entry per round (k_plan_sort): first plan entry, number of trades, "last round of its level".
                // The next round's trade and the descriptor after that are in flight while the current trades run.
                // blocks of chains that share their schedule of row pairs share its plan: planned once per block
                const int64_t plocal = p.plan_group ? (int64_t)((p.ix.chain0 + (uint64_t)u) / (uint64_t)p.plan_group - p.ix.chain0 / (uint64_t)p.plan_group) * p.ix.seg_len + j
                                                    : local;
                const uint2 *pl = p.plan + (size_t)plocal * T;
                const uint32_t *rt = p.rt + (size_t)plocal * p.rt_stride;
                const int NR = p.nrounds[plocal];
                const uint2 none = make_uint2(0u, 0u);
                uint32_t q = NR > 0 ? rt[0] : 0u;
                uint32_t q_n = NR > 1 ? rt[1] : 0u;
                uint2 ent_n = myg < (int)((q >> 16) & 0x7fffu) ? pl[(q & 0xffffu) + myg] : none;
                for (int r = 0; r < NR; ++r) {
                    const uint2 ent = ent_n;
                    const bool act = myg < (int)((q >> 16) & 0x7fffu);
                    const bool level_end = (q >> 31) != 0;
                    q = q_n;
                    ent_n = myg < (int)((q >> 16) & 0x7fffu) ? pl[(q & 0xffffu) + myg] : none;
                    q_n = r + 2 < NR ? rt[r + 2] : 0u;
                    // a warp without a trade in this round (the half-empty last round of a level) skips the call
                    bool ap = false;
                    if (__any_sync(0xffffffffu, act))
                        ap = trade_rows<G, V>(rows, RS, act, ent.x & 0xffffu, ent.x >> 16, ent.y, null_id, p.keys, sub);
                    my_applied += (ap && sub == 0) ? 1u : 0u;
                    if (level_end) __syncthreads();
                }
            }
            __syncthreads();

            // ---- outputs -----------------------------------------------------
            const bool emit = null_id >= p.ix.null_lo;
            if (emit && p.applied) {
                my_applied = __reduce_add_sync(0xffffffffu, my_applied);
                if (lane == 0 && my_applied) atomicAdd(&s_applied, my_applied);
            }
            if (emit && p.out_rows) {
                uint32_t *dst = p.out_rows + (size_t)local * R * p.wprL;
                for (int i = tid; i < R * p.wprL; i += blockDim.x) {
                    int r = i / p.wprL, c = i - r * p.wprL;
                    dst[i] = rows[(size_t)r * RS + c];
                }
            }
            if (!SMEM && emit && p.sc.n_profiles > 0 && p.sc.n_ufeat > 0) {
                // Rows are samples and the module features are columns: gather the columns of the distinct module features
                // ONCE per null into shared memory (n_ufeat x wprS words, one ballot per 32 samples and feature) and score
                // every profile from there like a features-as-rows matrix, instead of walking all sample rows per profile.
                uint32_t *cols = smem_rows;
                const int wprS = p.sc.wprS;
                for (int sb = wid; sb < wprS; sb += NW) {
                    const int s = sb * 32 + lane;
                    for (int d = 0; d < p.sc.n_ufeat; ++d) {
                        const int g = p.sc.ufeat[d];
                        const uint32_t bit = s < p.sc.S ? (rows[(size_t)s * RS + (g >> 5)] >> (g & 31)) & 1u : 0u;
                        const uint32_t word = __ballot_sync(0xffffffffu, bit != 0);
                        if (lane == 0) cols[(size_t)d * wprS + sb] = word;
                    }
                }
                __syncthreads();
                for (int pr = wid; pr < p.sc.n_profiles; pr += NW) {
                    SetScore sc = score_module_call(cols, wprS, false, p.sc.umod + (size_t)pr * p.sc.k, p.sc.k,
                                                    p.sc.w + (size_t)pr * p.sc.S, p.sc.mask + (size_t)pr * p.sc.wprS, p.sc.S, lane);
                    if (lane == 0) {
                        if (sc.W >= p.sc.z_obs[pr]) atomicAdd(&p.sc.n_ge[pr], 1ull);
                        if (p.sc.null_scores) p.sc.null_scores[(size_t)pr * p.sc.n_total + (null_id - p.ix.null_lo)] = sc.W;
                    }
                }
            } else if (emit && p.sc.n_profiles > 0) {
                for (int pr = wid; pr < p.sc.n_profiles; pr += NW) {
                    SetScore sc = score_module_call(rows, RS, p.sc.rows_are_samples != 0, p.sc.modules + (size_t)pr * p.sc.k,
                                                    p.sc.k, p.sc.w + (size_t)pr * p.sc.S, p.sc.mask + (size_t)pr * p.sc.wprS,
                                                    p.sc.S, lane);
                    if (lane == 0) {
                        if (sc.W >= p.sc.z_obs[pr]) atomicAdd(&p.sc.n_ge[pr], 1ull);
                        if (p.sc.null_scores) p.sc.null_scores[(size_t)pr * p.sc.n_total + (null_id - p.ix.null_lo)] = sc.W;
                    }
                }
            }
            __syncthreads();
            if (emit && p.applied && tid == 0) p.applied[null_id - p.ix.null_lo] = s_applied;
        }
        if (p.save_carry) {
            uint32_t *dst = p.carry + (size_t)u * R * RS;
            for (int i = tid; i < nvec; i += blockDim.x)
                reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(rows)[i];
        }
        __syncthreads();
    }
}

