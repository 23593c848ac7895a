// sdx_curveball.cu — Curveball null generation (+ fused fixed-set scoring).
//
// Replaces, behind the C ABI of include/sdx.h:
//   src/curveball.py:17-40            curve_ball
//   src/generate_null_matrices.py:71-74  the null loop over one mutable presence list
//   src/superdendrix.py:166-197       the null loop / exceedance count (stat = fixed)
//
// Pipeline per batch of nulls (all on the ctx stream):
//   k_plan_levels   [sdx_plan.cuh] one THREAD per null draws the row pair of each of its trades (PAIR stream), walks them in
//                   order and assigns each the level
//                   1 + max(level of the last trade that touched row a, row b).
//                   Trades of one level touch pairwise disjoint rows, and every
//                   earlier trade on those rows is in a lower level, so running
//                   level by level is identical to the sequential order.
//   k_plan_sort     one WARP per null: counting sort of the trades by level into
//                   the plan (a | b << 16, t) plus the level offsets.
//   k_trade         [sdx_trade_kernel.cuh] one CTA per chain segment: matrix resident in shared memory
//                   (or in L2/global when it does not fit), groups of G lanes
//                   execute the trades of a level concurrently (plan entries
//                   prefetched one round ahead), __syncthreads between levels;
//                   then the fused scorer and the outputs.
//   experimental variants (opt-in, bit-identical): k_trade_sm [sdx_trade_sm.cuh], k_trade_sparse [sdx_trade_sparse.cuh]
// This file: host side (launch geometry, burn-in pool, chains across launches), replay, the C entry points.
#include <stdlib.h>

#include <algorithm>

#include "sdx_plan.cuh"
#include "sdx_trade_kernel.cuh"
#include "sdx_trade_sm.cuh"
#include "sdx_trade_sparse.cuh"

// score the observed matrix with the very same device code (z_obs for the exceedance test)
__global__ void k_score_observed(const uint32_t *rows, int RS, ScoreArgs sc, double *z_out) {
    const int lane = threadIdx.x & 31;
    const int pr = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pr >= sc.n_profiles) return;
    SetScore s = score_module(rows, RS, sc.rows_are_samples != 0, sc.modules + (size_t)pr * sc.k, sc.k,
                              sc.w + (size_t)pr * sc.S, sc.mask + (size_t)pr * sc.wprS, sc.S, lane);
    if (lane == 0) z_out[pr] = s.W;
}

// ---------------------------------------------------------------------------
// replay of a recorded schedule: one warp, strictly sequential
// ---------------------------------------------------------------------------
__global__ void k_replay(uint32_t *rows, int RS, int wprL, const int32_t *a, const int32_t *b,
                         const uint32_t *to_a, int64_t T, unsigned long long *applied) {
    const int lane = threadIdx.x;
    unsigned long long ap = 0;
    for (int64_t t = 0; t < T; ++t) {
        uint32_t *ra = rows + (size_t)a[t] * RS, *rb = rows + (size_t)b[t] * RS;
        const uint32_t *m = to_a + (size_t)t * wprL;
        int la = 0, lb = 0;
        for (int i = lane; i < wprL; i += 32) {
            uint32_t x = ra[i], y = rb[i];
            la += __popc(x & ~y); lb += __popc(y & ~x);
        }
        la = warp_sum_i32(la); lb = warp_sum_i32(lb);
        if (la != 0 && lb != 0) {
            for (int i = lane; i < wprL; i += 32) {
                uint32_t x = ra[i], y = rb[i], ab = x & y, sym = x ^ y, mk = m[i];
                ra[i] = ab | (sym & mk);
                rb[i] = ab | (sym & ~mk);
            }
            ++ap;
        }
        __syncwarp();
    }
    if (lane == 0) *applied = ap;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct TradeConfig {
    int slots;         // CTAs of the trade kernel resident on the whole GPU (one wave)
    int G, V, LB;
    bool smem;
    int threads;
    size_t smem_bytes;
    int grid_cap;      // max CTAs (global variant keeps the working set near L2 size)
    int persist_slots; // k_trade_sm: null slots per SM (0 = not applicable)
    size_t persist_smem;
};

static int pick_config(sdx_ctx *c, TradeConfig *cfg) {
    const int w = c->wprL;
    if (w <= 32) { cfg->G = 4; cfg->V = 8; }
    else if (w <= 64) { cfg->G = 8; cfg->V = 8; }
    else if (w <= 128) { cfg->G = 32; cfg->V = 4; }
    else if (w <= 160) { cfg->G = 32; cfg->V = 5; }
    else if (w <= 256) { cfg->G = 32; cfg->V = 8; }
    else return sdx_fail(c, SDX_ERR_UNSUPPORTED, "trade rows longer than 8192 bits are not supported");
    if (c->R > 65535) return sdx_fail(c, SDX_ERR_UNSUPPORTED, "more than 65535 trade rows are not supported");
    // tuning knobs (benchmarks only): SDX_TRADE_G in {2,4,8,16,32}, SDX_TRADE_WARPS
    int warps_env = 0;
    if (const char *e = getenv("SDX_TRADE_G")) {
        const int g = atoi(e);
        static const int gv[][2] = {{2, 16}, {4, 8}, {8, 4}, {8, 8}, {8, 20}, {16, 4}, {32, 4}, {32, 5}, {32, 8}};
        for (auto &q : gv)
            if (q[0] == g && q[0] * q[1] >= c->RS) { cfg->G = q[0]; cfg->V = q[1]; break; }
    }
    if (const char *e = getenv("SDX_TRADE_WARPS")) warps_env = atoi(e);
    const size_t bytes = (size_t)c->R * c->RS * 4;
    const size_t budget = c->smem_optin > 2048 ? c->smem_optin - 1024 : 0;
    cfg->smem = bytes <= budget;
    cfg->persist_slots = 0;
    cfg->persist_smem = 0;
    if (cfg->smem) {
        // persistent work-queue kernel: as many null slots as fit next to its control block, all outstanding rounds in the ring
        const size_t per_slot = (bytes < 16 ? 16 : bytes) + SM_OFFS_CAP * sizeof(uint16_t);
        const size_t avail = c->smem_optin > sizeof(SmShared) + 256 ? c->smem_optin - sizeof(SmShared) - 256 : 0;
        int ns = (int)(avail / per_slot);
        if (ns > SM_MAX_SLOTS) ns = SM_MAX_SLOTS;
        const int gpw = 32 / cfg->G;
        const int rounds_per_level = (c->R / 2 + gpw - 1) / gpw + SM_EPI_ROUNDS;
        const char *pe = getenv("SDX_TRADE_PERSIST");
        // EXPERIMENTAL (SDX_TRADE_PERSIST=1): bit-identical, but measured 37 % slower than k_trade on C1 — see
        // profiles/r01_k_trade_ncu_summary.md ("work-queue kernel")
        if (ns >= 2 && ns * rounds_per_level <= SM_RING - 32 && pe && atoi(pe) == 1) {
            cfg->persist_slots = ns;
            cfg->persist_smem = ns * per_slot;
        }
        cfg->smem_bytes = bytes < 16 ? 16 : bytes;
        if (const char *e = getenv("SDX_TRADE_PAD_SMEM")) {       // experiments only: fewer resident CTAs per SM
            const size_t pad = (size_t)atol(e);
            if (cfg->smem_bytes + pad <= budget) cfg->smem_bytes += pad;
        }
        int per_sm = (int)((228 * 1024) / (cfg->smem_bytes + 1024 + 64));
        if (per_sm < 1) per_sm = 1;
        int warps = 40 / per_sm;                  // aim at ~40 resident warps per SM
        if (warps < 4) warps = 4;
        if (warps > 16) warps = 16;
        if (per_sm >= 4 && warps > 4) warps = 4;  // register file: 5 CTAs x 128 threads x 96 registers
        if (warps_env >= 1 && warps_env <= 16) warps = warps_env;
        cfg->threads = warps * 32;
        cfg->LB = per_sm >= 4 ? (cfg->threads <= 128 ? 1 : (cfg->threads <= 256 ? 2 : 0)) : 0;
        if (const char *e = getenv("SDX_TRADE_LB")) {
            const int lb = atoi(e);
            if (lb >= 0 && lb <= 9 && cfg->threads <= lb_threads(lb)) cfg->LB = lb;
        }
        cfg->grid_cap = 1 << 30;
        {
            int resident = per_sm;                                      // shared memory limit
            const int by_threads = 2048 / cfg->threads;
            if (resident > by_threads) resident = by_threads;
            if (cfg->LB != 0 && resident > lb_blocks(cfg->LB)) resident = lb_blocks(cfg->LB);
            cfg->slots = resident * c->sm_count;
        }
    } else {
        cfg->smem_bytes = 0;
        // Rows live in global memory / L2 (C3: 1 000 rows of 157 words).  Measured on the C3 slice, 4 096 nulls, trade kernel only:
        //   G = 32, V = 5, 16 warps, 1 CTA / SM (128 registers), 150 CTAs on 148 SMs      53.9 ms  (two waves: the first version)
        //   the same with a grid of whole waves (148 CTAs)                                 32.1 ms
        //   G = 32, V = 5, 16 warps, 2 CTAs / SM at 64 registers (launch bounds (512, 2))  29.6 ms
        //   G = 8, V = 20, 4 warps, 4 CTAs / SM at 128 registers                           18.7 ms  <- rows of 129..160 words
        // Fewer lanes per trade means more trades per instruction for the serial parts (Philox block, Floyd loop), and
        // the grid is always a whole number of resident CTAs per SM.
        int warps = 16, ctas_per_sm = 1;
        if (c->RS > 128 && c->RS <= 160 && !getenv("SDX_TRADE_G")) { cfg->G = 8; cfg->V = 20; }
        if (cfg->G == 8 && cfg->V == 20) { cfg->LB = 0; warps = 4; ctas_per_sm = 4; }
        else if (cfg->V <= 5) { cfg->LB = 10; ctas_per_sm = 2; }      // 8 words per lane spill too much at 64 registers
        else cfg->LB = 0;
        if (const char *e = getenv("SDX_TRADE_LB")) { cfg->LB = atoi(e) == 10 ? 10 : 0; }
        if (warps_env >= 1 && warps_env <= 16) warps = warps_env;
        cfg->threads = warps * 32;
        size_t per = bytes;
        const int per_wave = ctas_per_sm * c->sm_count;
        int cap = (int)((size_t)192 * 1024 * 1024 / per);
        cap = cap < per_wave ? per_wave : (cap / per_wave) * per_wave;
        if (const char *e = getenv("SDX_TRADE_GRID_CAP")) { const int v = atoi(e); if (v >= 1) cap = v; }   // experiments only
        cfg->grid_cap = cap;
        cfg->slots = cap;
    }
    return SDX_OK;
}

template <int G, int V>
static cudaError_t launch_trade(const TradeConfig &cfg, int grid, cudaStream_t st, const TradeArgs &args, bool persist) {
    if (persist) {
        cudaError_t e = cudaFuncSetAttribute(k_trade_sm<G, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.persist_smem);
        if (e != cudaSuccess) return e;
        k_trade_sm<G, V><<<grid, SM_THREADS, cfg.persist_smem, st>>>(args, cfg.persist_slots);
        return cudaGetLastError();
    }
    if (cfg.smem) {
        auto kern = cfg.LB == 1 ? k_trade<G, V, true, 1> : (cfg.LB == 2 ? k_trade<G, V, true, 2> : (cfg.LB == 3 ? k_trade<G, V, true, 3> : (cfg.LB == 4 ? k_trade<G, V, true, 4> : (cfg.LB == 5 ? k_trade<G, V, true, 5> : (cfg.LB == 6 ? k_trade<G, V, true, 6> : (cfg.LB == 7 ? k_trade<G, V, true, 7> : (cfg.LB == 8 ? k_trade<G, V, true, 8> : (cfg.LB == 9 ? k_trade<G, V, true, 9> : k_trade<G, V, true, 0>))))))));
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<grid, cfg.threads, cfg.smem_bytes, st>>>(args);
    } else {
        const size_t cols = sizeof(uint32_t) * (size_t)args.sc.n_ufeat * args.sc.wprS;      // column cache of the scorer (<= 48 KB)
        if (cfg.LB == 10) k_trade<G, V, false, 10><<<grid, cfg.threads, cols, st>>>(args);
        else k_trade<G, V, false, 0><<<grid, cfg.threads, cols, st>>>(args);
    }
    return cudaGetLastError();
}

static cudaError_t dispatch_trade(const TradeConfig &cfg, int grid, cudaStream_t st, const TradeArgs &args, bool persist) {
    if (cfg.G == 2 && cfg.V == 16) return launch_trade<2, 16>(cfg, grid, st, args, persist);
    if (cfg.G == 4 && cfg.V == 8) return launch_trade<4, 8>(cfg, grid, st, args, persist);
    if (cfg.G == 8 && cfg.V == 4) return launch_trade<8, 4>(cfg, grid, st, args, persist);
    if (cfg.G == 8 && cfg.V == 8) return launch_trade<8, 8>(cfg, grid, st, args, persist);
    if (cfg.G == 16 && cfg.V == 4) return launch_trade<16, 4>(cfg, grid, st, args, persist);
    if (cfg.G == 8 && cfg.V == 20) return launch_trade<8, 20>(cfg, grid, st, args, persist);
    if (cfg.G == 32 && cfg.V == 4) return launch_trade<32, 4>(cfg, grid, st, args, persist);
    if (cfg.G == 32 && cfg.V == 5) return launch_trade<32, 5>(cfg, grid, st, args, persist);
    return launch_trade<32, 8>(cfg, grid, st, args, persist);
}

struct HostScore {               // optional fused scoring request
    int k = 0;
    const int32_t *modules = nullptr;   // host, P x k
    const double *z_obs = nullptr;      // host, P (or null)
    int64_t *n_ge = nullptr;            // host out, P
    double *null_scores = nullptr;      // host out, P x n (or null)
    double *z_obs_out = nullptr;        // host out, P (or null)
};

enum { SL_LVL = 0, SL_PLAN, SL_AUX, SL_MOD, SL_Z, SL_SCORES, SL_CARRY, SL_OUT, SL_APPLIED, SL_WORK, SL_OFFS = 12, SL_PAIRS = 13, SL_RT = 16 };      // 10, 11, 14, 15: sdx_scan.cu

// Builds (once per uploaded matrix) the sparse layout of the trade rows: rows of at most SP_TS carriers become
// index lists, the others stay bit-packed.  Applicable when the rows are features, fit 32 words, the lists are the
// majority and the footprint shrinks at least by half.  EXPERIMENTAL: measured at 10.5 ms per 32 768 C1 nulls against
// 9.8 ms for the bit-packed kernel (profiles/r01_k_trade_ncu_summary.md), so it is only used with SDX_TRADE_SPARSE=1.
static int sparse_layout(sdx_ctx *c) {
    if (c->sp_state != 0) return SDX_OK;           // 1 = in use, -1 = not applicable
    c->sp_state = -1;
    const int R = c->R, RS = c->RS;
    if (c->rows_are_samples || RS > SP_G * SP_V || c->L > 65535 || R < 2) return SDX_OK;
    std::vector<int> size((size_t)R);
    SDX_CUDA(c, cudaMemcpyAsync(size.data(), c->d_rowsize, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<uint32_t> desc((size_t)R);
    int n_dense = 0, n_list = 0, n_sparse_rows = 0;
    for (int r = 0; r < R; ++r) {
        if (size[r] > SP_TS) desc[r] = sp_desc(true, size[r] > 0x7fff ? 0x7fff : size[r], n_dense++);
        else { desc[r] = sp_desc(false, size[r], n_list); n_list += size[r]; ++n_sparse_rows; }
    }
    const int list_words = (((n_list + 1) / 2) + 3) & ~3;
    const int fp_words = n_dense * RS + list_words;
    bool use = n_list <= 65535 && n_dense <= 65535 && 2 * n_sparse_rows >= R && 2 * fp_words <= R * RS;
    const char *env = getenv("SDX_TRADE_SPARSE");
    use = use && env != nullptr && atoi(env) != 0;
    if (!use) return SDX_OK;
    c->sp_n_dense = n_dense; c->sp_fp_words = fp_words > 4 ? fp_words : 4;
    SDX_CUDA(c, cudaMalloc(&c->d_sp_desc, sizeof(uint32_t) * R));
    SDX_CUDA(c, cudaMalloc(&c->d_sp_obs, sizeof(uint32_t) * c->sp_fp_words));
    SDX_CUDA(c, cudaMemsetAsync(c->d_sp_obs, 0, sizeof(uint32_t) * c->sp_fp_words, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(c->d_sp_desc, desc.data(), sizeof(uint32_t) * R, cudaMemcpyHostToDevice, c->stream));
    k_to_sparse<<<ceil_div(R, 8), 256, 0, c->stream>>>(c->d_trade_obs, R, RS, c->d_sp_desc, n_dense, c->d_sp_obs);
    c->total_launches++;
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    SDX_CUDA(c, cudaGetLastError());
    c->sp_state = 1;
    return SDX_OK;
}

// The common driver behind sdx_curveball and sdx_null_pvalue (stat = fixed).
static int run_nulls_impl(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                          uint32_t *out_rows, int64_t *applied, const HostScore *hs, uint32_t *save_states, bool use_pool);


// The burn-in pool of sdx_set_burn_in: pool_size chains of `burn` rounds from the observed matrix; their final states
// (in the trade kernel's own state layout) are the start states of the real chains.  Cached per (seed, T).
static int ensure_pool(sdx_ctx *c, uint64_t seed, int64_t T, size_t state_words) {
    if (c->burn <= 0) return SDX_OK;
    if (c->pool_valid && c->pool_seed == seed && c->pool_T == T && c->pool_state_words == state_words) return SDX_OK;
    c->pool_valid = false;
    SDX_TRY(sdx_reserve(c, c->d_pool, c->cap_pool, sizeof(uint32_t) * state_words * (size_t)c->pool_size));
    {   // pair set of the burn-in trades: the max(2, ceil(R / 10)) densest rows, size descending, index ascending
        const int R = c->R;
        std::vector<int> size((size_t)R);
        SDX_CUDA(c, cudaMemcpyAsync(size.data(), c->d_rowsize, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream));
        SDX_CUDA(c, cudaStreamSynchronize(c->stream));
        std::vector<uint16_t> top((size_t)R);
        for (int r = 0; r < R; ++r) top[r] = (uint16_t)r;
        std::stable_sort(top.begin(), top.end(), [&](uint16_t x, uint16_t y) { return size[x] > size[y]; });
        int K = (R + 9) / 10;
        if (K < 2) K = 2;
        if (K > R) K = R;
        SDX_TRY(sdx_reserve(c, c->d_top, c->cap_top, sizeof(uint16_t) * (size_t)R));
        SDX_CUDA(c, cudaMemcpyAsync(c->d_top, top.data(), sizeof(uint16_t) * K, cudaMemcpyHostToDevice, c->stream));
        SDX_CUDA(c, cudaStreamSynchronize(c->stream));
        c->top_K = K;
    }
    const uint64_t base_chain = SDX_POOL_BASE / (uint64_t)c->burn;
    int rc = run_nulls_impl(c, seed, base_chain * (uint64_t)c->burn, (int64_t)c->pool_size * c->burn, T, c->burn, nullptr, nullptr, nullptr,
                            c->d_pool, false);
    if (rc) return rc;
    c->pool_valid = true; c->pool_seed = seed; c->pool_T = T; c->pool_state_words = state_words;
    return SDX_OK;
}

int sdx_run_nulls(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                  uint32_t *out_rows, int64_t *applied, const HostScore *hs) {
    return run_nulls_impl(c, seed, null0, n_nulls, n_trades, chain_len, out_rows, applied, hs, nullptr, true);
}

static int run_nulls_impl(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                          uint32_t *out_rows, int64_t *applied, const HostScore *hs, uint32_t *save_states, bool use_pool) {
    SDX_REQUIRE(c, c->d_trade_obs, SDX_ERR_STATE, "sdx_upload_matrix has not been called");
    SDX_REQUIRE(c, n_nulls >= 0 && chain_len >= 1, SDX_ERR_INVALID, "n_nulls must be >= 0 and chain_len >= 1");
    SDX_REQUIRE(c, c->R >= 2, SDX_ERR_INVALID, "Curveball needs at least two trade rows");
    SDX_REQUIRE(c, null0 + (uint64_t)n_nulls < (1ull << 62) && (!use_pool || null0 + (uint64_t)n_nulls <= SDX_BURN_ID_MIN), SDX_ERR_INVALID,
                "null index out of range");
    const int64_t T = n_trades < 0 ? 5 * (int64_t)(c->S < c->F ? c->S : c->F) : n_trades;
    SDX_REQUIRE(c, T <= 60000, SDX_ERR_UNSUPPORTED, "more than 60000 trades per null are not supported");
    TradeConfig cfg;
    int rc = pick_config(c, &cfg);
    if (rc) return rc;
    if ((rc = sparse_layout(c))) return rc;
    const bool sparse = c->sp_state == 1;
    const size_t sp_scr = std::max(sizeof(uint32_t) * SP_GPW * 2 * (size_t)c->RS, sizeof(uint16_t) * 2 * (SP_TS + 1) * 32);
    const size_t sp_smem = sparse ? sizeof(uint32_t) * (size_t)c->sp_fp_words + sp_scr : 0;

    const int R = c->R, RS = c->RS;
    if (use_pool && c->burn > 0) {
        const int64_t Tp_ = n_trades < 0 ? 5 * (int64_t)(c->S < c->F ? c->S : c->F) : n_trades;
        if ((rc = ensure_pool(c, seed, Tp_, sparse ? (size_t)c->sp_fp_words : (size_t)R * RS))) return rc;
    }
    LaunchTimer timer(c);
    const PhiloxKeys keys = make_keys(seed);
    const size_t mat_words = sparse ? (size_t)c->sp_fp_words : (size_t)R * RS;      // state of one chain (carry buffer)
    const size_t out_per = (size_t)R * c->wprL;

    // ---- fused scoring set-up ------------------------------------------------
    ScoreArgs sc{};
    unsigned long long *d_nge = nullptr;
    double *d_scores = nullptr, *d_zobs = nullptr;
    if (hs) {
        SDX_REQUIRE(c, c->d_w, SDX_ERR_STATE, "sdx_upload_weights has not been called");
        SDX_REQUIRE(c, hs->k >= 1 && hs->k <= SDX_MAX_K, SDX_ERR_INVALID, "module size k must be in 1..SDX_MAX_K");
        SDX_REQUIRE(c, hs->modules && hs->n_ge, SDX_ERR_INVALID, "modules and n_ge are required");
        for (int64_t i = 0; i < (int64_t)c->P * hs->k; ++i)
            SDX_REQUIRE(c, hs->modules[i] >= -1 && hs->modules[i] < c->F, SDX_ERR_INVALID, "module feature index out of range");
        void *pm, *pz, *ps = nullptr;
        const size_t n_mod = (size_t)c->P * hs->k;
        // column cache of the fused scorer (global-memory variant with samples as rows): distinct module features + remapped modules
        std::vector<int32_t> ufeat, umod;
        if (c->rows_are_samples && !cfg.smem && !sparse) {
            std::vector<int32_t> slot((size_t)c->F, -1);
            umod.assign(n_mod, -1);
            for (size_t i = 0; i < n_mod; ++i) {
                const int32_t g = hs->modules[i];
                if (g < 0) continue;
                if (slot[g] < 0) { slot[g] = (int32_t)ufeat.size(); ufeat.push_back(g); }
                umod[i] = slot[g];
            }
            if (ufeat.empty() || sizeof(uint32_t) * ufeat.size() * c->wprS > 48 * 1024) { ufeat.clear(); umod.clear(); }
        }
        if ((rc = sdx_scratch(c, SL_MOD, sizeof(int32_t) * (2 * n_mod + ufeat.size()) + 16, &pm))) return rc;
        if ((rc = sdx_scratch(c, SL_Z, (sizeof(double) + sizeof(unsigned long long)) * c->P + 16, &pz))) return rc;
        d_zobs = (double *)pz;
        d_nge = (unsigned long long *)((char *)pz + sizeof(double) * c->P);
        SDX_CUDA(c, cudaMemcpyAsync(pm, hs->modules, sizeof(int32_t) * c->P * hs->k, cudaMemcpyHostToDevice, c->stream));
        SDX_CUDA(c, cudaMemsetAsync(d_nge, 0, sizeof(unsigned long long) * c->P, c->stream));
        if (hs->null_scores && n_nulls > 0) {
            if ((rc = sdx_scratch(c, SL_SCORES, sizeof(double) * c->P * n_nulls, &ps))) return rc;
            d_scores = (double *)ps;
        }
        sc.n_profiles = c->P; sc.k = hs->k; sc.modules = (const int32_t *)pm; sc.w = c->d_w; sc.mask = c->d_mask;
        sc.z_obs = d_zobs; sc.n_ge = d_nge; sc.null_scores = d_scores; sc.n_total = n_nulls;
        sc.S = c->S; sc.wprS = c->wprS; sc.rows_are_samples = c->rows_are_samples;
        if (!ufeat.empty()) {
            int32_t *d_umod = (int32_t *)pm + n_mod, *d_ufeat = d_umod + n_mod;
            SDX_CUDA(c, cudaMemcpyAsync(d_umod, umod.data(), sizeof(int32_t) * n_mod, cudaMemcpyHostToDevice, c->stream));
            SDX_CUDA(c, cudaMemcpyAsync(d_ufeat, ufeat.data(), sizeof(int32_t) * ufeat.size(), cudaMemcpyHostToDevice, c->stream));
            SDX_CUDA(c, cudaStreamSynchronize(c->stream));          // the host vectors go out of scope before the launch
            sc.umod = d_umod; sc.ufeat = d_ufeat; sc.n_ufeat = (int)ufeat.size();
        }
        if (hs->z_obs) {
            SDX_CUDA(c, cudaMemcpyAsync(d_zobs, hs->z_obs, sizeof(double) * c->P, cudaMemcpyHostToDevice, c->stream));
        } else {
            k_score_observed<<<ceil_div(c->P, 4), 128, 0, c->stream>>>(c->d_trade_obs, RS, sc, d_zobs);
            timer.count();
        }
    }

    float plan_ms = 0, trade_ms = 0;
    int n_trade_launches = 0;
    if (n_nulls > 0) {
        // chains overlapping [null0, null0 + n)
        const uint64_t null_hi = null0 + (uint64_t)n_nulls;
        const uint64_t chain_first = null0 / (uint64_t)chain_len;
        const uint64_t chain_last = (null_hi - 1) / (uint64_t)chain_len;
        const int64_t n_chains = (int64_t)(chain_last - chain_first + 1);
        // Nulls planned per launch.  A launch ends with a ragged tail (its slowest chains), so fewer, larger launches are
        // faster: 100 000 C1 nulls in chains of 16 run at 2.10 M/s with 32 768 per launch and 2.31 M/s with 131 072
        // (one launch).  Bounded by ~4 GB of plan scratch (18 bytes per trade) and, when the nulls themselves are
        // returned, by the staging buffer.
        int64_t NB = 131072;
        {
            const int64_t by_mem = ((int64_t)4 << 30) / (18 * (T > 0 ? T : 1));
            if (NB > by_mem) NB = by_mem;
            if (out_rows && NB > 32768) NB = 32768;
            if (NB < 1024) NB = 1024;
        }
        if (const char *e = getenv("SDX_NULLS_PER_LAUNCH")) { const long v = atol(e); if (v >= 256 && v <= (1 << 20)) NB = v; }
        const int seg_cap = (int)(chain_len < 256 ? chain_len : 256);
        int64_t chains_per_launch = NB / seg_cap;
        // Expected depth of the level schedule is a small multiple of the trades per row (2T/R; C0: 10 -> depth ~41).
        // dcap bounds the histogram of k_plan_sort; a null whose schedule is deeper runs in sequential order.
        int dcap = (int)(16 * T / R) + 64;
        if (dcap < 128) dcap = 128;
        if (dcap > SDX_HCAP) dcap = SDX_HCAP;
        const bool persist = !sparse && cfg.persist_slots > 0 && 2 * dcap <= SM_OFFS_CAP;
        {   // one CTA per chain and every chain of a launch does the same work: launch whole waves to avoid a ragged tail
            const int64_t slots = sparse ? (int64_t)16 * c->sm_count : (persist ? (int64_t)cfg.persist_slots * c->sm_count : (int64_t)cfg.slots);
            if (slots > 0 && chains_per_launch > slots) chains_per_launch = (chains_per_launch / slots) * slots;
        }
        if (chains_per_launch > n_chains) chains_per_launch = n_chains;
        const int64_t n_segs = (chain_len + seg_cap - 1) / seg_cap;
        const int64_t n_local_max = chains_per_launch * seg_cap;
        const int Tn = (int)(T > 0 ? T : 1);

        // level kernel geometry: last[] (R x PL_NT u16) in shared memory when it fits, else in global memory
        const int lv_threads = PL_NT;
        const size_t smem_cap = c->smem_optin > 4096 ? c->smem_optin - 2048 : 0;
        const size_t lv_tiles = sizeof(uint32_t) * PL_NT * (PL_TT + 1) + sizeof(uint16_t) * PL_NT * (PL_TT + 2);
        const size_t lt = dcap <= 255 ? 1 : 2;                          // bytes per entry of the last-level table (k_plan_levels<LT>)
        const size_t last_bytes = (((size_t)PL_NT * R * lt + 1) / 2) * 2;
        const bool last_in_smem = lv_tiles + last_bytes <= smem_cap;
        // EXPERIMENTAL level capacity (k_plan_levels, SDX_PLAN_CAP=<trades per level>): with 32 (one round of the trade
        // kernel's CTA per level) a C0 null needs 62 rounds instead of 79, but every round then ends in a barrier and the
        // measured step is slower (trade 11.1 vs 10.8 ms per launch, plan kernels 8.8 % -> 14.7 % of the step); 64 gives
        // trade 10.5 ms and the same plan cost.  Default: the earliest-level schedule (0).
        int level_cap = 0;
        if (T > 0) {
            if (const char *e = getenv("SDX_PLAN_CAP")) level_cap = atoi(e);
            if (level_cap < 0) level_cap = 0;
            const size_t cnt_bytes = sizeof(uint16_t) * PL_NT * ((size_t)dcap + 1);
            if (lv_tiles + (last_in_smem ? last_bytes : 0) + cnt_bytes > smem_cap || level_cap > 65535) level_cap = 0;
        }
        const size_t lv_smem = lv_tiles + (last_in_smem ? last_bytes : 0) + (level_cap > 0 ? sizeof(uint16_t) * PL_NT * ((size_t)dcap + 1) : 0);
        if (lv_smem > 48 * 1024) {
            if (lt == 1) SDX_CUDA(c, cudaFuncSetAttribute(k_plan_levels<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lv_smem));
            else SDX_CUDA(c, cudaFuncSetAttribute(k_plan_levels<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lv_smem));
        }

        void *p_lvl, *p_plan, *p_offs, *p_aux = nullptr, *p_carry = nullptr, *p_work = nullptr, *p_out = nullptr, *p_app = nullptr;
        const int offs_stride = (int)((2 * (T < dcap ? T : dcap) + 2 + 1) & ~1ll);
        const size_t sort_bins = sizeof(uint32_t) * ((size_t)dcap * SDX_NCLS + 2);
        const int cache_bins = sort_bins + 2 * (size_t)Tn + 4 <= 20 * 1024;                     // + one u16 per trade when two warps still fit
        const size_t sort_smem_warp = sort_bins + (cache_bins ? sizeof(uint32_t) * (size_t)((Tn + 1) / 2) : 0);
        const int sort_warps = sort_smem_warp * 8 <= 40 * 1024 ? 8 : (sort_smem_warp * 4 <= 40 * 1024 ? 4 : 2);
        if ((rc = sdx_scratch(c, SL_OFFS, sizeof(uint16_t) * (size_t)n_local_max * offs_stride + 64, &p_offs))) return rc;
        const size_t lvl_bytes = ((sizeof(uint16_t) * (size_t)n_local_max * Tn + 15) / 16) * 16;
        if ((rc = sdx_scratch(c, SL_LVL, lvl_bytes + sizeof(int) * n_local_max + 64, &p_lvl))) return rc;
        int *d_depth = (int *)((char *)p_lvl + lvl_bytes);
        if ((rc = sdx_scratch(c, SL_PLAN, (sparse ? sizeof(uint4) : sizeof(uint2)) * (size_t)n_local_max * Tn, &p_plan))) return rc;
        // round table of k_trade (k_plan_sort): at most one round per trade
        void *p_rt = nullptr;
        const int rt_stride = (Tn + 4) & ~3;
        const int per_round = (cfg.threads / 32) * (32 / cfg.G);
        int *d_nrounds = nullptr;
        if (!sparse) {
            if ((rc = sdx_scratch(c, SL_RT, sizeof(uint32_t) * (size_t)n_local_max * rt_stride + sizeof(int) * (size_t)n_local_max + 64, &p_rt))) return rc;
            d_nrounds = (int *)((uint32_t *)p_rt + (size_t)n_local_max * rt_stride);
        }
        void *p_pairs;
        const int Tp = (Tn + 3) & ~3;                       // pairs row length: 16-byte aligned rows
        if ((rc = sdx_scratch(c, SL_PAIRS, sizeof(uint32_t) * (size_t)n_local_max * Tp + 64, &p_pairs))) return rc;
        if (!last_in_smem) {
            const size_t nb = (size_t)ceil_div(n_local_max, lv_threads) * lv_threads;
            if ((rc = sdx_scratch(c, SL_AUX, lt * nb * R, &p_aux))) return rc;
        }
        const bool need_carry = n_segs > 1;
        SDX_REQUIRE(c, !save_states || n_segs == 1, SDX_ERR_INVALID, "burn-in chains must fit one segment");
        if (need_carry && (rc = sdx_scratch(c, SL_CARRY, sizeof(uint32_t) * mat_words * chains_per_launch, &p_carry))) return rc;
        if (out_rows && (rc = sdx_scratch(c, SL_OUT, sizeof(uint32_t) * out_per * n_local_max, &p_out))) return rc;
        if (applied) {
            if ((rc = sdx_scratch(c, SL_APPLIED, sizeof(unsigned long long) * n_nulls, &p_app))) return rc;
            SDX_CUDA(c, cudaMemsetAsync(p_app, 0, sizeof(unsigned long long) * n_nulls, c->stream));
        }
        int grid_cap = cfg.grid_cap;
        if (!cfg.smem) {
            if (grid_cap > chains_per_launch) grid_cap = (int)chains_per_launch;
            if ((rc = sdx_scratch(c, SL_WORK, sizeof(uint32_t) * mat_words * grid_cap, &p_work))) return rc;
        }

        for (int64_t cb = 0; cb < n_chains; cb += chains_per_launch) {
            const int64_t nch = (n_chains - cb) < chains_per_launch ? (n_chains - cb) : chains_per_launch;
            for (int64_t sg = 0; sg < n_segs; ++sg) {
                NullIndexing ix;
                ix.chain0 = chain_first + (uint64_t)cb; ix.chain_len = chain_len; ix.seg_off = sg * seg_cap;
                const int64_t left = chain_len - sg * seg_cap;
                ix.seg_len = (int)(left < seg_cap ? left : seg_cap);
                ix.null_lo = null0; ix.null_hi = null_hi;
                const int n_local = (int)(nch * ix.seg_len);
                cudaEventRecord(c->ev[2], c->stream);
                if (T > 0) {
                    if (lt == 1)
                        k_plan_levels<uint8_t><<<ceil_div(n_local, lv_threads), lv_threads, lv_smem, c->stream>>>(
                            keys, ix, n_local, R, (int)T, Tp, (uint32_t *)p_pairs, (uint16_t *)p_lvl, d_depth, (uint8_t *)p_aux, level_cap, dcap, c->d_top, c->top_K);
                    else
                        k_plan_levels<uint16_t><<<ceil_div(n_local, lv_threads), lv_threads, lv_smem, c->stream>>>(
                            keys, ix, n_local, R, (int)T, Tp, (uint32_t *)p_pairs, (uint16_t *)p_lvl, d_depth, (uint16_t *)p_aux, level_cap, dcap, c->d_top, c->top_K);
                    k_plan_sort<<<ceil_div(n_local, sort_warps), sort_warps * 32, sort_smem_warp * sort_warps, c->stream>>>(
                                                                          n_local, R, (int)T, Tp, dcap, (const uint32_t *)p_pairs,
                                                                          (const uint16_t *)p_lvl, d_depth, c->d_rowsize, sparse ? c->d_sp_desc : nullptr,
                                                                          sparse ? SP_TS : -1, p_plan,
                                                                          (uint16_t *)p_offs, offs_stride, (uint32_t *)p_rt, rt_stride, d_nrounds, per_round, cache_bins);
                    timer.count(2);
                }
                cudaEventRecord(c->ev[3], c->stream);
                TradeArgs ta{};
                ta.keys = keys; ta.ix = ix; ta.observed = c->d_trade_obs; ta.work = (uint32_t *)p_work;
                ta.carry = (uint32_t *)p_carry; ta.plan = p_plan; ta.n_units = (int)nch;
                ta.offs = (const uint16_t *)p_offs; ta.depth = d_depth; ta.offs_stride = offs_stride; ta.dcap = dcap;
                ta.rt = (const uint32_t *)p_rt; ta.nrounds = d_nrounds; ta.rt_stride = rt_stride;
                ta.R = R; ta.RS = RS; ta.T = (int)T; ta.wprL = c->wprL;
                ta.save_carry = need_carry && sg + 1 < n_segs;
                if (save_states) { ta.carry = save_states + (size_t)cb * mat_words; ta.save_carry = 1; }
                ta.pool = (use_pool && c->burn > 0) ? c->d_pool : nullptr;
                ta.pool_size = c->pool_size;
                ta.out_rows = (uint32_t *)p_out; ta.applied = (unsigned long long *)p_app; ta.sc = sc;
                const int grid = (int)(nch < grid_cap ? nch : grid_cap);
                cudaError_t e;
                if (sparse) {
                    ta.desc = c->d_sp_desc; ta.obs_state = c->d_sp_obs; ta.n_dense = c->sp_n_dense; ta.fp_words = c->sp_fp_words;
                    e = cudaFuncSetAttribute(k_trade_sparse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp_smem);
                    if (e == cudaSuccess) {
                        k_trade_sparse<<<(unsigned)nch, 32, sp_smem, c->stream>>>(ta);
                        e = cudaGetLastError();
                    }
                } else if (persist) {
                    e = dispatch_trade(cfg, (int)(nch < c->sm_count ? nch : c->sm_count), c->stream, ta, true);
                } else {
                    e = dispatch_trade(cfg, grid, c->stream, ta, false);
                }
                if (e != cudaSuccess) return sdx_fail(c, SDX_ERR_CUDA, std::string("trade kernel launch failed: ") + cudaGetErrorString(e));
                timer.count();
                ++n_trade_launches;
                cudaEventRecord(c->ev[4], c->stream);
                if (out_rows) {
                    // staged rows are indexed by the launch-local null index; copy out the requested ids
                    const bool contiguous = (n_segs == 1);
                    const int64_t n_copies = contiguous ? 1 : nch;
                    for (int64_t ci = 0; ci < n_copies; ++ci) {
                        const int64_t l0 = contiguous ? 0 : ci * ix.seg_len;
                        const int64_t l1 = contiguous ? n_local : l0 + ix.seg_len;
                        uint64_t lo = ix.id(l0), hi = ix.id(l1 - 1) + 1;
                        const uint64_t first = lo;
                        if (lo < null0) lo = null0;
                        if (hi > null_hi) hi = null_hi;
                        if (hi > lo)
                            SDX_CUDA(c, cudaMemcpyAsync(out_rows + (size_t)(lo - null0) * out_per,
                                                        (uint32_t *)p_out + (size_t)(l0 + (int64_t)(lo - first)) * out_per,
                                                        sizeof(uint32_t) * out_per * (hi - lo), cudaMemcpyDefault, c->stream));
                    }
                }
                SDX_CUDA(c, cudaStreamSynchronize(c->stream));
                float a_ms = 0, b_ms = 0;
                cudaEventElapsedTime(&a_ms, c->ev[2], c->ev[3]);
                cudaEventElapsedTime(&b_ms, c->ev[3], c->ev[4]);
                plan_ms += a_ms; trade_ms += b_ms;
            }
        }
        if (applied) {
            std::vector<unsigned long long> tmp((size_t)n_nulls);
            SDX_CUDA(c, cudaMemcpyAsync(tmp.data(), p_app, sizeof(unsigned long long) * n_nulls, cudaMemcpyDeviceToHost, c->stream));
            SDX_CUDA(c, cudaStreamSynchronize(c->stream));
            for (int64_t i = 0; i < n_nulls; ++i) applied[i] = (int64_t)tmp[i];
        }
    }

    if (hs) {
        std::vector<unsigned long long> ge((size_t)c->P);
        SDX_CUDA(c, cudaMemcpyAsync(ge.data(), d_nge, sizeof(unsigned long long) * c->P, cudaMemcpyDeviceToHost, c->stream));
        if (hs->z_obs_out) SDX_CUDA(c, cudaMemcpyAsync(hs->z_obs_out, d_zobs, sizeof(double) * c->P, cudaMemcpyDeviceToHost, c->stream));
        if (hs->null_scores && n_nulls > 0)
            SDX_CUDA(c, cudaMemcpyAsync(hs->null_scores, d_scores, sizeof(double) * c->P * n_nulls, cudaMemcpyDeviceToHost, c->stream));
        SDX_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int p = 0; p < c->P; ++p) hs->n_ge[p] = (int64_t)ge[p];
    }
    rc = timer.finish();
    c->timing.plan_ms = plan_ms; c->timing.trade_ms = trade_ms; c->timing.trade_launches = n_trade_launches;
    return rc;
}

// nulls null0 .. null0+n-1 as dense trade rows (n x R x wprL words) into DEVICE memory (used by the max_k statistic)
int sdx_generate_to_device(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n, int64_t n_trades, int64_t chain_len,
                           uint32_t *d_out) {
    return sdx_run_nulls(ctx, seed, null0, n, n_trades, chain_len, d_out, nullptr, nullptr);
}

extern "C" int sdx_curveball(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                             int64_t chain_len, uint32_t *out_trade_rows, int64_t *applied) {
    if (!ctx) return SDX_ERR_INVALID;
    return sdx_run_nulls(ctx, seed, null0, n_nulls, n_trades, chain_len, out_trade_rows, applied, nullptr);
}

extern "C" int sdx_curveball_replay(sdx_ctx *c, const uint32_t *start, const int32_t *a, const int32_t *b,
                                    const uint32_t *to_a, int64_t T, uint32_t *out, int64_t *applied) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_trade_obs, SDX_ERR_STATE, "sdx_upload_matrix has not been called");
    SDX_REQUIRE(c, T >= 0 && (T == 0 || (a && b && to_a)), SDX_ERR_INVALID, "schedule arrays missing");
    const int R = c->R, RS = c->RS, wprL = c->wprL;
    for (int64_t t = 0; t < T; ++t)
        SDX_REQUIRE(c, a[t] >= 0 && a[t] < R && b[t] >= 0 && b[t] < R && a[t] != b[t], SDX_ERR_INVALID, "schedule row index out of range");
    LaunchTimer timer(c);
    void *p_rows, *p_sched;
    int rc;
    const size_t mat = sizeof(uint32_t) * (size_t)R * RS;
    if ((rc = sdx_scratch(c, 6, mat + 64, &p_rows))) return rc;
    const size_t sched_bytes = (size_t)T * (8 + 4 * (size_t)wprL) + 64;
    if ((rc = sdx_scratch(c, 1, sched_bytes, &p_sched))) return rc;
    uint32_t *d_rows = (uint32_t *)p_rows;
    unsigned long long *d_ap = (unsigned long long *)((char *)p_rows + mat);
    if (start) {
        SDX_CUDA(c, cudaMemsetAsync(d_rows, 0, mat, c->stream));
        SDX_CUDA(c, cudaMemcpy2DAsync(d_rows, (size_t)RS * 4, start, (size_t)wprL * 4, (size_t)wprL * 4, R, cudaMemcpyHostToDevice, c->stream));
    } else {
        SDX_CUDA(c, cudaMemcpyAsync(d_rows, c->d_trade_obs, mat, cudaMemcpyDeviceToDevice, c->stream));
    }
    int32_t *d_a = (int32_t *)p_sched, *d_b = d_a + T;
    uint32_t *d_m = (uint32_t *)(d_b + T);
    if (T > 0) {
        SDX_CUDA(c, cudaMemcpyAsync(d_a, a, sizeof(int32_t) * T, cudaMemcpyHostToDevice, c->stream));
        SDX_CUDA(c, cudaMemcpyAsync(d_b, b, sizeof(int32_t) * T, cudaMemcpyHostToDevice, c->stream));
        SDX_CUDA(c, cudaMemcpyAsync(d_m, to_a, sizeof(uint32_t) * T * wprL, cudaMemcpyHostToDevice, c->stream));
    }
    k_replay<<<1, 32, 0, c->stream>>>(d_rows, RS, wprL, d_a, d_b, d_m, T, d_ap);
    timer.count();
    unsigned long long ap = 0;
    if (out) SDX_CUDA(c, cudaMemcpy2DAsync(out, (size_t)wprL * 4, d_rows, (size_t)RS * 4, (size_t)wprL * 4, R, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(&ap, d_ap, sizeof ap, cudaMemcpyDeviceToHost, c->stream));
    rc = timer.finish();
    if (applied) *applied = (int64_t)ap;
    return rc;
}

int sdx_null_pvalue_maxk(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                         const int32_t *modules, int k, const double *z_obs, int64_t *n_ge, double *null_scores,
                         double *z_obs_out);      // sdx_maxk.cu

static int sdx_null_pvalue_fixed(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                                 int64_t chain_len, const int32_t *modules, int k, const double *z_obs,
                                 int64_t *n_ge, double *null_scores, double *z_obs_out) {
    HostScore hs;
    hs.k = k; hs.modules = modules; hs.z_obs = z_obs; hs.n_ge = n_ge; hs.null_scores = null_scores; hs.z_obs_out = z_obs_out;
    return sdx_run_nulls(ctx, seed, null0, n_nulls, n_trades, chain_len, nullptr, nullptr, &hs);
}

extern "C" int sdx_null_pvalue(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades,
                               int64_t chain_len, int stat, const int32_t *modules, int k, const double *z_obs,
                               int64_t *n_ge, double *null_scores, double *z_obs_out) {
    if (!ctx) return SDX_ERR_INVALID;
    if (stat == SDX_STAT_FIXED)
        return sdx_null_pvalue_fixed(ctx, seed, null0, n_nulls, n_trades, chain_len, modules, k, z_obs, n_ge, null_scores, z_obs_out);
    if (stat == SDX_STAT_MAXK)
        return sdx_null_pvalue_maxk(ctx, seed, null0, n_nulls, n_trades, chain_len, modules, k, z_obs, n_ge, null_scores, z_obs_out);
    return sdx_fail(ctx, SDX_ERR_INVALID, "unknown null statistic");
}
