// sdx_plan.cuh — the trade plan of a batch of nulls: pairs (PAIR stream), level schedule, counting sort by (level, size class).
// Included by sdx_curveball.cu only.
#pragma once
#include "sdx_trade.cuh"

// ---------------------------------------------------------------------------
struct NullIndexing {          // local null index (within a launch) -> global null id
    uint64_t chain0;           // first chain of this launch
    int64_t chain_len;
    int64_t seg_off;           // chain-relative offset of the segment
    int seg_len;               // nulls per chain in this launch
    uint64_t null_lo, null_hi; // requested range [null_lo, null_hi)
    int pair_group;            // sdx_set_pair_group: 1, or 32 = aligned blocks of 32 chains share the PAIR stream of their first chain
    int chain_stride;          // 1; pair_group when the index runs over the first chains of the blocks only (the plan kernels: one plan per block)
    __host__ __device__ uint64_t id(int64_t local) const {
        return (chain0 + (uint64_t)(local / seg_len) * (uint64_t)chain_stride) * (uint64_t)chain_len + (uint64_t)seg_off + (uint64_t)(local % seg_len);
    }
    // stream id the row pairs of null `id` are drawn from: round j of chain c uses the pairs of round j of
    // chain c - c % pair_group, so a block of pair_group chains walks the same schedule of row pairs
    __host__ __device__ uint64_t pair_id(uint64_t id) const {
        if (pair_group <= 1) return id;
        const uint64_t c = id / (uint64_t)chain_len, j = id - c * (uint64_t)chain_len;
        return (c - c % (uint64_t)pair_group) * (uint64_t)chain_len + j;
    }
};

// ---------------------------------------------------------------------------
// plan: levels
// ---------------------------------------------------------------------------
// One THREAD per null draws and walks its trades in order; pairs and levels go out through shared-memory
// tiles of PL_TT trades x PL_NT nulls so that every global access is a coalesced row segment.
#define PL_NT 64
#define PL_TT 32
// cap > 0: at most `cap` trades per level (online first fit: the earliest level after both predecessors that still has
// room).  With cap = the number of trades one CTA of the trade kernel runs at a time, every level is exactly one
// round of that CTA, and the rounds per null drop from 79 (earliest-level schedule, half-empty second rounds) to 62
// at C0 (ideal: T / cap = 60; tests/tools/exp_levels.py).  Any assignment that puts a trade after its two
// predecessors is a valid schedule, so the nulls do not change.  A schedule that would need dcap levels or more is
// reported as depth dcap (the trade kernel then runs that null in sequential order).
// LT: type of the per-row "level of the last trade" table — u8 when dcap <= 255 (half the shared memory, twice the resident
// CTAs), else u16.  Levels are clamped to dcap (a null that reaches it runs in sequential order anyway).
template <typename LT>
__global__ void __launch_bounds__(PL_NT) k_plan_levels(const __grid_constant__ PhiloxKeys keys, NullIndexing ix, int n_local, int R, int T,
                                                       int Tp, uint32_t *__restrict__ pairs, uint16_t *__restrict__ lvl,
                                                       int *__restrict__ depth, LT *__restrict__ last_global, int cap, int dcap,
                                                       const uint16_t *__restrict__ top, int top_K) {
    extern __shared__ __align__(16) uint16_t plan_smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n0 = blockIdx.x * PL_NT, n = n0 + tid;
    // layout: tile_p[PL_NT][PL_TT + 1] u32 | tile_l[PL_NT][PL_TT + 2] u16 | last[R][PL_NT] LT (unless in global memory)
    uint32_t *tile_p = reinterpret_cast<uint32_t *>(plan_smem);
    uint16_t *tile_l = reinterpret_cast<uint16_t *>(tile_p + PL_NT * (PL_TT + 1));
    uint16_t *after_tiles = tile_l + PL_NT * (PL_TT + 2);
    LT *last = last_global ? last_global + (size_t)blockIdx.x * PL_NT * R + tid : reinterpret_cast<LT *>(after_tiles) + tid;
    // fill[level][tid]: trades already assigned to a level (only with cap > 0); it follows last[] when that is in shared memory
    uint16_t *fill = after_tiles + (last_global ? 0 : ((size_t)PL_NT * R * sizeof(LT) + 1) / 2) + tid;
    for (int r = 0; r < R; ++r) last[(size_t)r * PL_NT] = 0;
    if (cap > 0)
        for (int l = 0; l <= dcap; ++l) fill[(size_t)l * PL_NT] = 0;
    int d = 0, lo = 1;                                             // lo: lowest level that still has room
    // The thread draws the row pair of each of its trades itself (PAIR stream): the Philox arithmetic is independent of the
    // shared-memory chain of the level walk and fills its latency.  Pairs and levels leave through the tiles, coalesced.
    const uint64_t null_id = ix.id(n < n_local ? n : 0);
    const bool burn = top_K >= 2 && null_id >= SDX_BURN_ID_MIN;     // burn-in rounds of the pool (the whole launch or none of it)
    const uint64_t pair_id = ix.pair_id(null_id);
    for (int t0 = 0; t0 < T; t0 += PL_TT) {
        const int cnt = T - t0 < PL_TT ? T - t0 : PL_TT;
        for (int c = 0; c < cnt; ++c) {
            uint32_t a, b;
            if (burn) trade_pair_burn(keys, null_id, (uint32_t)(t0 + c), (uint32_t)R, top, (uint32_t)top_K, a, b);
            else trade_pair(keys, pair_id, (uint32_t)(t0 + c), (uint32_t)R, a, b);
            tile_p[tid * (PL_TT + 1) + c] = a | (b << 16);
            const int la = last[(size_t)a * PL_NT], lb = last[(size_t)b * PL_NT];
            int l = (la > lb ? la : lb) + 1;
            if (cap > 0) {
                if (l < lo) l = lo;
                while (l < dcap && fill[(size_t)l * PL_NT] >= cap) ++l;
                if (l > dcap) l = dcap;
                ++fill[(size_t)l * PL_NT];
                while (lo < dcap && fill[(size_t)lo * PL_NT] >= cap) ++lo;
            }
            if (l > dcap) l = dcap;
            last[(size_t)a * PL_NT] = (LT)l;
            last[(size_t)b * PL_NT] = (LT)l;
            tile_l[tid * (PL_TT + 2) + c] = (uint16_t)l;
            d = l > d ? l : d;
        }
        __syncthreads();
        for (int i = wid; i < PL_NT; i += PL_NT / 32)
            if (n0 + i < n_local && t0 + lane < T) {
                lvl[(size_t)(n0 + i) * T + t0 + lane] = tile_l[i * (PL_TT + 2) + lane];
                pairs[(size_t)(n0 + i) * Tp + t0 + lane] = tile_p[i * (PL_TT + 1) + lane];
            }
        __syncthreads();
    }
    if (n < n_local) depth[n] = d;
}

// One WARP per null — the variant for few nulls (the burn-in pool: a few thousand rounds of 5 * R_b trades), where the
// kernel above is pure latency (one thread draws and walks ~4 000 trades: 1.3 ms however few nulls there are).  The lanes
// draw the pairs of 32 trades at once, lane 0 walks them through the warp's last[] table (a single thread: program
// order, no barrier inside the walk), and pairs and levels leave coalesced.  Same schedule as k_plan_levels with cap = 0.
// Shared memory per warp: last[R] (LT, padded to 16 bytes) | 32 pairs u32 | 32 levels u16.
#define PLW_WARPS 4
template <typename LT>
__global__ void __launch_bounds__(PLW_WARPS * 32) k_plan_levels_warp(const __grid_constant__ PhiloxKeys keys, NullIndexing ix, int n_local, int R, int T,
                                                                     int Tp, uint32_t *__restrict__ pairs, uint16_t *__restrict__ lvl,
                                                                     int *__restrict__ depth, int dcap, const uint16_t *__restrict__ top, int top_K) {
    extern __shared__ __align__(16) uint16_t plan_smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = blockIdx.x * PLW_WARPS + wid;
    if (n >= n_local) return;                                      // whole warps
    const size_t last_bytes = (((size_t)R * sizeof(LT) + 15) / 16) * 16;
    unsigned char *base = reinterpret_cast<unsigned char *>(plan_smem) + (size_t)wid * (last_bytes + 32 * 4 + 32 * 2);
    LT *last = reinterpret_cast<LT *>(base);
    volatile uint32_t *st_p = reinterpret_cast<uint32_t *>(base + last_bytes);
    volatile uint16_t *st_l = reinterpret_cast<uint16_t *>(base + last_bytes + 32 * 4);
    for (int r = lane; r < R; r += 32) last[r] = 0;
    const uint64_t null_id = ix.id(n);
    const bool burn = top_K >= 2 && null_id >= SDX_BURN_ID_MIN;
    const uint64_t pair_id = ix.pair_id(null_id);
    int d = 0;
    __syncwarp();
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane, cnt = T - t0 < 32 ? T - t0 : 32;
        uint32_t e = 0;
        if (t < T) {
            uint32_t a, b;
            if (burn) trade_pair_burn(keys, null_id, (uint32_t)t, (uint32_t)R, top, (uint32_t)top_K, a, b);
            else trade_pair(keys, pair_id, (uint32_t)t, (uint32_t)R, a, b);
            e = a | (b << 16);
        }
        st_p[lane] = e;
        __syncwarp();
        if (lane == 0) {
#pragma unroll 4
            for (int i = 0; i < cnt; ++i) {
                const uint32_t q = st_p[i];
                const uint32_t a = q & 0xffffu, b = q >> 16;
                const int la = last[a], lb = last[b];
                int l = (la > lb ? la : lb) + 1;
                if (l > dcap) l = dcap;
                last[a] = (LT)l;
                last[b] = (LT)l;
                st_l[i] = (uint16_t)l;
                d = l > d ? l : d;
            }
        }
        __syncwarp();
        if (t < T) {
            lvl[(size_t)n * T + t] = st_l[lane];
            pairs[(size_t)n * Tp + t] = e;
        }
        __syncwarp();
    }
    if (lane == 0) depth[n] = d;
}

// ---------------------------------------------------------------------------
// plan: counting sort by level (one warp per null)
// ---------------------------------------------------------------------------
#define SDX_HCAP 1024          // upper bound of the deepest level schedule the plan kernels sort (see dcap below)
#define SDX_NCLS 4             // trade size classes inside a level (largest first)

// Row sizes never change (Curveball preserves them), so min(|a|, |b|) bounds the number of picks of a
// trade and |a| + |b| bounds its symmetric difference.  Inside a level the plan lists the large
// trades first and groups similar sizes, so the groups of one warp run loops of similar length and
// the register fast path of trade_rows (n <= 64) is taken by whole warps.
__device__ __forceinline__ int trade_class(int sa, int sb) {
    const int kb = sa < sb ? sa : sb;
    if (sa + sb > 64) return 0;
    return kb > 8 ? 1 : (kb > 3 ? 2 : 3);
}

// Round table (k_trade): one u32 per round = the `ng` trades a CTA of the trade kernel runs at a time:
//   bits 0-15 first plan entry, bits 16-30 number of trades, bit 31 last round of its level (barrier after it).
// The kernel then needs no level bookkeeping at all.
__global__ void k_plan_sort(int n_local, int R, int T, int Tp, int dcap, const uint32_t *__restrict__ pairs,
                            const uint16_t *__restrict__ lvl, const int *__restrict__ depth, const int *__restrict__ rowsize,
                            uint2 *__restrict__ plan, uint32_t *__restrict__ rt, int rt_stride, int *__restrict__ nrounds, int ng,
                            int cache_bins) {
    extern __shared__ uint32_t hist_smem[];        // per warp: dcap * SDX_NCLS + 2 bins + T u16 trade bins
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = blockIdx.x * (blockDim.x >> 5) + wid;
    if (n >= n_local) return;
    const int D = depth[n];
    // per warp: dcap * NCLS + 2 bins, then the bin of every trade (u16) so that the second pass does not look it up again
    // (cache_bins = 0 when T is too large for that: the second pass then recomputes the bin)
    const size_t per_warp = (size_t)(dcap * SDX_NCLS + 2) + (cache_bins ? (size_t)((T + 1) / 2) : 0);
    uint32_t *hist = hist_smem + (size_t)wid * per_warp;
    uint16_t *bin = reinterpret_cast<uint16_t *>(hist + (dcap * SDX_NCLS + 2));
    const uint16_t *lv = lvl + (size_t)n * T;
    // plan entry: (a | b << 16, t)
    uint2 *out2 = plan + (size_t)n * T;
    auto emit = [&](uint32_t pos, uint32_t e, uint32_t t) { out2[pos] = make_uint2(e, t); };
    const uint32_t *pr = pairs + (size_t)n * Tp;
    uint32_t *rtn = rt ? rt + (size_t)n * rt_stride : nullptr;
    if (D >= dcap) {            // very deep schedule: fall back to the sequential order (always valid)
        for (int t = lane; t < T; t += 32) {
            emit((uint32_t)t, pr[t], (uint32_t)t);
            if (rtn) rtn[t] = (uint32_t)t | (1u << 16) | (1u << 31);       // one trade per round, a barrier after each
        }
        if (rtn && lane == 0) nrounds[n] = T;
        return;                 // k_trade runs such a null one trade per level, in order
    }
    const int NB = D * SDX_NCLS;                       // bin = (level - 1) * NCLS + class
    for (int i = lane; i <= NB; i += 32) hist[i] = 0;
    __syncwarp();
    for (int t = lane; t < T; t += 32) {
        const uint32_t e = pr[t];
        const int b = (lv[t] - 1) * SDX_NCLS + trade_class(rowsize[e & 0xffffu], rowsize[e >> 16]);
        if (cache_bins) bin[t] = (uint16_t)b;
        atomicAdd(&hist[b], 1u);
    }
    __syncwarp();
    // exclusive scan of the bins -> start offsets
    uint32_t carry = 0;
    for (int base = 0; base < NB; base += 32) {
        const int i = base + lane;
        const uint32_t c = (i < NB) ? hist[i] : 0u;
        uint32_t inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        if (i < NB) hist[i] = carry + inc - c;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
    if (rtn) {
        // rounds of level l: ceil(n_l / ng); their positions in the table = prefix sum over the levels
        uint32_t rcarry = 0;
        for (int base = 0; base < D; base += 32) {
            const int l = base + lane;
            uint32_t s0 = 0, nl = 0;
            if (l < D) {
                s0 = hist[l * SDX_NCLS];
                nl = (l + 1 < D ? hist[(l + 1) * SDX_NCLS] : (uint32_t)T) - s0;
            }
            const uint32_t nr = (nl + ng - 1) / ng;
            uint32_t inc = nr;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            uint32_t pos = rcarry + inc - nr;
            for (uint32_t q = 0; q < nr; ++q) {
                const uint32_t left = nl - q * ng;
                rtn[pos + q] = (s0 + q * ng) | ((left < (uint32_t)ng ? left : (uint32_t)ng) << 16) | (q + 1 == nr ? 1u << 31 : 0u);
            }
            rcarry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (lane == 0) nrounds[n] = (int)rcarry;
    }
    __syncwarp();
    for (int t = lane; t < T; t += 32) {
        const uint32_t e = pr[t];
        const int b = cache_bins ? (int)bin[t] : (lv[t] - 1) * SDX_NCLS + trade_class(rowsize[e & 0xffffu], rowsize[e >> 16]);
        const uint32_t pos = atomicAdd(&hist[b], 1u);
        emit(pos, e, (uint32_t)t);
    }
}

