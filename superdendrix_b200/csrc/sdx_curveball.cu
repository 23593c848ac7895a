// sdx_curveball.cu — Curveball null generation (+ fused fixed-set scoring).
//
// Replaces, behind the C ABI of include/sdx.h:
//   src/curveball.py:17-40            curve_ball
//   src/generate_null_matrices.py:71-74  the null loop over one mutable presence list
//   src/superdendrix.py:166-197       the null loop / exceedance count (stat = fixed)
//
// Pipeline per batch of nulls (all on the ctx stream):
//   k_plan_levels   one THREAD per null: draws the trade pairs (PAIR stream) in
//                   order and assigns each trade the level
//                   1 + max(level of the last trade that touched row a, row b).
//                   Trades of one level touch pairwise disjoint rows, and every
//                   earlier trade on those rows is in a lower level, so running
//                   level by level is identical to the sequential order.
//   k_plan_sort     one WARP per null: counting sort of the trades by level into
//                   the plan (a | b << 16, t) plus the level offsets.
//   k_trade         one CTA per chain segment: matrix resident in shared memory
//                   (or in L2/global when it does not fit), groups of G lanes
//                   execute the trades of a level concurrently (plan entries
//                   prefetched one round ahead), __syncthreads between levels;
//                   then the fused scorer and the outputs.
#include <stdlib.h>

#include <algorithm>

#include "sdx_sparse.cuh"

// ---------------------------------------------------------------------------
struct NullIndexing {          // local null index (within a launch) -> global null id
    uint64_t chain0;           // first chain of this launch
    int64_t chain_len;
    int64_t seg_off;           // chain-relative offset of the segment
    int seg_len;               // nulls per chain in this launch
    uint64_t null_lo, null_hi; // requested range [null_lo, null_hi)
    __host__ __device__ uint64_t id(int64_t local) const {
        return (chain0 + (uint64_t)(local / seg_len)) * (uint64_t)chain_len + (uint64_t)seg_off + (uint64_t)(local % seg_len);
    }
};

// ---------------------------------------------------------------------------
// plan: levels
// ---------------------------------------------------------------------------
// pairs of every trade of every null of the launch, one thread per (null, trade): pairs[n][t] = a | b << 16
__global__ void k_plan_pairs(const __grid_constant__ PhiloxKeys keys, NullIndexing ix, int n_local, int R, int T, int Tp,
                             uint32_t *__restrict__ pairs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n = (int)(i / Tp), t = (int)(i - (int64_t)n * Tp);
    if (n >= n_local) return;
    uint32_t a = 0, b = 0;
    if (t < T) trade_pair(keys, ix.id(n), (uint32_t)t, (uint32_t)R, a, b);
    pairs[i] = a | (b << 16);
}

// One THREAD per null walks its trades in order; pairs come in and levels go out through shared-memory
// tiles of PL_TT trades x PL_NT nulls so that every global access is a coalesced row segment.
#define PL_NT 64
#define PL_TT 32
__global__ void __launch_bounds__(PL_NT) k_plan_levels(int n_local, int R, int T, int Tp, const uint32_t *__restrict__ pairs,
                                                       uint16_t *__restrict__ lvl, int *__restrict__ depth,
                                                       uint16_t *__restrict__ last_global) {
    extern __shared__ __align__(16) uint16_t plan_smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n0 = blockIdx.x * PL_NT, n = n0 + tid;
    // layout: tile_p[PL_NT][PL_TT + 1] u32 | tile_l[PL_NT][PL_TT + 2] u16 | last[R][PL_NT] u16 (unless in global memory)
    uint32_t *tile_p = reinterpret_cast<uint32_t *>(plan_smem);
    uint16_t *tile_l = reinterpret_cast<uint16_t *>(tile_p + PL_NT * (PL_TT + 1));
    uint16_t *last = last_global ? last_global + (size_t)blockIdx.x * PL_NT * R + tid : tile_l + PL_NT * (PL_TT + 2) + tid;
    for (int r = 0; r < R; ++r) last[(size_t)r * PL_NT] = 0;
    int d = 0;
    constexpr int PER = PL_NT / (PL_NT / 32);                      // null rows each warp loads per tile
    uint32_t nx[PER];                                              // next tile, in flight while the current one is processed
    auto fetch = [&](int t0) {
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            const int i = wid + q * (PL_NT / 32);
            nx[q] = (n0 + i < n_local && t0 + lane < Tp) ? pairs[(size_t)(n0 + i) * Tp + t0 + lane] : 0u;      // coalesced 128-byte segments
        }
    };
    fetch(0);
    for (int t0 = 0; t0 < T; t0 += PL_TT) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < PER; ++q) tile_p[(wid + q * (PL_NT / 32)) * (PL_TT + 1) + lane] = nx[q];
        __syncthreads();
        if (t0 + PL_TT < T) fetch(t0 + PL_TT);
        const int cnt = T - t0 < PL_TT ? T - t0 : PL_TT;
        for (int c = 0; c < cnt; ++c) {
            const uint32_t e = tile_p[tid * (PL_TT + 1) + c];
            const uint32_t a = e & 0xffffu, b = e >> 16;
            const int la = last[(size_t)a * PL_NT], lb = last[(size_t)b * PL_NT];
            const int l = (la > lb ? la : lb) + 1;
            last[(size_t)a * PL_NT] = (uint16_t)l;
            last[(size_t)b * PL_NT] = (uint16_t)l;
            tile_l[tid * (PL_TT + 2) + c] = (uint16_t)l;
            d = l > d ? l : d;
        }
        __syncthreads();
        for (int i = wid; i < PL_NT; i += PL_NT / 32)
            if (n0 + i < n_local && t0 + lane < T) lvl[(size_t)(n0 + i) * T + t0 + lane] = tile_l[i * (PL_TT + 2) + lane];
    }
    if (n < n_local) depth[n] = d;
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
// With the sparse layout (ts >= 0: rows of at most ts carriers are index lists, sdx_sparse.cuh) classes 0 and 1 are
// the trades that involve a bit-packed row and classes 2 and 3 the list x list trades (one per thread), split by the
// total list length so that the threads of a warp walk lists of similar length.
__device__ __forceinline__ int trade_class(int sa, int sb, int ts) {
    const int kb = sa < sb ? sa : sb;
    if (ts >= 0) {
        if (sa > ts || sb > ts) return sa + sb > 64 ? 0 : 1;
        return sa + sb > 12 ? 2 : 3;
    }
    if (sa + sb > 64) return 0;
    return kb > 8 ? 1 : (kb > 3 ? 2 : 3);
}

__global__ void k_plan_sort(int n_local, int R, int T, int Tp, int dcap, const uint32_t *__restrict__ pairs,
                            const uint16_t *__restrict__ lvl, const int *__restrict__ depth, const int *__restrict__ rowsize,
                            const uint32_t *__restrict__ desc, int ts, void *__restrict__ plan, uint16_t *__restrict__ offs,
                            int offs_stride) {
    extern __shared__ uint32_t hist_smem[];        // per warp: dcap * SDX_NCLS + 2 bins
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n = blockIdx.x * (blockDim.x >> 5) + wid;
    if (n >= n_local) return;
    const int D = depth[n];
    uint32_t *hist = hist_smem + (size_t)wid * (dcap * SDX_NCLS + 2);
    const uint16_t *lv = lvl + (size_t)n * T;
    // plan entry: (a | b << 16, t) for k_trade; (a | b << 16, t, descriptor of a, descriptor of b) for k_trade_sparse
    uint2 *out2 = reinterpret_cast<uint2 *>(plan) + (size_t)n * T;
    uint4 *out4 = reinterpret_cast<uint4 *>(plan) + (size_t)n * T;
    auto emit = [&](uint32_t pos, uint32_t e, uint32_t t) {
        if (desc) out4[pos] = make_uint4(e, t, desc[e & 0xffffu], desc[e >> 16]);
        else out2[pos] = make_uint2(e, t);
    };
    const uint32_t *pr = pairs + (size_t)n * Tp;
    if (D >= dcap) {            // very deep schedule: fall back to the sequential order (always valid)
        for (int t = lane; t < T; t += 32) emit((uint32_t)t, pr[t], (uint32_t)t);
        return;                 // k_trade runs such a null one trade per level, in order
    }
    const int NB = D * SDX_NCLS;                       // bin = (level - 1) * NCLS + class
    for (int i = lane; i <= NB; i += 32) hist[i] = 0;
    __syncwarp();
    for (int t = lane; t < T; t += 32) {
        const uint32_t e = pr[t];
        atomicAdd(&hist[(lv[t] - 1) * SDX_NCLS + trade_class(rowsize[e & 0xffffu], rowsize[e >> 16], ts)], 1u);
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
    // of[2l] = first entry of level l+1, of[2l+1] = its first class-2 entry (list x list trades start there), of[2D] = T
    uint16_t *of = offs + (size_t)n * offs_stride;
    for (int l = lane; l < D; l += 32) { of[2 * l] = (uint16_t)hist[l * SDX_NCLS]; of[2 * l + 1] = (uint16_t)hist[l * SDX_NCLS + 2]; }
    if (lane == 0) of[2 * D] = (uint16_t)T;
    __syncwarp();
    for (int t = lane; t < T; t += 32) {
        const uint32_t e = pr[t];
        uint32_t pos = atomicAdd(&hist[(lv[t] - 1) * SDX_NCLS + trade_class(rowsize[e & 0xffffu], rowsize[e >> 16], ts)], 1u);
        emit(pos, e, (uint32_t)t);
    }
}

// ---------------------------------------------------------------------------
// trade kernel
// ---------------------------------------------------------------------------
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
};

struct TradeArgs {
    PhiloxKeys keys;
    NullIndexing ix;
    const uint32_t *observed;         // R x RS
    uint32_t *work;                   // gridDim.x x R x RS (global-resident variant only)
    uint32_t *carry;                  // n_units x R x RS: chain state between launches (or null)
    const void *plan;                 // n_local x T: uint2 (a | b << 16, t), or uint4 with the two row descriptors (sparse layout)
    const uint16_t *offs;             // n_local x offs_stride level offsets
    int dcap;                         // schedules at least this deep run in sequential order
    const int *depth;                 // n_local
    int offs_stride;
    int n_units;                      // chain segments in this launch
    int R, RS, T, wprL;
    int save_carry;
    uint32_t *out_rows;               // staging: rows of launch-local null l go to out_rows + l * R * wprL
    unsigned long long *applied;      // (i - ix.null_lo)
    ScoreArgs sc;
    const uint32_t *pool;             // burn-in pool: pool_size start states in the kernel's state layout, or null
    int pool_size;
    // sparse layout (k_trade_sparse only)
    const uint32_t *desc;             // R row descriptors
    const uint32_t *obs_state;        // fp_words: observed matrix in the sparse layout
    int n_dense, fp_words;
};

// out-of-line copy of the scorer for the trade kernel: keeps its registers out of the trade loop's budget
static __device__ __noinline__ SetScore score_module_call(const uint32_t *rows, int RS, bool rows_are_samples, const int32_t *mod, int k,
                                                          const double *w, const uint32_t *mask, int S, int lane) {
    return score_module(rows, RS, rows_are_samples, mod, k, w, mask, S, lane);
}

// LB selects the launch bounds: 0 = (512, 1), 1 = (128, 5), 2 = (256, 5), 3 = (256, 4), 4 = (192, 5), 5 = (256, 3), 6 = (64, 5), 7 = (96, 5) — all
// but 0 bound the registers so that five CTAs (five nulls) stay resident per SM when the matrix fits five times in shared memory.
constexpr int lb_threads(int lb) { return lb == 0 ? 512 : (lb == 1 ? 128 : (lb == 4 ? 192 : (lb == 6 ? 64 : (lb == 7 ? 96 : 256)))); }
constexpr int lb_blocks(int lb) { return lb == 0 ? 1 : (lb == 3 ? 4 : (lb == 5 ? 3 : 5)); }
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
                const uint2 *pl = reinterpret_cast<const uint2 *>(p.plan) + (size_t)local * T;
                const uint16_t *of = p.offs + (size_t)local * p.offs_stride;
                int D = p.depth[local];
                const bool seq = D >= p.dcap;            // very deep schedule: one trade per level, in order
                if (seq) D = T;
                // level l covers plan entries [OFF(l), OFF(l+1)); e2 = end of level l+1, fetched one level ahead
                auto OFF = [&](int lv) -> int { return seq ? lv : (int)of[2 * lv]; };
                int l = 0, base = OFF(0), e = OFF(1);
                int e2 = (D >= 2) ? OFF(2) : e;
                const uint2 none = make_uint2(0u, 0u);
                uint2 ent_n = (base + myg < e) ? pl[base + myg] : none;
                while (l < D) {
                    const uint2 ent = ent_n;
                    const bool act = base + myg < e;
                    int nbase = base + NG, nl = l, ne = e;
                    const bool level_end = nbase >= e;
                    if (level_end) { nl = l + 1; nbase = e; ne = e2; }
                    ent_n = (nl < D && nbase + myg < ne) ? pl[nbase + myg] : none;
                    if (level_end) e2 = (nl + 2 <= D) ? OFF(nl + 2) : ne;
                    const bool ap = trade_rows<G, V>(rows, RS, act, ent.x & 0xffffu, ent.x >> 16, ent.y, null_id, p.keys, sub);
                    my_applied += (ap && sub == 0) ? 1u : 0u;
                    if (level_end) __syncthreads();
                    base = nbase; l = nl; e = ne;
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
            if (emit && p.sc.n_profiles > 0) {
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

// ---------------------------------------------------------------------------
// Persistent variant (experimental, SDX_TRADE_PERSIST=1; needs the matrix to fit shared memory at least twice): ONE CTA per SM owns NS null
// slots and its 20 warps pull ROUNDS — the 32/G trades one warp runs in lock step, all of one level of one slot — from
// a ring queue in shared memory.  A slot publishes the rounds of level l+1 when the last round of level l reports
// back, so the only thing that ever waits is a slot, never a warp: where k_trade parks a warp at the level barrier of
// its null, this kernel hands it a round of another null.  Epilogue work (scoring, row output) is queued the same way.
// Trades, streams and outputs are those of k_trade: results are bit-identical.
// ---------------------------------------------------------------------------
#define SM_MAX_SLOTS 8
#define SM_OFFS_CAP 1024             // staged level offsets per slot (u16): schedules of up to 511 levels
#define SM_EPI_ROUNDS 4              // epilogue rounds per null
#define SM_RING 512                  // ring entries (power of two); the host checks that all outstanding rounds fit
#define SM_THREADS 640
#ifdef SDX_SM_DEBUG
#define SM_DBG(...) do { if (lane == 0 && blockIdx.x == 0) printf(__VA_ARGS__); } while (0)
#else
#define SM_DBG(...) do { } while (0)
#endif

struct SmSlot {
    unsigned int pending;            // rounds of the current level still running or queued
    unsigned int applied;
    int level;                       // current level; level == D is the epilogue
    int D;                           // trade levels of the current null
    int seq;                         // very deep schedule: one trade per level, in plan order
    int unit, j, emit;
    long long local;                 // launch-local null index
    unsigned long long null_id;
};

struct SmShared {
    SmSlot slot[SM_MAX_SLOTS];
    unsigned long long ring[SM_RING];         // (ticket = index + 1) << 32 | payload
    unsigned int head, reserve, n_finished;
};

// payload: slot (bits 0-2) | epilogue flag (bit 3) | trades in the round (bits 4-9) | first plan entry or epilogue round (bits 10-25)
__device__ __forceinline__ unsigned int sm_payload(int s, int epi, int cnt, int base) {
    return (unsigned)s | ((unsigned)epi << 3) | ((unsigned)cnt << 4) | ((unsigned)base << 10);
}

// D, seq: the slot's constants for the current null, read by the caller BEFORE any lane may change the slot (lanes of
// a warp are not in lock step: a lane that lags must not see what lane 0 writes further down).
template <int G>
__device__ __forceinline__ void sm_push_level(SmShared &sh, const uint16_t *of_s, int s, int lv, int D, int seq, int lane) {
    constexpr int GPW = 32 / G;
    SmSlot &sl = sh.slot[s];
    const bool epi = lv == D;
    int base = 0, e = SM_EPI_ROUNDS * GPW;
    if (!epi) {
        base = seq ? lv : (int)of_s[2 * lv];
        e = seq ? lv + 1 : (int)of_s[2 * lv + 2];
    }
    const int nr = (e - base + GPW - 1) / GPW;
    __syncwarp();
    unsigned int t0 = 0;
    if (lane == 0) {
        sl.level = lv;
        sl.pending = (unsigned)nr;
        __threadfence_block();
        t0 = atomicAdd(&sh.reserve, (unsigned)nr);
    }
    t0 = __shfl_sync(0xffffffffu, t0, 0);
    SM_DBG("push s=%d lv=%d D=%d epi=%d base=%d e=%d nr=%d t0=%u\n", s, lv, D, (int)epi, base, e, nr, t0);
    for (int i = lane; i < nr; i += 32) {
        const int b = base + i * GPW;
        const int cnt = e - b < GPW ? e - b : GPW;
        const unsigned int pay = epi ? sm_payload(s, 1, 0, i) : sm_payload(s, 0, cnt, b);
        const unsigned int idx = t0 + (unsigned)i;
        *reinterpret_cast<volatile unsigned long long *>(&sh.ring[idx & (SM_RING - 1)]) = ((unsigned long long)(idx + 1u) << 32) | pay;
    }
    __syncwarp();          // every entry is in the ring before any lane of this warp starts polling it (a lane spinning
                           // on an entry that a sibling lane still has to write would starve that lane)
}

// Moves slot s to its next null (or retires it) and queues the first rounds.  One warp, all lanes.
template <int G>
__device__ __noinline__ void sm_next_null(const TradeArgs &p, SmShared &sh, uint32_t *rows, uint16_t *of_s, int s, int NS, bool first, int lane) {
    SmSlot &sl = sh.slot[s];
    const int R = p.R, RS = p.RS, T = p.T;
    const int nvec = R * RS / 4;
    int unit = sl.unit, j = first ? 0 : sl.j + 1;
    __syncwarp();                      // every lane has read the slot before lane 0 rewrites it below
    for (;;) {
        bool have = unit < p.n_units && j < p.ix.seg_len;
        long long local = 0;
        unsigned long long null_id = 0;
        if (have) {
            local = (long long)unit * p.ix.seg_len + j;
            null_id = p.ix.id(local);
            have = null_id < p.ix.null_hi;
        }
        if (!have) {
            if (unit < p.n_units && p.save_carry && !(first && j == 0)) {
                uint4 *dst = reinterpret_cast<uint4 *>(p.carry + (size_t)unit * R * RS);
                for (int i = lane; i < nvec; i += 32) dst[i] = reinterpret_cast<const uint4 *>(rows)[i];
            }
            if (unit < p.n_units) unit += NS * (int)gridDim.x;
            j = 0;
            first = true;
            if (unit >= p.n_units) {
                if (lane == 0) { sl.unit = unit; __threadfence_block(); atomicAdd(&sh.n_finished, 1u); }
                return;
            }
            continue;
        }
        const bool restart = (null_id % (unsigned long long)p.ix.chain_len) == 0;
        if (restart || j == 0) {
            const uint32_t *src = p.carry + (size_t)unit * R * RS;
            if (restart) src = p.pool ? p.pool + (size_t)((null_id / (unsigned long long)p.ix.chain_len) % (unsigned long long)p.pool_size) * R * RS : p.observed;
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(rows);
            int i = lane;
            for (; i + 96 < nvec; i += 128) {                  // four loads in flight per lane
                const uint4 a = s4[i], b = s4[i + 32], c = s4[i + 64], d = s4[i + 96];
                d4[i] = a; d4[i + 32] = b; d4[i + 64] = c; d4[i + 96] = d;
            }
            for (; i < nvec; i += 32) d4[i] = s4[i];
        }
        int D = T > 0 ? p.depth[local] : 0;
        const int seq = D >= p.dcap;
        if (seq) D = T;
        else {
            const uint16_t *of = p.offs + (size_t)local * p.offs_stride;
            for (int i = lane; i <= 2 * D; i += 32) of_s[i] = of[i];
        }
        const int emit = null_id >= p.ix.null_lo;
        SM_DBG("null s=%d unit=%d j=%d local=%lld id=%llu D=%d seq=%d emit=%d restart=%d\n", s, unit, j, local, null_id, D, seq, emit, (int)restart);
        if (lane == 0) {
            sl.unit = unit; sl.j = j; sl.local = local; sl.null_id = null_id; sl.D = D; sl.seq = seq; sl.applied = 0; sl.emit = emit;
        }
        __syncwarp();
        __threadfence_block();
        const bool has_epi = emit && (p.out_rows || p.sc.n_profiles > 0);
        if (D > 0 || has_epi) {
            sm_push_level<G>(sh, of_s, s, 0, D, seq, lane);    // level 0 == D means: epilogue only
            return;
        }
        if (emit && p.applied && lane == 0) p.applied[null_id - p.ix.null_lo] = 0ull;
        first = false;
        ++j;
    }
}

template <int G, int V>
__global__ void __launch_bounds__(SM_THREADS, 1) k_trade_sm(const __grid_constant__ TradeArgs p, const int NS) {
    extern __shared__ __align__(16) uint32_t smem_rows[];
    __shared__ SmShared sh;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int GPW = 32 / G;
    const int sub = lane & (G - 1), grp = lane / G;
    const int R = p.R, RS = p.RS, T = p.T;
    const size_t slot_words = (size_t)R * RS;
    uint16_t *offs_all = reinterpret_cast<uint16_t *>(smem_rows + (size_t)NS * slot_words);

    if (tid < SM_MAX_SLOTS) {
        sh.slot[tid].pending = 0; sh.slot[tid].level = 0; sh.slot[tid].D = 0;
        sh.slot[tid].unit = tid * (int)gridDim.x + (int)blockIdx.x;          // slot s starts with unit s * grid + b
        sh.slot[tid].j = 0;
    }
    if (tid == 0) { sh.head = 0; sh.reserve = 0; sh.n_finished = 0; }
    for (int i = tid; i < SM_RING; i += blockDim.x) sh.ring[i] = 0ull;
    __syncthreads();
    if (wid < NS) sm_next_null<G>(p, sh, smem_rows + (size_t)wid * slot_words, offs_all + (size_t)wid * SM_OFFS_CAP, wid, NS, true, lane);

    for (;;) {
        // ---- take a round from the queue (lane 0 polls) ---------------------------------------------------------
        unsigned int pay = 0xffffffffu;
        if (lane == 0) {
            for (;;) {
                const unsigned int h = *reinterpret_cast<volatile unsigned int *>(&sh.head);
                const unsigned int rsv = *reinterpret_cast<volatile unsigned int *>(&sh.reserve);
                if ((int)(rsv - h) > 0) {
                    if (atomicCAS(&sh.head, h, h + 1u) != h) continue;
                    unsigned long long e;
                    do { e = *reinterpret_cast<volatile unsigned long long *>(&sh.ring[h & (SM_RING - 1)]); } while ((unsigned int)(e >> 32) != h + 1u);
                    pay = (unsigned int)e;
                    break;
                }
                if (*reinterpret_cast<volatile unsigned int *>(&sh.n_finished) >= (unsigned)NS) break;
                __nanosleep(64);
            }
        }
        pay = __shfl_sync(0xffffffffu, pay, 0);
        if (pay == 0xffffffffu) break;
        __threadfence_block();
        const int s = (int)(pay & 7u), epi = (int)((pay >> 3) & 1u), cnt = (int)((pay >> 4) & 63u), base = (int)((pay >> 10) & 0xffffu);
        SmSlot &sl = sh.slot[s];
        uint32_t *rows = smem_rows + (size_t)s * slot_words;
        SM_DBG("take w=%d s=%d epi=%d cnt=%d base=%d local=%lld\n", wid, s, epi, cnt, base, sl.local);
        const long long local = sl.local;
        const unsigned long long null_id = sl.null_id;

        if (!epi) {
            const uint2 *pl = reinterpret_cast<const uint2 *>(p.plan) + (size_t)local * T;
            const bool act = grp < cnt;
            const uint2 ent = act ? pl[base + grp] : make_uint2(0u, 0u);
            const bool ap = trade_rows<G, V>(rows, RS, act, ent.x & 0xffffu, ent.x >> 16, ent.y, null_id, p.keys, sub);
            if (p.applied) {
                const unsigned int n_ap = __popc(__ballot_sync(0xffffffffu, ap && sub == 0));
                if (lane == 0 && n_ap) atomicAdd(&sl.applied, n_ap);
            }
        } else if (sl.emit) {
            const int r = base;                                   // epilogue round r of SM_EPI_ROUNDS
            if (p.out_rows) {
                uint32_t *dst = p.out_rows + (size_t)local * R * p.wprL;
                for (int row = r; row < R; row += SM_EPI_ROUNDS)
                    for (int c = lane; c < p.wprL; c += 32) dst[(size_t)row * p.wprL + c] = rows[(size_t)row * RS + c];
            }
            for (int pr = r; pr < p.sc.n_profiles; pr += SM_EPI_ROUNDS) {
                SetScore sc = score_module_call(rows, RS, p.sc.rows_are_samples != 0, p.sc.modules + (size_t)pr * p.sc.k, p.sc.k,
                                                p.sc.w + (size_t)pr * p.sc.S, p.sc.mask + (size_t)pr * p.sc.wprS, p.sc.S, lane);
                if (lane == 0) {
                    if (sc.W >= p.sc.z_obs[pr]) atomicAdd(&p.sc.n_ge[pr], 1ull);
                    if (p.sc.null_scores) p.sc.null_scores[(size_t)pr * p.sc.n_total + (null_id - p.ix.null_lo)] = sc.W;
                }
            }
        }

        // ---- report back; the warp that completes a level moves the slot on ----------------------------------------
        __syncwarp();
        unsigned int left = 1;
        if (lane == 0) { __threadfence_block(); left = atomicSub(&sl.pending, 1u) - 1u; }
        left = __shfl_sync(0xffffffffu, left, 0);
        SM_DBG("done w=%d s=%d left=%u\n", wid, s, left);
        if (left == 0) {
            __threadfence_block();
            int st = 0, D = 0;                                   // lane 0 reads the slot, the others get copies
            if (lane == 0) { st = sl.level | (sl.seq << 30) | (sl.emit << 29); D = sl.D; }
            st = __shfl_sync(0xffffffffu, st, 0);
            D = __shfl_sync(0xffffffffu, D, 0);
            const int lv = st & 0xffffff, seq = (st >> 30) & 1, emit = (st >> 29) & 1;
            const bool has_epi = emit && (p.out_rows || p.sc.n_profiles > 0);
            uint16_t *of_s = offs_all + (size_t)s * SM_OFFS_CAP;
            if (lv + 1 < D || (lv + 1 == D && has_epi)) {
                sm_push_level<G>(sh, of_s, s, lv + 1, D, seq, lane);
            } else {
                if (emit && p.applied && lane == 0) p.applied[null_id - p.ix.null_lo] = sl.applied;
                sm_next_null<G>(p, sh, rows, of_s, s, NS, false, lane);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Sparse layout: conversion of the observed matrix and the one-warp-per-null trade kernel (sdx_sparse.cuh)
#ifndef SP_MINB
#define SP_MINB 16
#endif
// ---------------------------------------------------------------------------
__global__ void k_to_sparse(const uint32_t *__restrict__ rows, int R, int RS, const uint32_t *__restrict__ desc, int n_dense,
                            uint32_t *__restrict__ state) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    const uint32_t d = desc[r];
    const uint32_t w = lane < RS ? rows[(size_t)r * RS + lane] : 0u;     // RS <= 32: lane i holds word i
    if (sp_is_dense(d)) {
        if (lane < RS) state[(size_t)sp_off(d) * RS + lane] = w;
        return;
    }
    uint16_t *list = reinterpret_cast<uint16_t *>(state + (size_t)n_dense * RS) + sp_off(d);
    int inc = __popc(w);
    const int cnt = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    int pos = inc - cnt;
    for (uint32_t x = w; x; x &= x - 1) list[pos++] = (uint16_t)(lane * 32 + __ffs(x) - 1);
}

__global__ void __launch_bounds__(32, SP_MINB) k_trade_sparse(const __grid_constant__ TradeArgs p) {
    extern __shared__ __align__(16) uint32_t sp_smem[];
    __shared__ int32_t s_mod[SDX_MAX_K];
    const int lane = threadIdx.x;
    const int sub = lane & (SP_G - 1), grp = lane / SP_G;
    const int R = p.R, RS = p.RS, T = p.T;
    // layout: state (bit-packed rows, then the list pool) | 2 scratch rows per group | list scratch of the 32 threads
    uint32_t *state = sp_smem;
    uint32_t *gscr = sp_smem + p.fp_words;
    uint16_t *lscr = reinterpret_cast<uint16_t *>(gscr);       // the two scratch areas are never live at the same time
    SparseView v;
    v.bits = state; v.lists = reinterpret_cast<uint16_t *>(state + (size_t)p.n_dense * RS); v.desc = p.desc; v.RS = RS;
    const int nvec = p.fp_words / 4;

    for (int u = blockIdx.x; u < p.n_units; u += gridDim.x) {
        for (int j = 0; j < p.ix.seg_len; ++j) {
            const int64_t local = (int64_t)u * p.ix.seg_len + j;
            const uint64_t null_id = p.ix.id(local);
            if (null_id >= p.ix.null_hi) break;
            const bool restart = (null_id % (uint64_t)p.ix.chain_len) == 0;
            if (restart || j == 0) {
                const uint32_t *src = p.carry + (size_t)u * p.fp_words;
                if (restart) src = p.pool ? p.pool + (size_t)((null_id / (uint64_t)p.ix.chain_len) % (uint64_t)p.pool_size) * p.fp_words : p.obs_state;
                for (int i = lane; i < nvec; i += 32)
                    reinterpret_cast<uint4 *>(state)[i] = reinterpret_cast<const uint4 *>(src)[i];
            }
            __syncwarp();

            unsigned int my_applied = 0;
            if (T > 0) {
                const uint4 *pl = reinterpret_cast<const uint4 *>(p.plan) + (size_t)local * T;
                const uint16_t *of = p.offs + (size_t)local * p.offs_stride;
                int D = p.depth[local];
                const bool seq = D >= p.dcap;            // very deep schedule: one trade per level, in order
                if (seq) D = T;
                // Level l = plan entries [s, e): [s, m) involve a bit-packed row (SP_GPW per round, 4 lanes each),
                // [m, e) are list x list trades (32 per round, one per thread).  Offsets are fetched two levels
                // ahead and the entry of a level's first round one level ahead.
                auto OFFS = [&](int lv, int &s_, int &m_, int &e_) {
                    if (lv >= D) { s_ = m_ = e_ = 0; }
                    else if (seq) { s_ = lv; m_ = lv + 1; e_ = lv + 1; }
                    else { s_ = of[2 * lv]; m_ = of[2 * lv + 1]; e_ = of[2 * lv + 2]; }
                };
                const uint4 none = make_uint4(0xffffffffu, 0u, 0u, 0u);
                auto ENTRY = [&](int s_, int m_, int e_, int r_) -> uint4 {       // this lane's entry of round r_ of a level
                    const int ug_ = (m_ - s_ + SP_GPW - 1) / SP_GPW;
                    const int idx = r_ < ug_ ? s_ + r_ * SP_GPW + grp : m_ + (r_ - ug_) * 32 + lane;
                    return idx < (r_ < ug_ ? m_ : e_) ? pl[idx] : none;
                };
                int s0, m0, e0, s1, m1, e1, s2, m2, e2;
                OFFS(0, s0, m0, e0);
                OFFS(1, s1, m1, e1);
                uint4 ent_n = ENTRY(s0, m0, e0, 0);
                for (int l = 0; l < D; ++l) {
                    OFFS(l + 2, s2, m2, e2);
                    const int ug = (m0 - s0 + SP_GPW - 1) / SP_GPW, un = ug + ((e0 - m0 + 31) >> 5);
                    for (int r = 0; r < un; ++r) {
                        const uint4 ent = ent_n;
                        ent_n = r + 1 < un ? ENTRY(s0, m0, e0, r + 1) : (l + 1 < D ? ENTRY(s1, m1, e1, 0) : none);
                        const bool act = ent.x != 0xffffffffu;
                        if (r < ug) {
                            // bit-packed path: list rows go through the group's scratch bitmaps
                            const bool la = act && !sp_is_dense(ent.z), lb = act && !sp_is_dense(ent.w);
                            uint32_t *pa = act ? (la ? gscr + (grp * 2) * RS : v.dense_row(ent.z)) : state;
                            uint32_t *pb = act ? (lb ? gscr + (grp * 2 + 1) * RS : v.dense_row(ent.w)) : state;
                            sp_group_expand(v.lists + sp_off(ent.z), sp_size(ent.z), pa, RS, sub, la);
                            sp_group_expand(v.lists + sp_off(ent.w), sp_size(ent.w), pb, RS, sub, lb);
                            const bool ap = trade_row_ptrs<SP_G, SP_V>(pa, pb, RS, act, ent.y, null_id, p.keys, sub);
                            __syncwarp();
                            sp_group_compact(pa, v.lists + sp_off(ent.z), RS, sub, la && ap);
                            sp_group_compact(pb, v.lists + sp_off(ent.w), RS, sub, lb && ap);
                            my_applied += (ap && sub == 0) ? 1u : 0u;
                        } else {
                            uint16_t *A = v.lists + sp_off(ent.z), *B = v.lists + sp_off(ent.w);
                            const bool ap = sp_trade_lists(A, sp_size(ent.z), B, sp_size(ent.w), act, ent.y, null_id, p.keys,
                                                           lscr + lane, lscr + (SP_TS + 1) * 32 + lane, 32);
                            my_applied += ap ? 1u : 0u;
                        }
                    }
                    __syncwarp();
                    s0 = s1; m0 = m1; e0 = e1; s1 = s2; m1 = m2; e1 = e2;
                }
            }
            __syncwarp();

            // ---- outputs -----------------------------------------------------
            const bool emit = null_id >= p.ix.null_lo;
            if (emit && p.applied) {
                my_applied = __reduce_add_sync(0xffffffffu, my_applied);
                if (lane == 0) p.applied[null_id - p.ix.null_lo] = my_applied;
            }
            if (emit && p.out_rows) {
                uint32_t *dst = p.out_rows + (size_t)local * R * p.wprL;
                for (int r = 0; r < R; ++r) {
                    sp_row_to_bitmap(v, p.desc[r], gscr, lane);
                    if (lane < p.wprL) dst[(size_t)r * p.wprL + lane] = gscr[lane];
                    __syncwarp();
                }
            }
            if (emit && p.sc.n_profiles > 0) {
                for (int pr = 0; pr < p.sc.n_profiles; ++pr) {
                    const int32_t *mod = p.sc.modules + (size_t)pr * p.sc.k;
                    for (int q = 0; q < p.sc.k; ++q) {             // the module's rows as bitmaps in the group scratch
                        const int g = mod[q];
                        if (g >= 0) sp_row_to_bitmap(v, p.desc[g], gscr + q * RS, lane);
                        if (lane == 0) s_mod[q] = g >= 0 ? q : -1;
                    }
                    __syncwarp();
                    SetScore sc = score_module(gscr, RS, false, s_mod, p.sc.k, p.sc.w + (size_t)pr * p.sc.S,
                                               p.sc.mask + (size_t)pr * p.sc.wprS, p.sc.S, lane);
                    if (lane == 0) {
                        if (sc.W >= p.sc.z_obs[pr]) atomicAdd(&p.sc.n_ge[pr], 1ull);
                        if (p.sc.null_scores) p.sc.null_scores[(size_t)pr * p.sc.n_total + (null_id - p.ix.null_lo)] = sc.W;
                    }
                    __syncwarp();
                }
            }
        }
        if (p.save_carry) {
            uint32_t *dst = p.carry + (size_t)u * p.fp_words;
            for (int i = lane; i < nvec; i += 32)
                reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(state)[i];
        }
        __syncwarp();
    }
}

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
        static const int gv[][2] = {{2, 16}, {4, 8}, {8, 4}, {8, 8}, {16, 4}, {32, 4}, {32, 5}, {32, 8}};
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
            if (lb >= 0 && lb <= 7 && cfg->threads <= lb_threads(lb)) cfg->LB = lb;
        }
        cfg->grid_cap = 1 << 30;
        {
            int resident = per_sm;                                      // shared memory limit
            const int by_threads = 2048 / cfg->threads;
            if (resident > by_threads) resident = by_threads;
            if (cfg->LB != 0 && resident > (cfg->LB == 3 ? 4 : (cfg->LB == 5 ? 3 : 5))) resident = cfg->LB == 3 ? 4 : (cfg->LB == 5 ? 3 : 5);
            cfg->slots = resident * c->sm_count;
        }
    } else {
        cfg->smem_bytes = 0;
        cfg->LB = 0;
        int warps = 16;
        if (warps_env >= 1 && warps_env <= 16) warps = warps_env;
        cfg->threads = warps * 32;
        size_t per = bytes;
        int cap = (int)((size_t)96 * 1024 * 1024 / per);   // working set ~ within L2
        if (cap < c->sm_count) cap = c->sm_count;
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
        auto kern = cfg.LB == 1 ? k_trade<G, V, true, 1> : (cfg.LB == 2 ? k_trade<G, V, true, 2> : (cfg.LB == 3 ? k_trade<G, V, true, 3> : (cfg.LB == 4 ? k_trade<G, V, true, 4> : (cfg.LB == 5 ? k_trade<G, V, true, 5> : (cfg.LB == 6 ? k_trade<G, V, true, 6> : (cfg.LB == 7 ? k_trade<G, V, true, 7> : k_trade<G, V, true, 0>))))));
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem_bytes);
        if (e != cudaSuccess) return e;
        kern<<<grid, cfg.threads, cfg.smem_bytes, st>>>(args);
    } else {
        k_trade<G, V, false, 0><<<grid, cfg.threads, 0, st>>>(args);
    }
    return cudaGetLastError();
}

static cudaError_t dispatch_trade(const TradeConfig &cfg, int grid, cudaStream_t st, const TradeArgs &args, bool persist) {
    if (cfg.G == 2 && cfg.V == 16) return launch_trade<2, 16>(cfg, grid, st, args, persist);
    if (cfg.G == 4 && cfg.V == 8) return launch_trade<4, 8>(cfg, grid, st, args, persist);
    if (cfg.G == 8 && cfg.V == 4) return launch_trade<8, 4>(cfg, grid, st, args, persist);
    if (cfg.G == 8 && cfg.V == 8) return launch_trade<8, 8>(cfg, grid, st, args, persist);
    if (cfg.G == 16 && cfg.V == 4) return launch_trade<16, 4>(cfg, grid, st, args, persist);
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

enum { SL_LVL = 0, SL_PLAN, SL_AUX, SL_MOD, SL_Z, SL_SCORES, SL_CARRY, SL_OUT, SL_APPLIED, SL_WORK, SL_OFFS = 12, SL_PAIRS = 13 };

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

#define SDX_POOL_BASE (1ull << 61)      // stream ids of the burn-in rounds live above every null index

// The burn-in pool of sdx_set_burn_in: pool_size chains of `burn` rounds from the observed matrix; their final states
// (in the trade kernel's own state layout) are the start states of the real chains.  Cached per (seed, T).
static int ensure_pool(sdx_ctx *c, uint64_t seed, int64_t T, size_t state_words) {
    if (c->burn <= 0) return SDX_OK;
    if (c->pool_valid && c->pool_seed == seed && c->pool_T == T && c->pool_state_words == state_words) return SDX_OK;
    c->pool_valid = false;
    SDX_TRY(sdx_reserve(c, c->d_pool, c->cap_pool, sizeof(uint32_t) * state_words * (size_t)c->pool_size));
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
    SDX_REQUIRE(c, null0 + (uint64_t)n_nulls < (1ull << 62) && (!use_pool || null0 + (uint64_t)n_nulls <= SDX_POOL_BASE), SDX_ERR_INVALID,
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
        if ((rc = sdx_scratch(c, SL_MOD, sizeof(int32_t) * c->P * hs->k + 16, &pm))) return rc;
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
        if (hs->z_obs) {
            SDX_CUDA(c, cudaMemcpyAsync(d_zobs, hs->z_obs, sizeof(double) * c->P, cudaMemcpyHostToDevice, c->stream));
        } else {
            k_score_observed<<<ceil_div(c->P, 4), 128, 0, c->stream>>>(c->d_trade_obs, RS, sc, d_zobs);
            timer.count();
        }
    }

    float plan_ms = 0, trade_ms = 0;
    if (n_nulls > 0) {
        // chains overlapping [null0, null0 + n)
        const uint64_t null_hi = null0 + (uint64_t)n_nulls;
        const uint64_t chain_first = null0 / (uint64_t)chain_len;
        const uint64_t chain_last = (null_hi - 1) / (uint64_t)chain_len;
        const int64_t n_chains = (int64_t)(chain_last - chain_first + 1);
        const int64_t NB = 32768;                                   // nulls planned per launch
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
        const bool last_in_smem = lv_tiles + (size_t)PL_NT * R * 2 <= smem_cap;
        const size_t lv_smem = lv_tiles + (last_in_smem ? (size_t)PL_NT * R * 2 : 0);
        if (lv_smem > 48 * 1024)
            SDX_CUDA(c, cudaFuncSetAttribute(k_plan_levels, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lv_smem));

        void *p_lvl, *p_plan, *p_offs, *p_aux = nullptr, *p_carry = nullptr, *p_work = nullptr, *p_out = nullptr, *p_app = nullptr;
        const int offs_stride = (int)((2 * (T < dcap ? T : dcap) + 2 + 1) & ~1ll);
        const size_t sort_smem_warp = sizeof(uint32_t) * ((size_t)dcap * SDX_NCLS + 2);
        const int sort_warps = sort_smem_warp * 8 <= 40 * 1024 ? 8 : (sort_smem_warp * 4 <= 40 * 1024 ? 4 : 2);
        if ((rc = sdx_scratch(c, SL_OFFS, sizeof(uint16_t) * (size_t)n_local_max * offs_stride + 64, &p_offs))) return rc;
        const size_t lvl_bytes = ((sizeof(uint16_t) * (size_t)n_local_max * Tn + 15) / 16) * 16;
        if ((rc = sdx_scratch(c, SL_LVL, lvl_bytes + sizeof(int) * n_local_max + 64, &p_lvl))) return rc;
        int *d_depth = (int *)((char *)p_lvl + lvl_bytes);
        if ((rc = sdx_scratch(c, SL_PLAN, (sparse ? sizeof(uint4) : sizeof(uint2)) * (size_t)n_local_max * Tn, &p_plan))) return rc;
        void *p_pairs;
        const int Tp = (Tn + 3) & ~3;                       // pairs row length: 16-byte aligned rows
        if ((rc = sdx_scratch(c, SL_PAIRS, sizeof(uint32_t) * (size_t)n_local_max * Tp + 64, &p_pairs))) return rc;
        if (!last_in_smem) {
            const size_t nb = (size_t)ceil_div(n_local_max, lv_threads) * lv_threads;
            if ((rc = sdx_scratch(c, SL_AUX, sizeof(uint16_t) * nb * R, &p_aux))) return rc;
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
                    k_plan_pairs<<<ceil_div((int64_t)n_local * Tp, 256), 256, 0, c->stream>>>(keys, ix, n_local, R, (int)T, Tp, (uint32_t *)p_pairs);
                    k_plan_levels<<<ceil_div(n_local, lv_threads), lv_threads, lv_smem, c->stream>>>(
                        n_local, R, (int)T, Tp, (const uint32_t *)p_pairs, (uint16_t *)p_lvl, d_depth, (uint16_t *)p_aux);
                    k_plan_sort<<<ceil_div(n_local, sort_warps), sort_warps * 32, sort_smem_warp * sort_warps, c->stream>>>(
                                                                          n_local, R, (int)T, Tp, dcap, (const uint32_t *)p_pairs,
                                                                          (const uint16_t *)p_lvl, d_depth, c->d_rowsize, sparse ? c->d_sp_desc : nullptr,
                                                                          sparse ? SP_TS : -1, p_plan,
                                                                          (uint16_t *)p_offs, offs_stride);
                    timer.count(3);
                }
                cudaEventRecord(c->ev[3], c->stream);
                TradeArgs ta{};
                ta.keys = keys; ta.ix = ix; ta.observed = c->d_trade_obs; ta.work = (uint32_t *)p_work;
                ta.carry = (uint32_t *)p_carry; ta.plan = p_plan; ta.n_units = (int)nch;
                ta.offs = (const uint16_t *)p_offs; ta.depth = d_depth; ta.offs_stride = offs_stride; ta.dcap = dcap;
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
    c->timing.plan_ms = plan_ms; c->timing.trade_ms = trade_ms;
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
