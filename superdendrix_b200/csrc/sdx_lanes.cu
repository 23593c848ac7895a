// sdx_lanes.cu — k_trade_lanes: Curveball null generation with one THREAD per null (+ fused fixed-set scoring).
//
// Replaces, behind the C ABI of include/sdx.h (sdx_curveball / sdx_null_pvalue with sdx_set_pair_group(ctx, 32)):
//   src/curveball.py:17-40               curve_ball
//   src/generate_null_matrices.py:71-74  the null loop over one mutable presence list
//   src/superdendrix.py:166-197          the null loop / exceedance count (stat = fixed)
//
// Why: k_trade (sdx_curveball.cu) spends ~109 warp instructions per trade because only 8 trades share an
// instruction (profiles/r01_k_trade_ncu_summary.md).  Here the 32 lanes of a warp are 32 NULLS that walk the same
// schedule of row pairs (sdx_set_pair_group: the PAIR stream is shared by aligned blocks of 32 chains, every null
// keeps its own DEAL stream).  Row sizes are invariants of Curveball, so the two rows of a trade have the same sizes
// in all 32 nulls: every loop bound is warp-uniform, a trade is plain per-thread code (sdx_lane_trade.cuh), there is
// no level plan, no plan memory and no barrier.
//
// Data movement: the state of a group of 32 nulls (C1: ~200 KB, lists of u16 indices for rows of <= TS elements,
// bit-packed words for the others, lanes interleaved) lives in global memory / L2.  One elected lane moves the LIST
// rows of a trade into shared memory with TMA bulk copies (cp.async.bulk + mbarrier, two trades ahead: the schedule is
// known, so a row pair is in flight while the previous two trades run) and the new lists back with bulk stores.  A trade
// that shares a list with one of the three trades before it cannot be prefetched blindly: it waits for the bulk stores
// in flight (cp.async.bulk.wait_group) or is loaded when its turn comes (~3 % of the trades).  The group states of a
// launch (C1: 3 125 x 220 KB) do not fit L2, so every row comes from HBM: the prefetch is what hides that latency.
// A bit-packed row (the few dense ones) of a bit-packed x list trade has ONE large buffer X per warp, filled one trade
// ahead; a trade between two bit-packed rows (1 %) works in place in the group state.  A warp needs 6 list buffers + X +
// the prefix popcounts (C1: 10 KB), so 22 warps share an SM.
//
// The burn-in pool (sdx_set_burn_in) is built by k_trade — 1 472 states are too few lanes for this kernel (a warp needs
// 6.8 ms per round however few groups there are), and its rounds trade the other orientation of the matrix (DESIGN.md
// 3b) — and converted once to the lane layout (k_ln_compact): whole
// groups when the pool size is a multiple of 32 (a chain's start state is then one block copy), compact states otherwise.
#include <algorithm>

#include "sdx_lane_trade.cuh"
#include "sdx_nulls.cuh"

// ---------------------------------------------------------------------------
// TMA bulk copy / mbarrier / proxy fence wrappers (PTX ISA 8.x; sm_90+)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();              // a copy that never lands is a bug: fail the launch, do not hang the GPU
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---------------------------------------------------------------------------
struct LaneArgs {
    PhiloxKeys keys;
    const uint2 *desc;                // R row descriptors
    int R, T, wprL;
    int slot_bytes;                   // one staging buffer
    int desc_bytes, warp_bytes;       // shared-memory layout: descriptor table, then warp_bytes per warp
    int x_bytes;                      // the buffer of a bit-packed row inside warp_bytes
    uint32_t gsb, rows_bytes, c_off, pp_off;      // group state: rows, rank-mask scratch, scorer partials
    uint8_t *states;                  // n_groups x gsb
    uint64_t chain0;                  // first chain of group 0 of this launch (a multiple of 32)
    int64_t chain_len;
    int64_t seg_off;                  // rounds [seg_off, seg_off + seg_len) of every chain run in this launch
    int seg_len;
    uint64_t null_lo, null_hi;        // requested nulls
    int n_groups;
    const uint8_t *src_states;        // start states: n_src groups of rows_bytes (src_groups), else n_src compact states of rows_bytes / 32
    int n_src;
    int src_groups;
    uint32_t *out_rows;               // (null_id - null_lo) x R x wprL, zero-filled, or null
    unsigned long long *applied;      // (null_id - null_lo), or null
    ScoreArgs sc;
};

// bit-packed states (n x R x RS words, the layout of k_trade; src_stride = 0: the same state n times) -> lane layout, one
// warp per (state, row).  interleave = 0: compact states of csb bytes each (entries contiguous); 1: groups of 32 states,
// csb bytes per group, state st in lane st % 32 of group st / 32.
__global__ void k_ln_compact(const uint32_t *__restrict__ rows, size_t src_stride, int n_states, int R, int RS, int wprL,
                             const uint2 *__restrict__ desc, uint32_t csb, int interleave, uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= (int64_t)n_states * R) return;
    const int st = (int)(job / R), r = (int)(job - (int64_t)st * R);
    const uint32_t *src = rows + (size_t)st * src_stride + (size_t)r * RS;
    const uint2 d = desc[r];
    const int stride = interleave ? LN_W : 1;
    uint8_t *dst = interleave ? out + (size_t)(st >> 5) * csb + d.x : out + (size_t)st * csb + d.x / LN_W;
    const int col = interleave ? (st & 31) : 0;
    if (ln_is_bits(d)) {
        for (int w = lane; w < wprL; w += 32) reinterpret_cast<uint32_t *>(dst)[(size_t)w * stride + col] = src[w];
        return;
    }
    uint16_t *list = reinterpret_cast<uint16_t *>(dst);
    int base = 0;
    for (int w0 = 0; w0 < wprL; w0 += 32) {
        const int w = w0 + lane;
        const uint32_t x = w < wprL ? src[w] : 0u;
        const int cnt = __popc(x);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        int pos = base + inc - cnt;
        for (uint32_t y = x; y; y &= y - 1) { list[(size_t)pos * stride + col] = (uint16_t)(w * 32 + __ffs(y) - 1); ++pos; }
        base += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ---------------------------------------------------------------------------
// the trade kernel: one warp per group of 32 chains
// ---------------------------------------------------------------------------
// entries of 32 trades, one per lane: a | b << 14 | flags << 28 (one copy of the PAIR draw in the kernel: code size)
static __device__ __noinline__ uint32_t ln_make_entries(const PhiloxKeys *keys, uint64_t pid, uint32_t t0, uint32_t prev, uint32_t T, uint32_t R) {
    const int lane = threadIdx.x & 31;
    const uint32_t u = t0 + (uint32_t)lane;
    uint32_t a = 0, b = 1;
    if (u < T) trade_pair(*keys, pid, u, R, a, b);
    uint32_t e = a | (b << 14), fl = 0;
#pragma unroll
    for (int d = 1; d <= 3; ++d) {
        const uint32_t q0 = __shfl_sync(0xffffffffu, e, (lane - d) & 31), q1 = __shfl_sync(0xffffffffu, prev, (lane - d) & 31);
        const uint32_t q = lane >= d ? q0 : q1;
        fl |= (u >= (uint32_t)d && ln_entry_rows_share(e, q)) ? (1u << (d - 1)) : 0u;
    }
    return e | (fl << 28);
}

#define LN_FLAG_PREV 1u               // shares a row with the trade before it: loaded when its turn comes
#define LN_FLAG_NEAR 6u               // shares a row with the second / third trade before it: wait for the stores in flight
#define LN_MAX_THREADS 704
// REGS: registers per thread.  80 keeps 22 warps (the groups of 100 000 nulls in ONE wave) resident per SM; 128 has no
// spills and is what a launch with fewer groups than that runs (registers are allocated in units of 512 per warp, so 88
// would already drop to 20 warps).
// SLOT_C: the size of a list buffer as a compile-time constant (1 088 B = lists of up to 16 entries + sentinel, the common case:
// every buffer of the warp is then an immediate offset from one base register), or 0 = taken from the arguments.
template <bool M64, int REGS, int SLOT_C>
__global__ void __maxnreg__(REGS) k_trade_lanes(const __grid_constant__ LaneArgs p) {
    extern __shared__ __align__(128) uint8_t ln_smem[];
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, WPC = blockDim.x >> 5;
    const int R = p.R, T = p.T, nw = p.wprL, SLOT = SLOT_C ? SLOT_C : p.slot_bytes;
    uint2 *s_desc = reinterpret_cast<uint2 *>(ln_smem);
    for (int i = threadIdx.x; i < R; i += blockDim.x) s_desc[i] = p.desc[i];
    // per warp: stage 0 (lists a, b) | out (a, b) | stage 1 (a, b) | X | PF4 | three mbarriers
    uint8_t *wbase = ln_smem + p.desc_bytes + (size_t)wid * p.warp_bytes;
    uint8_t *outA8 = wbase + 2 * (size_t)SLOT, *outB8 = outA8 + SLOT;
    uint16_t *outA = reinterpret_cast<uint16_t *>(outA8) + lane, *outB = reinterpret_cast<uint16_t *>(outB8) + lane;
    uint8_t *X8 = wbase + 6 * (size_t)SLOT;
    uint32_t *X = reinterpret_cast<uint32_t *>(X8) + lane;
    uint16_t *PF = reinterpret_cast<uint16_t *>(X8 + p.x_bytes) + lane;
    uint64_t *bars = reinterpret_cast<uint64_t *>(wbase + p.warp_bytes - 32);
    if (lane == 0) { mbar_init(smem_u32(&bars[0]), 1); mbar_init(smem_u32(&bars[1]), 1); mbar_init(smem_u32(&bars[2]), 1); }
    fence_mbar_init();
    fence_proxy_async();
    __syncthreads();

    uint32_t it = 0;                                        // trades issued so far by this warp: stage = it & 1, parity = (it >> 1) & 1
    uint32_t xit = 0;                                       // loads of X so far: parity = xit & 1
    for (int g = blockIdx.x * WPC + wid; g < p.n_groups; g += gridDim.x * WPC) {
        const uint64_t c_base = p.chain0 + (uint64_t)g * LN_W;
        const uint64_t my_chain = c_base + (uint64_t)lane;
        uint8_t *state = p.states + (size_t)g * p.gsb;
        if (p.seg_off == 0) {
            if (p.src_groups) {                             // the 32 start states are one block of the pool: copy it
                const uint4 *src = reinterpret_cast<const uint4 *>(p.src_states + (size_t)((c_base / LN_W) % (uint64_t)p.n_src) * p.rows_bytes);
                uint4 *dst = reinterpret_cast<uint4 *>(state);
                const int n16 = (int)(p.rows_bytes >> 4);
#pragma unroll 8
                for (int i = lane; i < n16; i += 32) dst[i] = __ldg(src + i);
            } else {
                const uint8_t *src = p.src_states + (size_t)(my_chain % (uint64_t)p.n_src) * (p.rows_bytes / LN_W);
                ln_load_state(state, lane, s_desc, R, nw, src);
            }
        }
        fence_proxy_async();                                // the bulk copies below read what the lanes just wrote
        __syncwarp();

        for (int j = 0; j < p.seg_len; ++j) {
            const uint64_t round = (uint64_t)p.seg_off + (uint64_t)j;
            const uint64_t pid = c_base * (uint64_t)p.chain_len + round;            // PAIR stream of the group (sdx_plan.cuh: pair_id)
            if (pid >= p.null_hi) break;                                            // no lane has a requested null from here on
            const uint64_t null_id = my_chain * (uint64_t)p.chain_len + round;
            unsigned int applied = 0;

            auto issue = [&](uint32_t st, uint2 dA, uint2 dB) {                     // lane 0: the lists of one trade -> stage st
                const uint32_t bytesA = ln_stage_bytes(dA), bytesB = ln_stage_bytes(dB);
                const uint32_t bar = smem_u32(&bars[st]);
                uint8_t *sA = wbase + (size_t)(st ? 4 : 0) * SLOT;
                mbar_expect_tx(bar, bytesA + bytesB);
                if (bytesA) bulk_g2s(smem_u32(sA), state + dA.x, bytesA, bar);
                if (bytesB) bulk_g2s(smem_u32(sA + SLOT), state + dB.x, bytesB, bar);
            };
            auto issue_x = [&](uint2 dA, uint2 dB) {                                // lane 0: the bit-packed row of a bit-packed x list trade -> X
                const uint32_t bar = smem_u32(&bars[2]);
                mbar_expect_tx(bar, (uint32_t)nw * 128u);
                bulk_g2s(smem_u32(X8), state + (ln_is_bits(dA) ? dA.x : dB.x), (uint32_t)nw * 128u, bar);
            };
            auto make_entries = [&](uint32_t t0, uint32_t prev) -> uint32_t { return ln_make_entries(&p.keys, pid, t0, prev, (uint32_t)T, (uint32_t)R); };

            uint32_t pr_cur = 0, pr_next = 0;
            bool x_busy = false;                                                    // the latest bulk store reads X
            if (T > 0) {
                pr_cur = make_entries(0, 0);
                pr_next = make_entries(32, pr_cur);
                // prologue: trades 0 and 1 (nothing is in flight: the pipeline is drained between rounds)
                const uint32_t e0 = __shfl_sync(FULL, pr_cur, 0), e1 = __shfl_sync(FULL, pr_cur, 1);
                if (lane == 0) {
                    const uint2 d0a = s_desc[e0 & 0x3fffu], d0b = s_desc[(e0 >> 14) & 0x3fffu];
                    issue(it & 1u, d0a, d0b);
                    if (ln_is_bits(d0a) != ln_is_bits(d0b)) issue_x(d0a, d0b);
                    if (T > 1 && !((e1 >> 28) & LN_FLAG_PREV)) issue((it + 1) & 1u, s_desc[e1 & 0x3fffu], s_desc[(e1 >> 14) & 0x3fffu]);
                }
            }
            for (int t = 0; t < T; ++t, ++it) {
                if (t > 0 && (t & 31) == 0) { pr_cur = pr_next; pr_next = make_entries((uint32_t)t + 32, pr_cur); }
                const uint32_t ent = __shfl_sync(FULL, pr_cur, t & 31);
                const uint32_t a = ent & 0x3fffu, b = (ent >> 14) & 0x3fffu, fl = ent >> 28;
                const uint2 dA = s_desc[a], dB = s_desc[b];
                const bool ba = ln_is_bits(dA), bb = ln_is_bits(dB);
                const bool use_x = ba != bb;
                const uint32_t st = it & 1u, par = (it >> 1) & 1u;
                if (t > 0 && (fl & LN_FLAG_PREV)) {                                 // could not be prefetched: the trade before wrote one of its lists
                    if (lane == 0) { bulk_wait<0>(); issue(st, dA, dB); }
                }
                mbar_wait(smem_u32(&bars[st]), par);
                if (use_x) { mbar_wait(smem_u32(&bars[2]), xit & 1u); ++xit; }

                uint8_t *sA8 = wbase + (size_t)(st ? 4 : 0) * SLOT;
                uint16_t *inA = reinterpret_cast<uint16_t *>(sA8) + lane, *inB = reinterpret_cast<uint16_t *>(sA8 + SLOT) + lane;
                // a bit-packed x list trade keeps RK in the list buffer the bit-packed row does not use and PK in its output
                // buffer: that one must not be under a bulk store any more
                if (use_x) {
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                }
                const LnRng rng{&p.keys, null_id, (uint32_t)t};
                uint32_t kmax = 0;
                LnTrade<M64> tr;
                ln_trade_phase1<M64>(tr, dA, dB, inA, inB, outA, outB, state, p.c_off, lane, X, PF, nw, rng, kmax);
                // the output buffers are free once the previous bulk store has read them; every store but the latest is complete
                if (lane == 0) { bulk_wait<1>(); bulk_wait_read<0>(); }
                __syncwarp();
                if (kmax != 0) {
                    ln_trade_phase2<M64>(tr, dA, dB, inA, inB, outA, outB, state, p.c_off, lane, X, PF, nw, kmax);
                    applied += tr.deal.go ? 1u : 0u;
                }
                fence_proxy_async_smem();                   // bulk stores read, and the next bulk load overwrites, what the lanes just wrote
                if (ba && bb) fence_proxy_async();          // two bit-packed rows were rewritten in place (generic proxy): a later bulk load of one must see it
                __syncwarp();
                // the rows of trade t + 1 / t + 2 that are fetched now
                const uint32_t u1 = (uint32_t)t + 1, u2 = (uint32_t)t + 2;
                const uint32_t e1 = __shfl_sync(FULL, (u1 >> 5) == ((uint32_t)t >> 5) ? pr_cur : pr_next, u1 & 31);
                const uint32_t e2 = __shfl_sync(FULL, (u2 >> 5) == ((uint32_t)t >> 5) ? pr_cur : pr_next, u2 & 31);
                if (lane == 0) {
                    if (kmax != 0) {
                        if (!ba) bulk_s2g(state + dA.x, smem_u32(outA8), ln_stage_bytes(dA));
                        else if (use_x) bulk_s2g(state + dA.x, smem_u32(X8), (uint32_t)nw * 128u);
                        if (!bb) bulk_s2g(state + dB.x, smem_u32(outB8), ln_stage_bytes(dB));
                        else if (use_x) bulk_s2g(state + dB.x, smem_u32(X8), (uint32_t)nw * 128u);
                    }
                    bulk_commit();                          // one group per trade (possibly empty): group index == trade index
                    x_busy = use_x && kmax != 0;
                    // the bit-packed row of trade t + 1 -> X (one trade ahead: X is a single buffer)
                    if (u1 < (uint32_t)T) {
                        const uint2 d1a = s_desc[e1 & 0x3fffu], d1b = s_desc[(e1 >> 14) & 0x3fffu];
                        if (ln_is_bits(d1a) != ln_is_bits(d1b)) {
                            if ((e1 >> 28) & 3u) bulk_wait<0>();            // shares a row with trade t or t - 1: their stores must have landed
                            else if (x_busy) bulk_wait_read<0>();           // the store of this trade still reads X
                            issue_x(d1a, d1b);
                        }
                    }
                    // the lists of trade t + 2 -> the stage that just became free
                    if (u2 < (uint32_t)T && !((e2 >> 28) & LN_FLAG_PREV)) {
                        if ((e2 >> 28) & LN_FLAG_NEAR) bulk_wait<0>();
                        issue(st, s_desc[e2 & 0x3fffu], s_desc[(e2 >> 14) & 0x3fffu]);
                    }
                }
            }
            // drain: every row is back in global memory before the lanes read their nulls
            if (lane == 0) bulk_wait<0>();
            __syncwarp();
            fence_proxy_async();

            // ---- outputs -----------------------------------------------------
            const bool emit = null_id >= p.null_lo && null_id < p.null_hi;
            if (emit && p.applied) p.applied[null_id - p.null_lo] = applied;
            if (emit && p.out_rows) ln_emit_rows(state, lane, s_desc, R, nw, p.out_rows + (size_t)(null_id - p.null_lo) * R * nw);
            if (p.sc.n_profiles > 0) {
                double *PP = reinterpret_cast<double *>(state + p.pp_off) + lane, *PN = PP + 32 * LN_W;
                for (int pr = 0; pr < p.sc.n_profiles; ++pr) {
                    double W = 0.0;
                    if (emit)
                        W = ln_score(state, lane, s_desc, p.sc.modules + (size_t)pr * p.sc.k, p.sc.k, p.sc.w + (size_t)pr * p.sc.S,
                                     p.sc.mask + (size_t)pr * p.sc.wprS, p.sc.S, PP, PN);
                    const unsigned ge = __ballot_sync(FULL, emit && W >= p.sc.z_obs[pr]);
                    if (lane == 0 && ge) atomicAdd(&p.sc.n_ge[pr], (unsigned long long)__popc(ge));
                    if (emit && p.sc.null_scores) p.sc.null_scores[(size_t)pr * p.sc.n_total + (null_id - p.null_lo)] = W;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
void sdx_lanes_invalidate(sdx_ctx *c, bool matrix_changed) {
    if (matrix_changed) { c->ln_state = 0; c->ln_geo_valid = false; }
    c->ln_pool_valid = false;
}

// d_states: n bit-packed states (stride 0: one state, n copies) -> groups of 32 (interleave) or compact states
static int ln_run_compact(sdx_ctx *c, const uint32_t *d_states, size_t src_stride, int n, bool interleave, uint8_t *d_out) {
    const int64_t jobs = (int64_t)n * c->R;
    const size_t bytes = interleave ? (size_t)((n + 31) / 32) * c->ln_rows_bytes : (size_t)n * (c->ln_rows_bytes / LN_W);
    SDX_CUDA(c, cudaMemsetAsync(d_out, 0, bytes, c->stream));
    k_ln_compact<<<(unsigned)((jobs + 7) / 8), 256, 0, c->stream>>>(d_states, src_stride, n, c->R, c->RS, c->wprL, c->d_ln_desc,
                                                                    interleave ? c->ln_rows_bytes : c->ln_rows_bytes / LN_W, interleave ? 1 : 0, d_out);
    c->timing.launches++; c->total_launches++;
    SDX_CUDA(c, cudaGetLastError());
    return SDX_OK;
}

// Lane layout of the uploaded matrix (once per matrix): which rows are lists, where every row sits in a group state.
static int ln_layout(sdx_ctx *c) {
    if (c->ln_state != 0) return SDX_OK;
    c->ln_state = -1;
    const int R = c->R, nw = c->wprL;
    if (R < 2 || R > 16384 || c->L > 65535 || c->knobs.lanes == 0) return SDX_OK;
    std::vector<int> size((size_t)R);
    SDX_CUDA(c, cudaMemcpyAsync(size.data(), c->d_rowsize, sizeof(int) * R, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    int ts = ln_default_ts(nw);
    if (c->knobs.lane_ts > 0) ts = std::min(c->knobs.lane_ts, LN_TS_MAX);
    std::vector<uint2> desc((size_t)R);
    const LnLayout lay = ln_build_layout(size.data(), R, nw, ts, desc.data());
    if (lay.x_bytes > 16 * 1024) return SDX_OK;          // rows too long for the per-warp buffer of a bit-packed row: k_trade runs this matrix
    c->ln_ts = lay.ts; c->ln_slot_bytes = lay.slot_bytes; c->ln_has_bits = lay.n_bits > 0; c->ln_gsb = lay.gsb;
    c->ln_rows_bytes = lay.rows_bytes; c->ln_c_off = lay.c_off; c->ln_pp_off = lay.pp_off; c->ln_pf_bytes = lay.pf_bytes; c->ln_x_bytes = lay.x_bytes;
    SDX_TRY(sdx_reserve(c, c->d_ln_desc, c->cap_ln_desc, sizeof(uint2) * (size_t)R));
    SDX_TRY(sdx_reserve(c, c->d_ln_obs, c->cap_ln_obs, (size_t)lay.rows_bytes));
    SDX_CUDA(c, cudaMemcpyAsync(c->d_ln_desc, desc.data(), sizeof(uint2) * R, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));       // desc goes out of scope
    SDX_TRY(ln_run_compact(c, c->d_trade_obs, 0, LN_W, true, c->d_ln_obs));
    c->ln_state = 1;
    return SDX_OK;
}

struct LnGeom { int wpc, grid, desc_bytes, warp_bytes; size_t smem; bool m64; const void *kern; };

#define LN_SLOT_FIXED 1088
static const void *ln_kernel(bool m64, bool wide, int slot_bytes) {
    if (m64) return wide ? (const void *)k_trade_lanes<true, 128, 0> : (const void *)k_trade_lanes<true, 80, 0>;
    if (slot_bytes == LN_SLOT_FIXED) return wide ? (const void *)k_trade_lanes<false, 128, LN_SLOT_FIXED> : (const void *)k_trade_lanes<false, 80, LN_SLOT_FIXED>;
    return wide ? (const void *)k_trade_lanes<false, 128, 0> : (const void *)k_trade_lanes<false, 80, 0>;
}

// Warps per CTA and grid for n_groups groups: as many resident warps per SM as shared memory and registers allow, but
// never fewer CTAs than SMs when the groups are few (one warp runs one group at a time).
static bool ln_geometry(sdx_ctx *c, int64_t n_groups, LnGeom *g) {
    if (c->ln_geo_groups == n_groups && c->ln_geo_valid) { *g = *reinterpret_cast<const LnGeom *>(c->ln_geo); return true; }
    g->m64 = 2 * c->ln_ts > 32;                          // the chosen ranks of a list x list trade need a register pair
    int wb = 6 * c->ln_slot_bytes + c->ln_x_bytes + c->ln_pf_bytes + 32;
    wb = ((wb + 127) / 128) * 128;
    g->warp_bytes = wb;
    g->desc_bytes = ((c->R * 8 + 127) / 128) * 128;
    const size_t budget = c->smem_optin > 2048 ? c->smem_optin - 1024 : 0;
    static const int cand[] = {1, 2, 3, 4, 6, 8, 11, 12, 16, 20, 22};
    int best = 0, best_grid = 0, best_cap = 0;
    double best_cost = 1e300;
    const void *best_kern = nullptr;
    for (int wide = 1; wide >= 0; --wide) {              // the 128-register variant first: it wins ties
        const void *kern = ln_kernel(g->m64, wide != 0, c->ln_slot_bytes);
        for (int wpc : cand) {
            if (c->knobs.lane_wpc > 0 && wpc != c->knobs.lane_wpc) continue;
            const size_t need = (size_t)g->desc_bytes + (size_t)wpc * wb;
            if (need > budget || wpc * 32 > LN_MAX_THREADS) continue;
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need) != cudaSuccess) { cudaGetLastError(); continue; }
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, wpc * 32, need) != cudaSuccess || nb < 1) { cudaGetLastError(); continue; }
            const int cap = nb * wpc;                    // resident warps per SM
            if (c->knobs.lane_debug) fprintf(stderr, "[lanes] %d regs, wpc %d: %zu B, %d CTAs / SM -> %d warps / SM\n", wide ? 128 : 80, wpc, need, nb, cap);
            const int64_t ctas = (n_groups + wpc - 1) / wpc;
            const int64_t slots = (int64_t)nb * c->sm_count;
            // time ~ waves x round time of a warp that shares its SM with `busiest - 1` others.  Measured at C1 (ms per
            // round of 1 920 trades): 6.8 alone, 7.7 / 7.9 / 8.8 / 13.7 with 5 / 7 / 11 / 22 warps per SM ~ 6.5 + 0.33 w.
            const int64_t busiest = ctas <= slots ? (ctas + c->sm_count - 1) / c->sm_count * wpc : cap;
            const double cost = (double)((ctas + slots - 1) / slots) * (6.5 + 0.33 * (double)busiest) + 1e-5 * wpc;
            if (cost < best_cost - 1e-9) {
                best_cost = cost; best = wpc; best_cap = cap; best_kern = kern;
                best_grid = (int)std::min<int64_t>(ctas, slots);
            }
        }
    }
    if (best == 0) return false;
    g->wpc = best; g->grid = best_grid; g->kern = best_kern;
    if (c->knobs.lane_debug) fprintf(stderr, "[lanes] %lld groups -> %d warps per CTA, grid %d\n", (long long)n_groups, best, best_grid);
    g->smem = (size_t)g->desc_bytes + (size_t)best * wb;
    static_assert(sizeof(LnGeom) <= sizeof(c->ln_geo), "LnGeom cache");
    memcpy(c->ln_geo, g, sizeof(LnGeom));                 // the occupancy queries above are per (matrix, number of groups)
    c->ln_geo_groups = n_groups; c->ln_geo_valid = true;
    return true;
}

bool sdx_lanes_applicable(sdx_ctx *c, int64_t n_chains, int64_t T, bool fused_scoring) {
    if (c->pair_group != LN_W || c->knobs.lanes == 0) return false;
    if (ln_layout(c) != SDX_OK || c->ln_state != 1) return false;
    if (fused_scoring && c->rows_are_samples) return false;      // the fused scorer reads feature rows
    if (T > 0x7fffffff) return false;
    // A warp runs the 1 920 trades of a C1 round one after the other (6.8 ms alone), k_trade runs a null level-parallel on
    // a CTA (0.3 ms): below ~4 groups per SM the request is latency bound and k_trade is faster (C1, measured: 12 500
    // nulls 5.6 vs 7.3 ms, 25 000 nulls 10.4 vs 8.3 ms).  Both kernels produce the same nulls.
    const int64_t min_chains = c->knobs.lane_min_chains > 0 ? c->knobs.lane_min_chains : (int64_t)4 * LN_W * c->sm_count;
    if (n_chains < min_chains) return false;
    LnGeom g;
    return ln_geometry(c, (n_chains + LN_W - 1) / LN_W, &g);
}

// sdx_curveball.cu
int sdx_score_setup(sdx_ctx *c, const HostScore *hs, int64_t n_nulls, bool colcache, ScoreArgs *sc, LaunchTimer *timer);
int sdx_score_finish(sdx_ctx *c, const HostScore *hs, int64_t n_nulls, const ScoreArgs &sc);

int sdx_run_nulls_lanes(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t T, int64_t chain_len,
                        uint32_t *out_rows, int64_t *applied, const HostScore *hs) {
    const int R = c->R, nw = c->wprL;
    int rc;
    // start states: the burn-in pool (built by k_trade, converted once) or the observed matrix
    const uint8_t *src = c->d_ln_obs;
    int n_src = 1, src_groups = 1;
    if (c->burn > 0) {
        const bool was_valid = c->pool_valid && c->pool_seed == seed && c->pool_T == T;
        if ((rc = sdx_ensure_pool(c, seed, T))) return rc;
        const bool groups = c->pool_size % LN_W == 0;
        if (!was_valid || !c->ln_pool_valid) {
            const size_t bytes = groups ? (size_t)(c->pool_size / LN_W) * c->ln_rows_bytes : (size_t)c->pool_size * (c->ln_rows_bytes / LN_W);
            SDX_TRY(sdx_reserve(c, c->d_ln_pool, c->cap_ln_pool, bytes));
            SDX_TRY(ln_run_compact(c, c->d_pool, (size_t)c->R * c->RS, c->pool_size, groups, c->d_ln_pool));
            c->ln_pool_valid = true; c->ln_pool_groups = groups;
        }
        src = c->d_ln_pool;
        src_groups = c->ln_pool_groups ? 1 : 0;
        n_src = src_groups ? c->pool_size / LN_W : c->pool_size;
    }
    LaunchTimer timer(c);
    ScoreArgs sc{};
    if (hs && (rc = sdx_score_setup(c, hs, n_nulls, false, &sc, &timer))) return rc;

    float trade_ms = 0;
    int n_launch = 0;
    if (n_nulls > 0) {
        const uint64_t null_hi = null0 + (uint64_t)n_nulls;
        const uint64_t chain_first = null0 / (uint64_t)chain_len, chain_last = (null_hi - 1) / (uint64_t)chain_len;
        const uint64_t g_first = chain_first / LN_W, g_last = chain_last / LN_W;
        const int64_t n_groups = (int64_t)(g_last - g_first + 1);
        const size_t out_per = (size_t)R * nw;
        void *p_out = nullptr, *p_app = nullptr, *p_states = nullptr;
        bool out_is_device = false;
        if (out_rows) {
            cudaPointerAttributes at{};
            out_is_device = cudaPointerGetAttributes(&at, out_rows) == cudaSuccess && at.type == cudaMemoryTypeDevice;
            cudaGetLastError();
        }
        if (applied) {
            SDX_TRY(sdx_scratch(c, SL_APPLIED, sizeof(unsigned long long) * n_nulls, &p_app));
            SDX_CUDA(c, cudaMemsetAsync(p_app, 0, sizeof(unsigned long long) * n_nulls, c->stream));
        }
        // groups per launch: bounded by the working states (3 GB) and, when the nulls are returned to the host, by the staging buffer (1 GB)
        int64_t groups_per_launch = std::max<int64_t>(1, (int64_t)(((size_t)3 << 30) / c->ln_gsb));
        if (out_rows && !out_is_device) {
            const int64_t by_out = std::max<int64_t>(1, (int64_t)(((size_t)1 << 30) / (sizeof(uint32_t) * out_per * LN_W * (size_t)chain_len + 1)));
            groups_per_launch = std::min(groups_per_launch, by_out);
        }
        if (groups_per_launch > n_groups) groups_per_launch = n_groups;
        SDX_TRY(sdx_scratch(c, SL_LN_STATE, (size_t)groups_per_launch * c->ln_gsb, &p_states));
        const int seg_cap = (int)std::min<int64_t>(chain_len, 64);
        const int64_t n_segs = (chain_len + seg_cap - 1) / seg_cap;
        for (int64_t gb = 0; gb < n_groups; gb += groups_per_launch) {
            const int64_t ng = std::min(groups_per_launch, n_groups - gb);
            LnGeom geo;
            if (!ln_geometry(c, ng, &geo)) return sdx_fail(c, SDX_ERR_UNSUPPORTED, "lane kernel: shared memory layout does not fit");
            SDX_CUDA(c, cudaFuncSetAttribute(geo.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)geo.smem));
            // nulls of this batch of groups: chains are contiguous in the null index
            const uint64_t b_lo = std::max<uint64_t>(null0, (g_first + (uint64_t)gb) * LN_W * (uint64_t)chain_len);
            const uint64_t b_hi = std::min<uint64_t>(null_hi, (g_first + (uint64_t)(gb + ng)) * LN_W * (uint64_t)chain_len);
            uint32_t *d_out = nullptr;                   // rows of null b_lo
            if (out_rows) {
                if (out_is_device) d_out = out_rows + (size_t)(b_lo - null0) * out_per;
                else { SDX_TRY(sdx_scratch(c, SL_OUT, sizeof(uint32_t) * out_per * (size_t)(b_hi - b_lo), &p_out)); d_out = (uint32_t *)p_out; }
                SDX_CUDA(c, cudaMemsetAsync(d_out, 0, sizeof(uint32_t) * out_per * (size_t)(b_hi - b_lo), c->stream));
            }
            for (int64_t sg = 0; sg < n_segs; ++sg) {
                LaneArgs a{};
                a.keys = make_keys(seed);
                a.desc = c->d_ln_desc; a.R = R; a.T = (int)T; a.wprL = nw;
                a.slot_bytes = c->ln_slot_bytes; a.desc_bytes = geo.desc_bytes; a.warp_bytes = geo.warp_bytes; a.x_bytes = c->ln_x_bytes;
                a.gsb = c->ln_gsb; a.rows_bytes = c->ln_rows_bytes; a.c_off = c->ln_c_off; a.pp_off = c->ln_pp_off;
                a.states = (uint8_t *)p_states;
                a.chain0 = (g_first + (uint64_t)gb) * LN_W; a.chain_len = chain_len;
                a.seg_off = sg * seg_cap;
                a.seg_len = (int)std::min<int64_t>(seg_cap, chain_len - sg * seg_cap);
                a.null_lo = null0; a.null_hi = null_hi; a.n_groups = (int)ng;
                a.src_states = src; a.n_src = n_src; a.src_groups = src_groups;
                a.out_rows = d_out ? d_out - (size_t)(b_lo - null0) * out_per : nullptr;      // the kernel indexes by null_id - null_lo
                a.applied = (unsigned long long *)p_app; a.sc = sc;
                cudaEventRecord(c->ev[3], c->stream);
                void *kargs[] = {&a};
                cudaError_t e = cudaLaunchKernel(geo.kern, dim3((unsigned)geo.grid), dim3((unsigned)geo.wpc * 32), kargs, geo.smem, c->stream);
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e != cudaSuccess) return sdx_fail(c, SDX_ERR_CUDA, std::string("lane kernel launch failed: ") + cudaGetErrorString(e));
                cudaEventRecord(c->ev[4], c->stream);
                timer.count();
                ++n_launch;
                SDX_CUDA(c, cudaStreamSynchronize(c->stream));
                float ms = 0;
                cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]);
                trade_ms += ms;
            }
            if (out_rows && !out_is_device && b_hi > b_lo) {
                SDX_CUDA(c, cudaMemcpyAsync(out_rows + (size_t)(b_lo - null0) * out_per, d_out, sizeof(uint32_t) * out_per * (size_t)(b_hi - b_lo),
                                            cudaMemcpyDeviceToHost, c->stream));
                SDX_CUDA(c, cudaStreamSynchronize(c->stream));
            }
        }
        if (applied) {
            std::vector<unsigned long long> tmp((size_t)n_nulls);
            SDX_CUDA(c, cudaMemcpyAsync(tmp.data(), p_app, sizeof(unsigned long long) * n_nulls, cudaMemcpyDeviceToHost, c->stream));
            SDX_CUDA(c, cudaStreamSynchronize(c->stream));
            for (int64_t i = 0; i < n_nulls; ++i) applied[i] = (int64_t)tmp[i];
        }
    }
    if (hs && (rc = sdx_score_finish(c, hs, n_nulls, sc))) return rc;
    rc = timer.finish();
    c->timing.plan_ms = 0; c->timing.trade_ms = trade_ms; c->timing.trade_launches = n_launch;
    return rc;
}
