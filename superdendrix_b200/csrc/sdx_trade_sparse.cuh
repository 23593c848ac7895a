// sdx_trade_sparse.cuh — k_to_sparse / k_trade_sparse: experimental one-warp-per-null kernel on the sparse layout (SDX_TRADE_SPARSE=1).
// Included by sdx_curveball.cu only.
#pragma once
#include "sdx_sparse.cuh"
#include "sdx_trade_kernel.cuh"

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

