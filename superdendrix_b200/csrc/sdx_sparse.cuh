// sdx_sparse.cuh — Curveball trades on a compact shared-memory form of the null matrix.
//
// The feature matrix of the path is very sparse (C0: 4 054 ones in 384 x 770), and Curveball preserves every row
// sum, so a row with s carriers never needs more than s index slots.  Rows with at most SP_TS carriers are kept
// as sorted 16-bit index lists in fixed slots; the few dense rows stay bit-packed.  A null then occupies ~8 KB of
// shared memory instead of 43 KB, ONE WARP runs a whole null (no CTA barrier between levels), and ~19 nulls are
// resident per SM instead of 5.
//
//   list x list trades     one trade per THREAD: two merge walks over the sorted lists (count, then deal)
//   anything with a dense  groups of 4 lanes on bit-packed rows (trade_row_ptrs); a list row is expanded into a
//   row                    scratch bitmap first and compacted back afterwards
//
// The canonical algorithm (ranks in ascending index order, Floyd picks, DEAL sub-stream per pick) is the one of
// sdx_trade.cuh, so outputs are bit-identical to the bit-packed kernel and to the oracle.
#pragma once
#include "sdx_trade.cuh"

#define SP_TS 16                       // rows with at most this many carriers are index lists
#define SP_G 4                         // lanes per trade on the bit-packed path
#define SP_V 8                         // words per lane: rows of at most 32 words (1 024 columns)
#define SP_GPW (32 / SP_G)

// row descriptor: bit 31 = bit-packed row, bits 16..30 = row size, bits 0..15 = list offset (in entries) or dense slot
__host__ __device__ __forceinline__ uint32_t sp_desc(bool dense, int size, int off) {
    return (dense ? 0x80000000u : 0u) | ((uint32_t)size << 16) | (uint32_t)off;
}
__device__ __forceinline__ bool sp_is_dense(uint32_t d) { return (d >> 31) != 0; }
__device__ __forceinline__ int sp_size(uint32_t d) { return (int)((d >> 16) & 0x7fffu); }
__device__ __forceinline__ int sp_off(uint32_t d) { return (int)(d & 0xffffu); }

struct SparseView {                    // one null in shared memory
    uint32_t *bits;                    // n_dense x RS words
    uint16_t *lists;                   // list pool
    const uint32_t *desc;              // R descriptors (global, read-only)
    int RS;
    __device__ __forceinline__ uint32_t *dense_row(uint32_t d) const { return bits + (size_t)sp_off(d) * RS; }
    __device__ __forceinline__ uint16_t *list_row(uint32_t d) const { return lists + sp_off(d); }
};

// row -> bit-packed copy in dst (RS words).  All 32 lanes; RS <= 32.
__device__ __forceinline__ void sp_row_to_bitmap(const SparseView &v, uint32_t d, uint32_t *dst, int lane) {
    if (sp_is_dense(d)) {
        if (lane < v.RS) dst[lane] = v.dense_row(d)[lane];
    } else {
        if (lane < v.RS) dst[lane] = 0u;
        __syncwarp();
        if (lane < sp_size(d)) { const uint32_t e = v.list_row(d)[lane]; atomicOr(&dst[e >> 5], 1u << (e & 31)); }
    }
    __syncwarp();
}

// Expand the list row of a group into its scratch bitmap / compact it back.  Called by the SP_G lanes of a group
// (all 32 lanes of the warp execute it together; `mine` selects the groups that really convert).
__device__ __forceinline__ void sp_group_expand(const uint16_t *list, int size, uint32_t *dst, int RS, int sub, bool mine) {
    if (mine) {
        for (int i = sub; i < RS; i += SP_G) dst[i] = 0u;
    }
    __syncwarp();
    if (mine) {
        for (int i = sub; i < size; i += SP_G) { const uint32_t e = list[i]; atomicOr(&dst[e >> 5], 1u << (e & 31)); }
    }
    __syncwarp();
}

__device__ __forceinline__ void sp_group_compact(const uint32_t *src, uint16_t *list, int RS, int sub, bool mine) {
    // lane `sub` owns words [sub * SP_V, sub * SP_V + SP_V): ascending lane order == ascending bit order
    uint32_t w[SP_V];
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < SP_V; ++i) {
        const int wi = sub * SP_V + i;
        w[i] = (mine && wi < RS) ? src[wi] : 0u;
        cnt += __popc(w[i]);
    }
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < SP_G; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, inc, o, SP_G);
        if (sub >= o) inc += x;
    }
    int pos = inc - cnt;
    if (mine) {
#pragma unroll
        for (int i = 0; i < SP_V; ++i) {
            uint32_t x = w[i];
            while (x) {
                list[pos++] = (uint16_t)((sub * SP_V + i) * 32 + __ffs(x) - 1);
                x &= x - 1;
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// list x list trade, one per thread.  A (na entries) and B (nb entries) are sorted ascending; na + nb <= 2 * SP_TS
// = 32, so the rank mask of the symmetric difference fits one register.  outA / outB: this thread's scratch
// (SP_TS + 1 entries each, stride `ostride` between entries).  All 32 lanes call it together.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool sp_trade_lists(uint16_t *A, int na, uint16_t *B, int nb, bool act, uint32_t t, uint64_t null_id,
                                               const PhiloxKeys &keys, uint16_t *outA, uint16_t *outB, int ostride) {
    constexpr unsigned FULL = 0xffffffffu;
    // first walk: only the number of common elements is needed (|a\\b| = na - common, |b\\a| = nb - common)
    uint32_t La = 0, Lb = 0;
    if (act) {
        const uint16_t *pa = A, *pb = B, *ea = A + na, *eb = B + nb;
        uint32_t common = 0;
        while (pa < ea && pb < eb) {
            const uint32_t x = *pa, y = *pb;
            common += x == y;
            pa += x <= y; pb += y <= x;
        }
        La = na - common; Lb = nb - common;
    }
    const bool go = La != 0 && Lb != 0;
    const bool small_is_a = La <= Lb;
    const uint32_t n = La + Lb;
    const uint32_t k = go ? (small_is_a ? La : Lb) : 0u;
    const uint32_t kmax = __reduce_max_sync(FULL, k);
    if (kmax == 0) return false;                        // warp-uniform
    uint32_t cm = 0;                                     // chosen ranks
#pragma unroll 1
    for (uint32_t blk = 0; blk * 4 < kmax; ++blk) {
        const uint4 xb = stream_block(keys, SDX_STREAM_DEAL, null_id, t, blk);
        const uint32_t x[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const uint32_t i = blk * 4 + c;
            if (i < k) {
                const uint32_t m = n - k + i + 1;       // Floyd: j = m - 1, r uniform in [0, j]
                const uint64_t prod = (uint64_t)x[c] * m;
                uint32_t r = (uint32_t)(prod >> 32);
                if ((uint32_t)prod < m) r = deal_below_slow(&keys, null_id, t, i, m, x[c]);
                if ((cm >> r) & 1u) r = m - 1;
                cm |= 1u << r;
            }
        }
    }
    if (go) {
        // second walk: elements in ascending order; common ones go to both rows, the rank-th element of the symmetric
        // difference goes to the smaller side when its rank was chosen, to the other side otherwise
        // (every element is stored to both scratch lists; only the list that takes it advances)
        const uint16_t *pa = A, *pb = B, *ea = A + na, *eb = B + nb;
        uint16_t *oa = outA, *ob = outB;
        while (pa < ea || pb < eb) {
            const uint32_t x = pa < ea ? *pa : 0xffffffffu, y = pb < eb ? *pb : 0xffffffffu;
            const uint32_t e = x < y ? x : y;
            const bool common = x == y;
            pa += x <= y; pb += y <= x;
            const bool picked = cm & 1u;
            cm >>= common ? 0 : 1;
            const bool to_a = common || picked == small_is_a;
            const bool to_b = common || picked != small_is_a;
            *oa = (uint16_t)e; *ob = (uint16_t)e;
            oa += to_a ? ostride : 0; ob += to_b ? ostride : 0;
        }
        for (int q = 0; q < na; ++q) A[q] = outA[q * ostride];
        for (int q = 0; q < nb; ++q) B[q] = outB[q * ostride];
    }
    return go;
}
