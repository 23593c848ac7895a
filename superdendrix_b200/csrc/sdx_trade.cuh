// sdx_trade.cuh — device code of the Curveball trade and of the set objective.
//
// Reference semantics: src/curveball.py:25-35 (one trade), src/superdendrix.py:387-391
// (SuperW).  Canonical random streams: DESIGN.md §3 / oracle/sdx_oracle.c.
#pragma once
#include "sdx_common.cuh"

// ---------------------------------------------------------------------------
// DEAL stream (oracle/sdx_oracle.c: deal_draw): draw `attempt`
// of pick i of trade t = word (i & 3) of block ((i >> 2) | (attempt << 16)).
// ---------------------------------------------------------------------------
// Rare branch of Lemire's method (taken with probability m / 2^32 per pick).
// Out of line so that the modulo is not speculated into the hot loop.
static __device__ __noinline__ uint32_t deal_below_slow(const PhiloxKeys *keys, uint64_t null_id, uint32_t t, uint32_t i,
                                                 uint32_t m, uint32_t x) {
    uint64_t prod = (uint64_t)x * m;
    uint32_t lo = (uint32_t)prod;
    const uint32_t thr = (0u - m) % m;
    uint32_t attempt = 0;
    while (lo < thr) {
        ++attempt;
        const uint4 b = stream_block(*keys, SDX_STREAM_DEAL, null_id, t, (i >> 2) | (attempt << 16));
        prod = (uint64_t)pick4(b, i & 3) * m;
        lo = (uint32_t)prod;
    }
    return (uint32_t)(prod >> 32);
}

// ---------------------------------------------------------------------------
// One trade per group of G consecutive lanes; lane `sub` of the group holds V
// consecutive words of each of the two rows (blocked layout: lane order ==
// ascending bit order).  ALL 32 lanes of the warp call it together (groups
// without a trade pass act = false), so every shuffle uses the full mask.
//
//   a' = (a&b) | picked ,  b' = (a&b) | rest     (or the other way round)
//
// `picked` is a uniformly random k-subset, k = min(|a\b|, |b\a|), of the n
// symmetric-difference elements: Floyd's algorithm over their ranks (ascending
// bit order), one sub-stream of the DEAL stream per pick.  The chosen ranks
// are kept as a compressed n-bit mask; while the trade runs both rows live in
// registers, so the storage of row a doubles as the scratch for that mask
// (lane 0 of the group performs the picks).  Afterwards each lane expands its
// slice of the mask onto the set bits of its own words.
// Returns true when the trade was applied (src/curveball.py:29 skip rule).
// ---------------------------------------------------------------------------
// row_a / row_b point at the first word of the two rows (RS words each, 16-byte aligned).
template <int G, int V>
__device__ __forceinline__ bool trade_row_ptrs(uint32_t *row_a, uint32_t *row_b, int RS, bool act,
                                               uint32_t t, uint64_t null_id, const PhiloxKeys &keys, int sub) {
    constexpr unsigned FULL = 0xffffffffu;
    uint32_t A[V], B[V], sym[V];
    uint32_t *ra = row_a + sub * V;
    uint32_t *rb = row_b + sub * V;
    if constexpr (V % 4 == 0) {
#pragma unroll
        for (int q = 0; q < V / 4; ++q) {
            uint4 va = make_uint4(0, 0, 0, 0), vb = va;
            if (act && sub * V + 4 * q < RS) {
                va = *reinterpret_cast<const uint4 *>(ra + 4 * q);
                vb = *reinterpret_cast<const uint4 *>(rb + 4 * q);
            }
            A[4 * q] = va.x; A[4 * q + 1] = va.y; A[4 * q + 2] = va.z; A[4 * q + 3] = va.w;
            B[4 * q] = vb.x; B[4 * q + 1] = vb.y; B[4 * q + 2] = vb.z; B[4 * q + 3] = vb.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const bool in = act && sub * V + i < RS;
            A[i] = in ? ra[i] : 0u;
            B[i] = in ? rb[i] : 0u;
        }
    }

    uint32_t ca = 0, cs = 0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sym[i] = A[i] ^ B[i];
        ca += __popc(A[i] & ~B[i]);
        cs += __popc(sym[i]);
    }
    // group-wide inclusive scan of (|a\b|, |a^b|) packed in one register
    const uint32_t packed = ca | (cs << 16);
    uint32_t inc = packed;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, inc, o, G);
        if (sub >= o) inc += v;
    }
    const uint32_t tot = __shfl_sync(FULL, inc, G - 1, G);
    const uint32_t La = tot & 0xffffu, n = tot >> 16, Lb = n - La;
    const bool go = La != 0 && Lb != 0;                 // act == false gives La == Lb == 0
    const bool small_is_a = La <= Lb;
    const uint32_t k = go ? (small_is_a ? La : Lb) : 0u;
    // one reduction for both maxima: k <= n / 2 and n < 2^15, and the group with the largest n that trades has k > 0
    const uint32_t kmax = __reduce_max_sync(FULL, k);
    if (kmax == 0) return false;                        // warp-uniform
    const uint32_t nmax = __any_sync(FULL, go && n > 64u) ? 65u : 0u;   // only "does any trade of the warp need the wide mask" matters
    uint32_t off = (inc - packed) >> 16;                // rank of this lane's first symmetric-difference element
    uint32_t pick[V];

    if (nmax <= 64) {
        // ---- fast path: the compressed mask of chosen ranks fits a register pair of lane 0 ----------
        uint32_t c0 = 0, c1 = 0;
#pragma unroll 1
        for (uint32_t blk0 = 0; blk0 * 4 < kmax; blk0 += G) {
            const uint4 blk = stream_block(keys, SDX_STREAM_DEAL, null_id, t, blk0 + sub);
#pragma unroll 1
            for (int q = 0; q < G; ++q) {
                const uint32_t i0 = (blk0 + q) * 4;
                if (i0 >= kmax) break;                  // warp-uniform
                uint32_t x[4];
                x[0] = __shfl_sync(FULL, blk.x, q, G);
                x[1] = __shfl_sync(FULL, blk.y, q, G);
                x[2] = __shfl_sync(FULL, blk.z, q, G);
                x[3] = __shfl_sync(FULL, blk.w, q, G);
                if (sub == 0 && i0 < k) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t i = i0 + c;
                        if (i < k) {
                            const uint32_t m = n - k + i + 1;           // Floyd: j = m - 1, r uniform in [0, j]
                            const uint64_t prod = (uint64_t)x[c] * m;
                            uint32_t r = (uint32_t)(prod >> 32);
                            if ((uint32_t)prod < m) r = deal_below_slow(&keys, null_id, t, i, m, x[c]);
                            const uint32_t w = (r & 32u) ? c1 : c0;
                            if ((w >> (r & 31)) & 1u) r = m - 1;
                            const uint32_t bit = 1u << (r & 31);
                            if (r & 32u) c1 |= bit; else c0 |= bit;
                        }
                    }
                }
            }
        }
        c0 = __shfl_sync(FULL, c0, 0, G);
        c1 = __shfl_sync(FULL, c1, 0, G);
        // drop the ranks below this lane's first element, then peel popc(sym) bits per word
        if (off & 32u) { c0 = c1; c1 = 0u; }
        c0 = __funnelshift_r(c0, c1, off & 31u);
        c1 >>= (off & 31u);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const uint32_t c = __popc(sym[i]);
            uint32_t ch = c0 & __funnelshift_rc(0xffffffffu, 0u, 32u - c);
            c0 = __funnelshift_rc(c0, c1, c);
            c1 = (c & 32u) ? 0u : (c1 >> c);
            uint32_t m = sym[i], pk = 0;
            while (ch) {
                const uint32_t low = m & (0u - m);
                m ^= low;
                if (ch & 1u) pk |= low;
                ch >>= 1;
            }
            pick[i] = pk;
        }
    } else {
        // ---- general path: the mask lives in the storage of row a (both rows are in registers) ------
        volatile uint32_t *cm = row_a;
        const uint32_t nw = go ? (n + 31) >> 5 : 0u;
        for (uint32_t i = sub; i < nw; i += G) cm[i] = 0u;
        __syncwarp();
#pragma unroll 1
        for (uint32_t blk0 = 0; blk0 * 4 < kmax; blk0 += G) {
            const uint4 blk = stream_block(keys, SDX_STREAM_DEAL, null_id, t, blk0 + sub);
#pragma unroll 1
            for (int q = 0; q < G; ++q) {
                const uint32_t i0 = (blk0 + q) * 4;
                if (i0 >= kmax) break;                  // warp-uniform
                uint32_t x[4];
                x[0] = __shfl_sync(FULL, blk.x, q, G);
                x[1] = __shfl_sync(FULL, blk.y, q, G);
                x[2] = __shfl_sync(FULL, blk.z, q, G);
                x[3] = __shfl_sync(FULL, blk.w, q, G);
                if (sub == 0 && i0 < k) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t i = i0 + c;
                        if (i < k) {
                            const uint32_t m = n - k + i + 1;
                            const uint64_t prod = (uint64_t)x[c] * m;
                            uint32_t r = (uint32_t)(prod >> 32);
                            if ((uint32_t)prod < m) r = deal_below_slow(&keys, null_id, t, i, m, x[c]);
                            uint32_t w = cm[r >> 5];
                            if ((w >> (r & 31)) & 1u) { r = m - 1; w = cm[r >> 5]; }
                            cm[r >> 5] = w | (1u << (r & 31));
                        }
                    }
                }
            }
        }
        __syncwarp();
        // expand: this lane's words take bits [off, off + popc) of the compressed mask, in rank order
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const uint32_t c = __popc(sym[i]);
            uint32_t pk = 0;
            if (go && c) {
                const uint32_t w0 = off >> 5;
                const uint32_t lo = cm[w0];
                const uint32_t hi = (w0 + 1 < nw) ? cm[w0 + 1] : 0u;
                uint32_t ch = __funnelshift_r(lo, hi, off & 31);
                if (c < 32) ch &= (1u << c) - 1u;
                uint32_t m = sym[i];
                while (ch) {
                    const uint32_t low = m & (0u - m);
                    m ^= low;
                    if (ch & 1u) pk |= low;
                    ch >>= 1;
                }
            }
            pick[i] = pk;
            off += c;
        }
    }
    __syncwarp();                                       // all reads of the scratch are done before row a is rewritten
    if (go) {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            const uint32_t ab = A[i] & B[i], rest = sym[i] ^ pick[i];
            A[i] = ab | (small_is_a ? pick[i] : rest);
            B[i] = ab | (small_is_a ? rest : pick[i]);
        }
        if constexpr (V % 4 == 0) {
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
                if (sub * V + 4 * q < RS) {
                    *reinterpret_cast<uint4 *>(ra + 4 * q) = make_uint4(A[4 * q], A[4 * q + 1], A[4 * q + 2], A[4 * q + 3]);
                    *reinterpret_cast<uint4 *>(rb + 4 * q) = make_uint4(B[4 * q], B[4 * q + 1], B[4 * q + 2], B[4 * q + 3]);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) {
                if (sub * V + i < RS) { ra[i] = A[i]; rb[i] = B[i]; }
            }
        }
    }
    return go;
}

template <int G, int V>
__device__ __forceinline__ bool trade_rows(uint32_t *rows, int RS, bool act, uint32_t a, uint32_t b,
                                           uint32_t t, uint64_t null_id, const PhiloxKeys &keys, int sub) {
    return trade_row_ptrs<G, V>(rows + (size_t)a * RS, rows + (size_t)b * RS, RS, act, t, null_id, keys, sub);
}

// ---------------------------------------------------------------------------
// Set objective for one module, computed by one warp (all 32 lanes call it).
//   W = 2 * sum_{s in U, w_s > 0} w_s  -  sum_g sum_{s in g} |w_s|
// rows: packed matrix with row stride RS words.  rows_are_samples = false:
// row g = carriers of feature g over samples.  true: row s = features of
// sample s (the trade orientation when F >= S).  mod[k]: feature indices, -1
// = empty slot.  w: S weights; mask: packed profile-sample mask.
// fp64 reduction is a fixed butterfly, so the result is a pure function of the
// bit content (equal sets give bit-identical W).
// ---------------------------------------------------------------------------
struct SetScore {
    double W;
    double sw;            // sum of w over the union (utils/compute_individual_contribution_SELECT.py:73-75)
    int cov, tot, act;
};

__device__ __forceinline__ SetScore score_module(const uint32_t *__restrict__ rows, int RS, bool rows_are_samples,
                                                 const int32_t *__restrict__ mod, int k,
                                                 const double *__restrict__ w, const uint32_t *__restrict__ mask,
                                                 int S, int lane) {
    double pos = 0.0, neg = 0.0, sw = 0.0;
    int cov = 0, tot = 0, act = 0;
    int g[SDX_MAX_K];
#pragma unroll
    for (int j = 0; j < SDX_MAX_K; ++j) g[j] = (j < k) ? mod[j] : -1;
    if (!rows_are_samples) {
        const int nw = (S + 31) >> 5;
        for (int wi = lane; wi < nw; wi += 32) {
            const uint32_t m = mask[wi];
            uint32_t r[SDX_MAX_K];
            uint32_t U = 0;
#pragma unroll
            for (int j = 0; j < SDX_MAX_K; ++j) {
                r[j] = (g[j] >= 0) ? (rows[(size_t)g[j] * RS + wi] & m) : 0u;
                U |= r[j];
                tot += __popc(r[j]);
            }
            cov += __popc(U);
            uint32_t x = U;
            while (x) {
                const int bp = __ffs(x) - 1;
                x &= x - 1;
                const double ws = w[wi * 32 + bp];
                int mult = 0;
#pragma unroll
                for (int j = 0; j < SDX_MAX_K; ++j) mult += (r[j] >> bp) & 1;
                if (ws > 0.0) { pos += ws; ++act; }
                neg += (double)mult * fabs(ws);
                sw += ws;
            }
        }
    } else {
        // lane <- the 32 samples of mask word wi, in ascending order: the same partition and order of the fp64 sums as
        // the branch above, so both orientations (and the column cache of k_trade) give bit-identical W for equal sets
        const int nw = (S + 31) >> 5;
        for (int wi = lane; wi < nw; wi += 32) {
            uint32_t m = mask[wi];
            while (m) {
                const int bp = __ffs(m) - 1;
                m &= m - 1;
                const int s = wi * 32 + bp;
                int mult = 0;
#pragma unroll
                for (int j = 0; j < SDX_MAX_K; ++j)
                    if (g[j] >= 0) mult += (rows[(size_t)s * RS + (g[j] >> 5)] >> (g[j] & 31)) & 1u;
                if (mult) {
                    const double ws = w[s];
                    ++cov; tot += mult;
                    if (ws > 0.0) { pos += ws; ++act; }
                    neg += (double)mult * fabs(ws);
                    sw += ws;
                }
            }
        }
    }
    pos = warp_sum_f64(pos);
    neg = warp_sum_f64(neg);
    SetScore o;
    o.W = 2.0 * pos - neg;
    o.sw = warp_sum_f64(sw);
    o.cov = warp_sum_i32(cov);
    o.tot = warp_sum_i32(tot);
    o.act = warp_sum_i32(act);
    return o;
}
