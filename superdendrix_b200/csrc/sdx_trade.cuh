// sdx_trade.cuh — device code of the Curveball trade and of the set objective.
//
// Reference semantics: src/curveball.py:25-35 (one trade), src/superdendrix.py:387-391
// (SuperW).  Canonical random streams: DESIGN.md §3 / oracle/sdx_oracle.c.
#pragma once
#include "sdx_common.cuh"

#define SDX_PLAN_END 0x80000000u   // plan entry flag: last trade of its level

// ---------------------------------------------------------------------------
// One trade, executed by a group of G consecutive lanes; lane `sub` of the
// group holds V consecutive words of each of the two rows (blocked layout, so
// lane order == ascending bit order).  All lanes of the group must call it
// together; groups of one warp may diverge from one another.
//
//   a' = (a&b) | picked ,  b' = (a&b) | rest     (or the other way round)
//
// where `picked` is a uniformly random k-subset, k = min(|a\b|, |b\a|), of the
// symmetric difference, drawn element by element ("r-th remaining element in
// ascending order", r exactly uniform) from the DEAL stream of (null, t).
// Returns true when the trade was applied (src/curveball.py:29 skip rule).
// ---------------------------------------------------------------------------
template <int G, int V>
__device__ __forceinline__ bool trade_rows(uint32_t *__restrict__ rows, int RS, uint32_t a, uint32_t b,
                                           uint32_t t, uint64_t null_id, const PhiloxKeys &keys,
                                           int sub, uint32_t gmask) {
    uint32_t A[V], B[V];
    uint32_t *ra = rows + (size_t)a * RS + sub * V;
    uint32_t *rb = rows + (size_t)b * RS + sub * V;
    if constexpr (V % 4 == 0) {
#pragma unroll
        for (int q = 0; q < V / 4; ++q) {
            uint4 va = make_uint4(0, 0, 0, 0), vb = va;
            if (sub * V + 4 * q < RS) {
                va = *reinterpret_cast<const uint4 *>(ra + 4 * q);
                vb = *reinterpret_cast<const uint4 *>(rb + 4 * q);
            }
            A[4 * q] = va.x; A[4 * q + 1] = va.y; A[4 * q + 2] = va.z; A[4 * q + 3] = va.w;
            B[4 * q] = vb.x; B[4 * q + 1] = vb.y; B[4 * q + 2] = vb.z; B[4 * q + 3] = vb.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) {
            bool in = sub * V + i < RS;
            A[i] = in ? ra[i] : 0u;
            B[i] = in ? rb[i] : 0u;
        }
    }

    uint32_t rem[V], pick[V];
    uint32_t ca = 0, cb = 0;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        uint32_t oa = A[i] & ~B[i], ob = B[i] & ~A[i];
        ca += __popc(oa); cb += __popc(ob);
        rem[i] = oa | ob; pick[i] = 0;
    }
    // group-wide inclusive scan of (ca, cb) packed in one register
    uint32_t packed = ca | (cb << 16);
    uint32_t inc = packed;
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        uint32_t v = __shfl_up_sync(gmask, inc, o, G);
        if (sub >= o) inc += v;
    }
    uint32_t tot = __shfl_sync(gmask, inc, G - 1, G);
    uint32_t La = tot & 0xffffu, Lb = tot >> 16;
    if (La == 0 || Lb == 0) return false;

    uint32_t exc = inc - packed;
    uint32_t pre = (exc & 0xffffu) + (exc >> 16);   // symmetric-difference elements in lower lanes
    uint32_t ctot = ca + cb;                        // ... in this lane
    const bool small_is_a = La <= Lb;
    const uint32_t k = small_is_a ? La : Lb;
    const uint32_t n = La + Lb;

    uint32_t di = 0;          // index of the next draw of the DEAL stream (group-uniform)
    uint4 blk = make_uint4(0, 0, 0, 0);
    for (uint32_t j = 0; j < k; ++j) {
        const uint32_t m = n - j;
        uint32_t r;
        for (;;) {
            if ((di & (4 * G - 1)) == 0)            // every lane refills its own block of this round
                blk = stream_block(keys, SDX_STREAM_DEAL, null_id, t, (di >> 2) + sub);
            uint32_t mine = pick4(blk, di & 3);
            uint32_t x = __shfl_sync(gmask, mine, (di >> 2) & (G - 1), G);
            ++di;
            uint64_t prod = (uint64_t)x * m;
            uint32_t lo = (uint32_t)prod;
            r = (uint32_t)(prod >> 32);
            if (lo >= m) break;                     // common case
            if (lo >= (0u - m) % m) break;          // Lemire rejection threshold
        }
        uint32_t d = r - pre;
        if (d < ctot) {                             // this lane owns the r-th remaining element
            bool done = false;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                uint32_t cw = __popc(rem[i]);
                bool here = !done && d < cw;
                if (here) {
                    uint32_t x = rem[i];
                    for (uint32_t q = 0; q < d; ++q) x &= x - 1;
                    uint32_t bit = x & (0u - x);
                    rem[i] ^= bit; pick[i] |= bit;
                }
                if (!done && !here) d -= cw;
                done |= here;
            }
            --ctot;
        } else if (r < pre) {
            --pre;
        }
    }

#pragma unroll
    for (int i = 0; i < V; ++i) {
        uint32_t ab = A[i] & B[i];
        A[i] = ab | (small_is_a ? pick[i] : rem[i]);
        B[i] = ab | (small_is_a ? rem[i] : pick[i]);
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
    return true;
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
    int cov, tot, act;
};

__device__ __forceinline__ SetScore score_module(const uint32_t *__restrict__ rows, int RS, bool rows_are_samples,
                                                 const int32_t *__restrict__ mod, int k,
                                                 const double *__restrict__ w, const uint32_t *__restrict__ mask,
                                                 int S, int lane) {
    double pos = 0.0, neg = 0.0;
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
            }
        }
    } else {
        for (int s = lane; s < S; s += 32) {
            if (!((mask[s >> 5] >> (s & 31)) & 1u)) continue;
            int mult = 0;
#pragma unroll
            for (int j = 0; j < SDX_MAX_K; ++j)
                if (g[j] >= 0) mult += (rows[(size_t)s * RS + (g[j] >> 5)] >> (g[j] & 31)) & 1u;
            if (mult) {
                const double ws = w[s];
                ++cov; tot += mult;
                if (ws > 0.0) { pos += ws; ++act; }
                neg += (double)mult * fabs(ws);
            }
        }
    }
    pos = warp_sum_f64(pos);
    neg = warp_sum_f64(neg);
    SetScore o;
    o.W = 2.0 * pos - neg;
    o.cov = warp_sum_i32(cov);
    o.tot = warp_sum_i32(tot);
    o.act = warp_sum_i32(act);
    return o;
}
