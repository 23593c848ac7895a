// sdx_lane_trade.cuh — Curveball trades with ONE THREAD PER NULL: the per-lane device code of k_trade_lanes (sdx_lanes.cu).
//
// Reference semantics: src/curveball.py:25-35 (one trade), src/superdendrix.py:387-391 (SuperW).  The canonical
// algorithm (ranks of the symmetric difference in ascending element order, Floyd picks for the smaller side, one DEAL
// sub-stream per pick: DESIGN.md §3, oracle/sdx_oracle.c: trade_canonical) is the one of sdx_trade.cuh, so the nulls are
// bit-identical to k_trade's and to the oracle's whatever the row representation.
//
// Layout ("group state"): 32 nulls that walk the same schedule of row pairs are stored interleaved.  A row with at
// most TS elements is a SORTED LIST of u16 element indices, entry i of lane l at  list[i * 32 + l]  (64 bytes per
// entry); a larger row stays BIT-PACKED, word w of lane l at  bits[w * 32 + l]  (128 bytes per word).  Row sizes are
// invariants of Curveball, so every row has a fixed slot and — because the 32 lanes trade the same two rows — every
// loop bound below is warp-uniform: no lock-step inflation, no shuffles, no level plan.
//
// Everything here is __host__ __device__ and touches memory only through the lane's own column pointers, so that
// tests/tools/lane_host.cu can run the very same code lane by lane on the CPU against the oracle (no GPU needed).
#pragma once
#include "sdx_common.cuh"

#define LN_W 32                  // lanes (nulls) per group
#define LN_SENT 0xFFFFu          // list sentinel: larger than every element (the lane kernel needs L <= 65535)
#define LN_TS_MAX 64             // longest list the kernel accepts (symmetric difference <= 128 ranks)

#ifdef __CUDA_ARCH__
#define LN_WARP_MAX(x) __reduce_max_sync(0xffffffffu, (x))
#define LN_LDCG16(p) __ldcg(p)
#define LN_LDCG32(p) __ldcg(p)
#define LN_POPC(x) __popc(x)
#define LN_FFS(x) __ffs(x)
#define LN_NOINLINE __noinline__
#else
#define LN_WARP_MAX(x) (x)
#define LN_LDCG16(p) (*(p))
#define LN_LDCG32(p) (*(p))
#define LN_POPC(x) __builtin_popcount(x)
#define LN_FFS(x) __builtin_ffs((int)(x))
#define LN_NOINLINE
#endif
#define LN_HD __host__ __device__ __forceinline__

// row descriptor: x = byte offset of the row inside the group state (multiple of 64), y = size | kind << 31 (1 = bit-packed)
LN_HD uint32_t ln_size(uint2 d) { return d.y & 0xffffu; }
LN_HD bool ln_is_bits(uint2 d) { return (d.y >> 31) != 0; }
LN_HD uint32_t ln_row_bytes(uint2 d, int nw) { return ln_is_bits(d) ? (uint32_t)nw * 128u : ln_size(d) * 64u; }

struct LnRng {                   // the DEAL stream of one trade of one null
    const PhiloxKeys *keys;
    uint64_t null_id;
    uint32_t t;
};

// Rare branch of Lemire's method (probability m / 2^32 per pick): same sub-stream rule as sdx_trade.cuh: deal_below_slow.
static __host__ __device__ LN_NOINLINE uint32_t ln_deal_slow(const PhiloxKeys *keys, uint64_t null_id, uint32_t t, uint32_t i,
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

LN_HD uint32_t ln_pick(const LnRng &g, uint32_t i, uint32_t m, uint32_t x) {      // uniform in [0, m) from the first draw x of pick i
    const uint64_t prod = (uint64_t)x * m;
    uint32_t r = (uint32_t)(prod >> 32);
    if ((uint32_t)prod < m) r = ln_deal_slow(g.keys, g.null_id, g.t, i, m, x);
    return r;
}

struct LnDeal {                  // what phase 1 of a trade hands to phase 2
    uint32_t cm;                 // chosen ranks (register form: n <= 32)
    uint32_t lb;                 // |b \ a|
    bool go;                     // the trade is applied (src/curveball.py:29: skipped when one row contains the other)
    bool small_is_a;
};

// ---------------------------------------------------------------------------------------------------------------
// list x list.  A, B: sorted lists of sa, sb entries followed by one LN_SENT entry.  Phase 1 counts the common
// elements (one merge walk) and draws the picks; phase 2 is the second merge walk that deals the elements into the
// output lists OA (sa entries) and OB (sb entries), both with room for one entry more.  The walks run sa + sb steps in
// every lane (the c steps that a lane with c common elements has left over find both lists at their sentinels).
// MREG: the chosen ranks fit one register (sa + sb <= 32); otherwise they live in C (u32 words at stride 32).
// ---------------------------------------------------------------------------------------------------------------
template <bool MREG>
LN_HD LnDeal ln_ll_phase1(const uint16_t *A, int sa, const uint16_t *B, int sb, uint32_t *C, const LnRng &g, uint32_t &kmax_out) {
    const int steps = sa + sb;
    const uint16_t *pa = A, *pb = B;
    uint32_t c = 0;
    for (int s = 0; s < steps; ++s) {
        const uint32_t x = *pa, y = *pb;
        const bool valid = (x & y) != LN_SENT;
        const bool le = x <= y, ge = y <= x;
        c += (le && ge && valid) ? 1u : 0u;
        pa += (le && valid) ? LN_W : 0;
        pb += (ge && valid) ? LN_W : 0;
    }
    const uint32_t La = (uint32_t)sa - c, Lb = (uint32_t)sb - c, n = La + Lb;
    LnDeal d;
    d.go = La != 0 && Lb != 0;
    d.small_is_a = La <= Lb;
    d.lb = Lb;
    d.cm = 0;
    const uint32_t k = d.go ? (d.small_is_a ? La : Lb) : 0u;
    const uint32_t kmax = LN_WARP_MAX(k);
    kmax_out = kmax;
    if (kmax == 0) return d;                            // warp-uniform: nobody trades
    if (!MREG) {
        const int nwc = (steps >> 5) + 2;
        for (int w = 0; w < nwc; ++w) C[w * LN_W] = 0u;
    }
    uint32_t cm = 0;
    for (uint32_t blk = 0; blk * 4 < kmax; ++blk) {
        const uint4 xb = stream_block(*g.keys, SDX_STREAM_DEAL, g.null_id, g.t, blk);
        const uint32_t x[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = blk * 4 + q;
            if (i < k) {
                const uint32_t m = n - k + i + 1;       // Floyd: j = m - 1, r uniform in [0, j]
                uint32_t r = ln_pick(g, i, m, x[q]);
                if (MREG) {
                    if ((cm >> r) & 1u) r = m - 1;
                    cm |= 1u << r;
                } else {
                    uint32_t w = C[(r >> 5) * LN_W];
                    if ((w >> (r & 31)) & 1u) { r = m - 1; w = C[(r >> 5) * LN_W]; }
                    C[(r >> 5) * LN_W] = w | (1u << (r & 31));
                }
            }
        }
    }
    d.cm = cm;
    return d;
}

template <bool MREG>
LN_HD void ln_ll_phase2(const uint16_t *A, int sa, const uint16_t *B, int sb, uint16_t *OA, uint16_t *OB, const uint32_t *C, const LnDeal &d) {
    const int steps = sa + sb;
    const uint16_t *pa = A, *pb = B;
    uint16_t *oa = OA, *ob = OB;
    // a lane that does not trade keeps its rows: every symmetric-difference element then comes from one row only, so
    // "picked for the smaller side a" = all (b \ a is empty) or none (a \ b is empty) reproduces the input
    const bool sia = d.go ? d.small_is_a : true;
    const uint32_t keep = d.lb == 0 ? 0xffffffffu : 0u;
    uint32_t cur = MREG ? (d.go ? d.cm : keep) : 0u, rank = 0;
    for (int s = 0; s < steps; ++s) {
        const uint32_t x = *pa, y = *pb;
        const bool valid = (x & y) != LN_SENT;
        const bool le = x <= y, ge = y <= x;
        const bool common = le && ge;
        const bool symel = valid && !common;
        if (!MREG) {
            if (symel && (rank & 31u) == 0) cur = d.go ? C[(rank >> 5) * LN_W] : keep;
        }
        const bool picked = (cur & 1u) != 0;
        if (symel) { cur >>= 1; ++rank; }
        const bool to_a = common || picked == sia;
        const bool to_b = common || picked != sia;
        const uint16_t e = (uint16_t)(le ? x : y);
        *oa = e; *ob = e;                               // stored to both lists; only the list that takes it advances
        oa += (valid && to_a) ? LN_W : 0;
        ob += (valid && to_b) ? LN_W : 0;
        pa += (le && valid) ? LN_W : 0;
        pb += (ge && valid) ? LN_W : 0;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bit-packed x bit-packed (nw words each, stride 32).  Phase 1: |a\b|, |b\a| and the picks as a rank mask in C
// (nw + 2 words).  Phase 2: every word takes its popc(a ^ b) bits of the rank mask and deposits them onto the set
// bits of a ^ b (software pdep); DA / DB may alias A / B (each word is read before it is written, by its own lane).
// ---------------------------------------------------------------------------------------------------------------
LN_HD uint32_t ln_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}

LN_HD LnDeal ln_bb_phase1(const uint32_t *A, const uint32_t *B, int nw, uint32_t *C, const LnRng &g, uint32_t &kmax_out) {
    uint32_t La = 0, Lb = 0;
    for (int w = 0; w < nw; ++w) {
        const uint32_t a = A[w * LN_W], b = B[w * LN_W];
        La += LN_POPC(a & ~b);
        Lb += LN_POPC(b & ~a);
    }
    const uint32_t n = La + Lb;
    LnDeal d;
    d.go = La != 0 && Lb != 0;
    d.small_is_a = La <= Lb;
    d.lb = Lb;
    d.cm = 0;
    const uint32_t k = d.go ? (d.small_is_a ? La : Lb) : 0u;
    const uint32_t kmax = LN_WARP_MAX(k);
    kmax_out = kmax;
    if (kmax == 0) return d;
    for (int w = 0; w < nw + 2; ++w) C[w * LN_W] = 0u;
    for (uint32_t blk = 0; blk * 4 < kmax; ++blk) {
        const uint4 xb = stream_block(*g.keys, SDX_STREAM_DEAL, g.null_id, g.t, blk);
        const uint32_t x[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = blk * 4 + q;
            if (i < k) {
                const uint32_t m = n - k + i + 1;
                uint32_t r = ln_pick(g, i, m, x[q]);
                uint32_t w = C[(r >> 5) * LN_W];
                if ((w >> (r & 31)) & 1u) { r = m - 1; w = C[(r >> 5) * LN_W]; }
                C[(r >> 5) * LN_W] = w | (1u << (r & 31));
            }
        }
    }
    return d;
}

LN_HD void ln_bb_phase2(const uint32_t *A, const uint32_t *B, uint32_t *DA, uint32_t *DB, int nw, const uint32_t *C, const LnDeal &d) {
    uint32_t rank = 0;
    for (int w = 0; w < nw; ++w) {
        const uint32_t a = A[w * LN_W], b = B[w * LN_W];
        uint32_t na = a, nb = b;
        const uint32_t sym = a ^ b, cnt = LN_POPC(sym);
        if (d.go && cnt) {
            const uint32_t w0 = rank >> 5;
            uint32_t ch = ln_funnel_r(C[w0 * LN_W], C[(w0 + 1) * LN_W], rank & 31u);
            if (cnt < 32) ch &= (1u << cnt) - 1u;
            uint32_t m = sym, pk = 0;
            while (ch) {
                const uint32_t low = m & (0u - m);
                m ^= low;
                if (ch & 1u) pk |= low;
                ch >>= 1;
            }
            const uint32_t ab = a & b, rest = sym ^ pk;
            na = ab | (d.small_is_a ? pk : rest);
            nb = ab | (d.small_is_a ? rest : pk);
        }
        DA[w * LN_W] = na; DB[w * LN_W] = nb;
        rank += cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// list <-> bit-packed scratch row (trades between a list row and a bit-packed row run as bit-packed trades)
// ---------------------------------------------------------------------------------------------------------------
LN_HD void ln_list_to_bits(const uint16_t *Lst, int size, uint32_t *X, int nw) {
    for (int w = 0; w < nw; ++w) X[w * LN_W] = 0u;
    for (int i = 0; i < size; ++i) {
        const uint32_t e = Lst[i * LN_W];
        X[(e >> 5) * LN_W] |= 1u << (e & 31);
    }
}

LN_HD void ln_bits_to_list(const uint32_t *X, int nw, uint16_t *Out) {
    uint16_t *o = Out;
    for (int w = 0; w < nw; ++w) {
        uint32_t x = X[w * LN_W];
        while (x) {
            *o = (uint16_t)(w * 32 + LN_FFS(x) - 1);
            o += LN_W;
            x &= x - 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One trade of one lane on staged rows.  inA / inB: the two rows as they sit in the group state (lists or words, lanes
// interleaved); outA / outB: the new rows in the same form; X: scratch row (nw words), C: chosen ranks — both at stride
// 32 and already offset to the lane.  Phase 1 only reads the rows (and writes the list sentinels), phase 2 writes the
// outputs; the kernel waits for the previous bulk store between the two.  A trade between a list row and a bit-packed
// row runs as a bit-packed trade on the scratch form of the list.
// ---------------------------------------------------------------------------------------------------------------
template <bool MREG>
LN_HD LnDeal ln_trade_phase1(uint2 dA, uint2 dB, uint8_t *inA, uint8_t *inB, uint32_t *X, uint32_t *C, int nw, int lane, const LnRng &g,
                             uint32_t &kmax) {
    const int sa = (int)ln_size(dA), sb = (int)ln_size(dB);
    const bool ba = ln_is_bits(dA), bb = ln_is_bits(dB);
    kmax = 0;
    LnDeal deal;
    deal.cm = 0; deal.lb = 0; deal.go = false; deal.small_is_a = true;
    if (sa == 0 || sb == 0) return deal;                  // an empty row never trades (warp-uniform)
    if (!ba && !bb) {
        uint16_t *A = reinterpret_cast<uint16_t *>(inA) + lane, *B = reinterpret_cast<uint16_t *>(inB) + lane;
        A[sa * LN_W] = (uint16_t)LN_SENT;
        B[sb * LN_W] = (uint16_t)LN_SENT;
        return ln_ll_phase1<MREG>(A, sa, B, sb, C, g, kmax);
    }
    if (!ba) ln_list_to_bits(reinterpret_cast<const uint16_t *>(inA) + lane, sa, X, nw);
    if (!bb) ln_list_to_bits(reinterpret_cast<const uint16_t *>(inB) + lane, sb, X, nw);
    const uint32_t *srcA = ba ? reinterpret_cast<const uint32_t *>(inA) + lane : X;
    const uint32_t *srcB = bb ? reinterpret_cast<const uint32_t *>(inB) + lane : X;
    return ln_bb_phase1(srcA, srcB, nw, C, g, kmax);
}

template <bool MREG>
LN_HD void ln_trade_phase2(uint2 dA, uint2 dB, const uint8_t *inA, const uint8_t *inB, uint8_t *outA, uint8_t *outB, uint32_t *X,
                           const uint32_t *C, int nw, int lane, const LnDeal &deal) {
    const int sa = (int)ln_size(dA), sb = (int)ln_size(dB);
    const bool ba = ln_is_bits(dA), bb = ln_is_bits(dB);
    if (!ba && !bb) {
        ln_ll_phase2<MREG>(reinterpret_cast<const uint16_t *>(inA) + lane, sa, reinterpret_cast<const uint16_t *>(inB) + lane, sb,
                           reinterpret_cast<uint16_t *>(outA) + lane, reinterpret_cast<uint16_t *>(outB) + lane, C, deal);
        return;
    }
    const uint32_t *srcA = ba ? reinterpret_cast<const uint32_t *>(inA) + lane : X;
    const uint32_t *srcB = bb ? reinterpret_cast<const uint32_t *>(inB) + lane : X;
    uint32_t *dstA = ba ? reinterpret_cast<uint32_t *>(outA) + lane : X;
    uint32_t *dstB = bb ? reinterpret_cast<uint32_t *>(outB) + lane : X;
    ln_bb_phase2(srcA, srcB, dstA, dstB, nw, C, deal);
    if (!ba) ln_bits_to_list(X, nw, reinterpret_cast<uint16_t *>(outA) + lane);
    if (!bb) ln_bits_to_list(X, nw, reinterpret_cast<uint16_t *>(outB) + lane);
}

// entry of the trade schedule: rows a, b (14 bits each) and, in bits 28..30, whether the trade shares a row with the
// first / second / third trade before it (those cannot be prefetched blindly)
LN_HD uint32_t ln_entry_rows_share(uint32_t e, uint32_t q) {
    const uint32_t a = e & 0x3fffu, b = (e >> 14) & 0x3fffu, qa = q & 0x3fffu, qb = (q >> 14) & 0x3fffu;
    return (qa == a || qa == b || qb == a || qb == b) ? 1u : 0u;
}

// ---------------------------------------------------------------------------------------------------------------
// SuperW of one module on the lane's null (rows = features), read from the group state in global memory.
//   W = 2 * sum_{s in U, w_s > 0} w_s - sum_g sum_{s in g} |w_s|            (src/superdendrix.py:387-391)
// The fp64 sums are formed exactly as score_module (sdx_trade.cuh) forms them — partial p = the samples of mask words
// p, p + 32, .. in ascending order, then the butterfly tree over the 32 partials — so a null scores bit-identically on
// both kernels and a null equal to the observed matrix scores exactly z_obs (ties count, src/superdendrix.py:192).
// PP / PN: 32 fp64 partials each at stride 32 (scratch of this lane).
// ---------------------------------------------------------------------------------------------------------------
LN_HD double ln_score(const uint8_t *state, int lane, const uint2 *desc, int nwL, const int32_t *mod, int k,
                      const double *w, const uint32_t *mask, int S, double *PP, double *PN) {
    const int nw = (S + 31) >> 5;
    for (int i = 0; i < 32; ++i) { PP[i * LN_W] = 0.0; PN[i * LN_W] = 0.0; }
    // per module slot: a cursor into a sorted list, or the bit-packed row
    const uint16_t *lp[SDX_MAX_K];
    const uint32_t *bp[SDX_MAX_K];
    int left[SDX_MAX_K];
    uint32_t head[SDX_MAX_K];
#pragma unroll
    for (int j = 0; j < SDX_MAX_K; ++j) {
        lp[j] = nullptr; bp[j] = nullptr; left[j] = 0; head[j] = 0xffffffffu;
        const int g = j < k ? mod[j] : -1;
        if (g >= 0) {
            const uint2 d = desc[g];
            if (ln_is_bits(d)) bp[j] = reinterpret_cast<const uint32_t *>(state + d.x) + lane;
            else {
                lp[j] = reinterpret_cast<const uint16_t *>(state + d.x) + lane;
                left[j] = (int)ln_size(d);
                if (left[j] > 0) head[j] = LN_LDCG16(lp[j]);
            }
        }
    }
    (void)nwL;
    for (int wi = 0; wi < nw; ++wi) {
        const uint32_t m = mask[wi];
        uint32_t r[SDX_MAX_K];
        uint32_t U = 0;
#pragma unroll
        for (int j = 0; j < SDX_MAX_K; ++j) {
            uint32_t x = 0;
            if (bp[j]) x = LN_LDCG32(bp[j] + (size_t)wi * LN_W);
            else {
                while (left[j] > 0 && (head[j] >> 5) == (uint32_t)wi) {
                    x |= 1u << (head[j] & 31);
                    lp[j] += LN_W; --left[j];
                    head[j] = left[j] > 0 ? (uint32_t)LN_LDCG16(lp[j]) : 0xffffffffu;
                }
            }
            r[j] = x & m;
            U |= r[j];
        }
        double pos = PP[(wi & 31) * LN_W], neg = PN[(wi & 31) * LN_W];
        uint32_t x = U;
        while (x) {
            const int b = LN_FFS(x) - 1;
            x &= x - 1;
            const double ws = w[wi * 32 + b];
            int mult = 0;
#pragma unroll
            for (int j = 0; j < SDX_MAX_K; ++j) mult += (r[j] >> b) & 1;
            if (ws > 0.0) pos += ws;
            neg += (double)mult * fabs(ws);
        }
        PP[(wi & 31) * LN_W] = pos; PN[(wi & 31) * LN_W] = neg;
    }
    for (int o = 16; o > 0; o >>= 1)
        for (int i = 0; i < o; ++i) {
            PP[i * LN_W] += PP[(i + o) * LN_W];
            PN[i * LN_W] += PN[(i + o) * LN_W];
        }
    return 2.0 * PP[0] - PN[0];
}

// the lane's null as R x wprL bit-packed trade rows (dst is zero-filled by the caller)
LN_HD void ln_emit_rows(const uint8_t *state, int lane, const uint2 *desc, int R, int wprL, uint32_t *dst) {
    for (int r = 0; r < R; ++r) {
        const uint2 d = desc[r];
        uint32_t *o = dst + (size_t)r * wprL;
        if (ln_is_bits(d)) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(state + d.x) + lane;
            for (int w = 0; w < wprL; ++w) o[w] = LN_LDCG32(q + (size_t)w * LN_W);
        } else {
            const uint16_t *p = reinterpret_cast<const uint16_t *>(state + d.x) + lane;
            const int n = (int)ln_size(d);
            uint32_t acc = 0, cw = 0;
            for (int i = 0; i < n; ++i) {
                const uint32_t e = LN_LDCG16(p + (size_t)i * LN_W);
                if ((e >> 5) != cw) { if (acc) o[cw] = acc; acc = 0; cw = e >> 5; }
                acc |= 1u << (e & 31);
            }
            if (acc) o[cw] = acc;
        }
    }
}

// start state of the lane's chain: a compact state (the group layout with one lane: list entries / words contiguous)
LN_HD void ln_load_state(uint8_t *state, int lane, const uint2 *desc, int R, int wprL, const uint8_t *compact) {
    for (int r = 0; r < R; ++r) {
        const uint2 d = desc[r];
        if (ln_is_bits(d)) {
            uint32_t *q = reinterpret_cast<uint32_t *>(state + d.x) + lane;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(compact + d.x / LN_W);
            for (int w = 0; w < wprL; ++w) q[(size_t)w * LN_W] = src[w];
        } else {
            uint16_t *p = reinterpret_cast<uint16_t *>(state + d.x) + lane;
            const uint16_t *src = reinterpret_cast<const uint16_t *>(compact + d.x / LN_W);
            const int n = (int)ln_size(d);
            for (int i = 0; i < n; ++i) p[(size_t)i * LN_W] = src[i];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host: the lane layout of a matrix — which rows are lists, where every row sits in a group state
// ---------------------------------------------------------------------------------------------------------------
struct LnLayout {
    int ts = 0, slot_bytes = 0, n_bits = 0, max_list = 0;
    uint32_t gsb = 0;
};

inline int ln_default_ts(int nw) {
    // a list trade costs ~28 instructions per element of the two rows, a bit-packed trade ~60 per word: short rows are lists
    int ts = nw / 2 + 4;
    if (ts < 16) ts = 16;
    if (ts > LN_TS_MAX) ts = LN_TS_MAX;
    return ts;
}

inline LnLayout ln_build_layout(const int *size, int R, int nw, int ts, uint2 *desc) {
    LnLayout lay;
    lay.ts = ts;
    uint32_t off = 0;
    for (int r = 0; r < R; ++r)
        if (size[r] > ts) { desc[r] = make_uint2(off, (uint32_t)(size[r] > 0xffff ? 0xffff : size[r]) | 0x80000000u); off += (uint32_t)nw * 128u; ++lay.n_bits; }
    for (int r = 0; r < R; ++r)
        if (size[r] <= ts) { desc[r] = make_uint2(off, (uint32_t)size[r]); off += (uint32_t)size[r] * 64u; if (size[r] > lay.max_list) lay.max_list = size[r]; }
    lay.gsb = ((off + 127u) / 128u) * 128u + 128u;
    int slot = (lay.max_list + 1) * 64;
    if (lay.n_bits && nw * 128 > slot) slot = nw * 128;
    lay.slot_bytes = ((slot + 127) / 128) * 128;
    return lay;
}
