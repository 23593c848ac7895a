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
// Three kinds of trade:
//   list x list        both rows staged in shared memory (TMA), two merge walks, picked ranks in a register mask
//   bit-packed x list  both staged (the bit-packed row in the warp's one large buffer X); nothing is converted: the
//                      list's entries are tested against the bits, a picked rank is mapped to its element by a select on
//                      the bits, and only the <= 2k elements that change rows are toggled
//   bit-packed x bit-packed  both rows read and updated IN PLACE in the group state (every lane touches only its own
//                      column), rank mask in the group's scratch words (rare: the few dense rows)
//
// Everything here is __host__ __device__ and touches memory only through the lane's own column pointers, so that
// tests/tools/lane_host.cu can run the very same code lane by lane on the CPU against the oracle (no GPU needed).
#pragma once
#include "sdx_common.cuh"

#define LN_W 32                  // lanes (nulls) per group
#define LN_SENT 0xFFFFu          // list sentinel: larger than every element (the lane kernel needs L <= 65535)
#define LN_TS_MAX 32             // longest list the kernel accepts (symmetric difference of two lists <= 64 ranks: a register pair)

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
LN_HD uint32_t ln_stage_bytes(uint2 d) { return ln_is_bits(d) ? 0u : ln_size(d) * 64u; }      // the list buffers take only lists

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

// chosen ranks of a list x list trade: one register when 2 TS <= 32, a pair otherwise
template <bool M64> struct LnMask;
template <> struct LnMask<false> {
    uint32_t v;
    LN_HD LnMask() : v(0) {}
    LN_HD explicit LnMask(uint32_t x) : v(x) {}
    LN_HD bool test(uint32_t r) const { return (v >> r) & 1u; }
    LN_HD void set(uint32_t r) { v |= 1u << r; }
    LN_HD bool low() const { return v & 1u; }
    LN_HD void shr() { v >>= 1; }
    LN_HD LnMask inv() const { return LnMask(~v); }
    static LN_HD LnMask fill(bool ones) { return LnMask(ones ? 0xffffffffu : 0u); }
};
template <> struct LnMask<true> {
    uint64_t v;
    LN_HD LnMask() : v(0) {}
    LN_HD explicit LnMask(uint64_t x) : v(x) {}
    LN_HD bool test(uint32_t r) const { return (v >> r) & 1ull; }
    LN_HD void set(uint32_t r) { v |= 1ull << r; }
    LN_HD bool low() const { return v & 1ull; }
    LN_HD void shr() { v >>= 1; }
    LN_HD LnMask inv() const { return LnMask(~v); }
    static LN_HD LnMask fill(bool ones) { return LnMask(ones ? ~0ull : 0ull); }
};

struct LnDeal {                  // what phase 1 of a trade hands to phase 2
    uint32_t k;                  // picks of this lane (0: the lane keeps its rows)
    uint32_t cmn;                // bit-packed x list: which list entries are common
    bool go;                     // the trade is applied (src/curveball.py:29: skipped when one row contains the other)
};

// ---------------------------------------------------------------------------------------------------------------
// list x list.  A, B: sorted lists of sa, sb entries followed by one LN_SENT entry.  Phase 1 counts the common
// elements (one merge walk) and draws the picks; phase 2 is the second merge walk that deals the elements into the
// output lists OA (sa entries) and OB (sb entries), both with room for one entry more.  The walks run sa + sb steps in
// every lane (the c steps that a lane with c common elements has left over find both lists at their sentinels).
// toA: bit r = the symmetric-difference element of rank r goes to row a.
// ---------------------------------------------------------------------------------------------------------------
// common elements of two sorted lists (one merge walk)
LN_HD uint32_t ln_ll_common(const uint16_t *A, int sa, const uint16_t *B, int sb) {
    const int steps = sa + sb;
    uint32_t c = 0;
#ifdef __CUDA_ARCH__
    // the compiler turns the predicates of the C loop below into register arithmetic (~18 instructions per step); spelled
    // out it is 10: two loads, four compares, one predicate AND, three predicated adds
    uint32_t pa = (uint32_t)__cvta_generic_to_shared(A), pb = (uint32_t)__cvta_generic_to_shared(B);
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
        asm volatile("{\n\t"
                     ".reg .pred px, py, pA, pB, pC;\n\t"
                     ".reg .b32 x, y;\n\t"
                     "ld.shared.u16 x, [%1];\n\t"
                     "ld.shared.u16 y, [%2];\n\t"
                     "setp.ne.u32 px, x, 0xFFFF;\n\t"
                     "setp.ne.u32 py, y, 0xFFFF;\n\t"
                     "setp.le.and.u32 pA, x, y, px;\n\t"
                     "setp.le.and.u32 pB, y, x, py;\n\t"
                     "and.pred pC, pA, pB;\n\t"
                     "@pC add.u32 %0, %0, 1;\n\t"
                     "@pA add.u32 %1, %1, 64;\n\t"
                     "@pB add.u32 %2, %2, 64;\n\t"
                     "}"
                     : "+r"(c), "+r"(pa), "+r"(pb) :: "memory");
    }
#else
    const uint16_t *pa = A, *pb = B;
    for (int s = 0; s < steps; ++s) {
        const uint32_t x = *pa, y = *pb;
        const bool adva = x <= y && x != LN_SENT, advb = y <= x && y != LN_SENT;
        if (adva && advb) ++c;
        if (adva) pa += LN_W;
        if (advb) pb += LN_W;
    }
#endif
    return c;
}

template <bool M64>
LN_HD void ln_ll_phase2(const uint16_t *A, int sa, const uint16_t *B, int sb, uint16_t *OA, uint16_t *OB, LnMask<M64> toA) {
    const int steps = sa + sb;
    const uint16_t *pa = A, *pb = B;
    uint16_t *oa = OA, *ob = OB;
    for (int s = 0; s < steps; ++s) {
        const uint32_t x = *pa, y = *pb;
        const bool adva = x <= y && x != LN_SENT, advb = y <= x && y != LN_SENT;      // neither: both lists are at their sentinels
        const bool bit = toA.low();
        const uint16_t e = (uint16_t)(x <= y ? x : y);
        *oa = e; *ob = e;                               // stored to both lists; only the list that takes it advances
        if (adva != advb) toA.shr();                    // a symmetric-difference element: dealt by its rank
        if (adva ? (advb || bit) : (advb && bit)) oa += LN_W;          // common, or dealt to a
        if (advb ? (adva || !bit) : (adva && !bit)) ob += LN_W;
        if (adva) pa += LN_W;
        if (advb) pb += LN_W;
    }
}

#ifdef __CUDA_ARCH__
// the same walk with the predicates spelled out (the compiler's version of the loop above runs ~38 instructions per step)
template <>
__device__ __forceinline__ void ln_ll_phase2<false>(const uint16_t *A, int sa, const uint16_t *B, int sb, uint16_t *OA, uint16_t *OB,
                                                    LnMask<false> toA) {
    const int steps = sa + sb;
    uint32_t pa = (uint32_t)__cvta_generic_to_shared(A), pb = (uint32_t)__cvta_generic_to_shared(B);
    uint32_t oa = (uint32_t)__cvta_generic_to_shared(OA), ob = (uint32_t)__cvta_generic_to_shared(OB);
    uint32_t m = toA.v;
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
        asm volatile("{\n\t"
                     ".reg .pred px, py, pA, pB, pS, pC, pbit, pTA, pTB, t1, t2;\n\t"
                     ".reg .b32 x, y, e, bit;\n\t"
                     "ld.shared.u16 x, [%0];\n\t"
                     "ld.shared.u16 y, [%1];\n\t"
                     "setp.ne.u32 px, x, 0xFFFF;\n\t"
                     "setp.ne.u32 py, y, 0xFFFF;\n\t"
                     "setp.le.and.u32 pA, x, y, px;\n\t"
                     "setp.le.and.u32 pB, y, x, py;\n\t"
                     "min.u32 e, x, y;\n\t"
                     "st.shared.u16 [%2], e;\n\t"
                     "st.shared.u16 [%3], e;\n\t"
                     "and.b32 bit, %4, 1;\n\t"
                     "setp.ne.u32 pbit, bit, 0;\n\t"
                     "xor.pred pS, pA, pB;\n\t"
                     "and.pred pC, pA, pB;\n\t"
                     "and.pred t1, pS, pbit;\n\t"
                     "and.pred t2, pS, !pbit;\n\t"
                     "or.pred pTA, pC, t1;\n\t"
                     "or.pred pTB, pC, t2;\n\t"
                     "@pS shr.u32 %4, %4, 1;\n\t"
                     "@pTA add.u32 %2, %2, 64;\n\t"
                     "@pTB add.u32 %3, %3, 64;\n\t"
                     "@pA add.u32 %0, %0, 64;\n\t"
                     "@pB add.u32 %1, %1, 64;\n\t"
                     "}"
                     : "+r"(pa), "+r"(pb), "+r"(oa), "+r"(ob), "+r"(m) :: "memory");
    }
}
#endif

// ---------------------------------------------------------------------------------------------------------------
// bit-packed x bit-packed (nw words each, stride 32, IN PLACE).  Phase 1: |a\b|, |b\a| and the picks as a rank mask in
// C (nw + 2 words).  Phase 2: every word takes its popc(a ^ b) bits of the rank mask and deposits them onto the set
// bits of a ^ b (software pdep).
// ---------------------------------------------------------------------------------------------------------------
LN_HD uint32_t ln_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#ifdef __CUDA_ARCH__
    return __funnelshift_r(lo, hi, s);
#else
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
#endif
}

LN_HD void ln_bb_count(const uint32_t *A, const uint32_t *B, int nw, uint32_t &La, uint32_t &Lb) {
    La = 0; Lb = 0;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) {
        const uint32_t a = LN_LDCG32(A + w * LN_W), b = LN_LDCG32(B + w * LN_W);      // L2: the rows are also rewritten by bulk stores
        La += LN_POPC(a & ~b);
        Lb += LN_POPC(b & ~a);
    }
}

LN_HD void ln_bb_phase2(uint32_t *A, uint32_t *B, int nw, const uint32_t *C, const LnDeal &d, bool small_is_a) {
    if (!d.go) return;                                  // in place: a lane that does not trade has nothing to write
    uint32_t rank = 0;
#pragma unroll 1
    for (int w = 0; w < nw; ++w) {
        const uint32_t a = LN_LDCG32(A + w * LN_W), b = LN_LDCG32(B + w * LN_W);
        const uint32_t sym = a ^ b, cnt = LN_POPC(sym);
        if (cnt) {
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
            A[w * LN_W] = ab | (small_is_a ? pk : rest);
            B[w * LN_W] = ab | (small_is_a ? rest : pk);
        }
        rank += cnt;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// bit-packed row (D: nw words STAGED in shared memory, sd elements) x list row (Lst: sl <= TS sorted entries, staged).
// Row kinds follow the row size (list iff size <= TS), so the list is the smaller side:
// |l\d| <= sl <= TS < sd - |d∩l| = |d\l|.  The list keeps its common elements and receives k = |l\d| of the n
// symmetric-difference elements, chosen by rank (Floyd, as everywhere).  Nothing is converted: the list's entries are
// tested against the bits, a picked rank is mapped to its element by a select on the bits, and only the <= 2k elements
// that change rows are toggled in D.
//   phase 1 (reads only): PF4[q] = elements of d below word 4q ((nw + 3) / 4 + 1 u16), per list entry RK[i] = its key
//                         (see ln_bl_prepare), the picked ranks PK[0 .. k) in ASCENDING order (insertion as they are drawn)
//   phase 2: resolves every pick against the ORIGINAL row d (a pick that is one of the list's own entries keeps it
//            there; any other is the y-th element of d, found by a search in PF4 and a select in the word), then
//            toggles the <= 2k bits of D and merges the entries that stay with the elements that arrive (both ascending)
//            into the new sorted list OL.
// All loop bounds (nw, sl, kmax) are warp-uniform.
// ---------------------------------------------------------------------------------------------------------------
LN_HD uint32_t ln_bl_prepare(const uint32_t *D, int nw, const uint16_t *Lst, int sl, uint16_t *PF4, uint16_t *RK, uint32_t &cmn_out) {
    const int nq = (nw + 3) >> 2;
    uint32_t acc = 0;
#pragma unroll 1
    for (int q = 0; q < nq; ++q) {
        PF4[q * LN_W] = (uint16_t)acc;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int w = 4 * q + v;
            if (w < nw) acc += LN_POPC(D[w * LN_W]);
        }
    }
    PF4[nq * LN_W] = (uint16_t)acc;
    uint32_t c = 0, cmn = 0;
#pragma unroll 1
    for (int i = 0; i < sl; ++i) {
        const uint32_t e = Lst[i * LN_W], w = e >> 5, sh = e & 31u;
        const uint32_t dw = D[w * LN_W];
        const uint32_t bit = (dw >> sh) & 1u;
        uint32_t below = (uint32_t)PF4[(w >> 2) * LN_W] + LN_POPC(dw & ((1u << sh) - 1u));     // elements of d below e
#pragma unroll
        for (int v = 0; v < 3; ++v) {
            const uint32_t ww = (w & ~3u) + (uint32_t)v;
            if (ww < w) below += LN_POPC(D[ww * LN_W]);
        }
        // key = elements of d\l below e + the list's own (not common) entries below e; the c common entries so far are all
        // below e.  For an own entry the key is its rank in the symmetric difference; a d\l element of rank r lies above a
        // common entry iff key <= r, above an own entry iff key < r.  Keys never decrease along the list.
        RK[i * LN_W] = (uint16_t)(below + (uint32_t)i - 2u * c);
        cmn |= bit << i;
        c += bit;
    }
    cmn_out = cmn;
    return c;
}

// position (0..31) of the j-th (0-based) set bit of x; j < popc(x)
LN_HD uint32_t ln_select32(uint32_t x, uint32_t j) {
    uint32_t pos = 0;
#pragma unroll
    for (int sh = 16; sh >= 1; sh >>= 1) {
        const uint32_t lowc = LN_POPC(x & ((1u << sh) - 1u));
        const bool up = j >= lowc;
        j -= up ? lowc : 0u;
        x = up ? (x >> sh) : x;
        pos += up ? (uint32_t)sh : 0u;
    }
    return pos;
}

LN_HD void ln_bl_phase2(uint32_t *D, int nw, const uint16_t *Lst, int sl, uint16_t *OL, const uint16_t *PF4, const uint16_t *RK, uint16_t *PK,
                        const LnDeal &d, uint32_t kmax) {
    const uint32_t k = d.k, cmn = d.cmn;
    const int nq = (nw + 3) >> 2;
    uint32_t keep = 0;
    int top = 1;
    while (top < nq) top <<= 1;
    // resolve the picks (the row d is still the original one)
#pragma unroll 1
    for (uint32_t j = 0; j < kmax; ++j) {
        uint32_t el = LN_SENT;
        if (j < k) {
            const uint32_t r = PK[j * LN_W];
            uint32_t nlt = 0, nle = 0;                  // entries with key < r, key <= r: prefixes of the list
#pragma unroll 2
            for (int i = 0; i < sl; ++i) {
                const uint32_t v = RK[i * LN_W];
                if (v < r) ++nlt;
                if (v <= r) ++nle;
            }
            const uint32_t mlt = nlt >= 32 ? 0xffffffffu : (1u << nlt) - 1u, mle = nle >= 32 ? 0xffffffffu : (1u << nle) - 1u;
            const uint32_t own = ~cmn & (mle ^ mlt);    // the own entry of rank r, if there is one
            keep |= own;
            if (!own) {
                // the y-th element of d: rank r minus the own entries below it plus the common entries below it
                const uint32_t y = r - (uint32_t)LN_POPC(~cmn & mlt) + (uint32_t)LN_POPC(cmn & mle);
                int lo = 0;                             // largest block of 4 words with PF4[lo] <= y
#pragma unroll 1
                for (int step = top >> 1; step >= 1; step >>= 1) {
                    const int cand = lo + step;
                    if (cand < nq && (uint32_t)PF4[cand * LN_W] <= y) lo = cand;
                }
                uint32_t base = PF4[lo * LN_W], w = 4u * (uint32_t)lo;
                bool found = false;
#pragma unroll
                for (int v = 0; v < 3; ++v) {           // the word inside the block (the last one needs no test)
                    const uint32_t cnt = LN_POPC(D[w * LN_W]);
                    if (!found && y >= base + cnt) { base += cnt; ++w; } else found = true;
                }
                el = w * 32u + ln_select32(D[w * LN_W], y - base);
            }
        }
        PK[j * LN_W] = (uint16_t)el;                     // the element that moves from d to the list (ascending in j), or LN_SENT
    }
    // the list's own entries that were not picked move to d; the picked elements of d\l leave it
    const uint32_t stay = cmn | keep;
#pragma unroll 1
    for (int i = 0; i < sl; ++i)
        if (!((stay >> i) & 1u)) {
            const uint32_t e = Lst[i * LN_W];
            D[(e >> 5) * LN_W] |= 1u << (e & 31u);
        }
#pragma unroll 1
    for (uint32_t j = 0; j < kmax; ++j) {
        const uint32_t el = PK[j * LN_W];
        if (el != LN_SENT) D[(el >> 5) * LN_W] &= ~(1u << (el & 31u));
    }
    // the new list: the entries that stay and the elements that arrive, both ascending, merged.  Every step consumes
    // one list entry or one pick (sl + kmax steps); an entry that left and a pick that was an own entry emit nothing.
    // (x == e never happens for two elements: an arriving element was not in the list.)
    const uint16_t *pl = Lst, *pp = PK;
    uint16_t *ol = OL;
    uint32_t st = stay, il = 0, ip = 0;
    const uint32_t steps = (uint32_t)sl + kmax;
#pragma unroll 1
    for (uint32_t s = 0; s < steps; ++s) {
        const uint32_t e = il < (uint32_t)sl ? (uint32_t)*pl : 0x10000u;         // past the end: larger than every element
        const uint32_t x = ip < kmax ? (uint32_t)*pp : 0x10000u;
        const bool take_l = e < x && x != LN_SENT;                               // a pick that was an own entry (LN_SENT) is skipped at once
        const bool emit = take_l ? (st & 1u) != 0 : (x < LN_SENT);
        *ol = (uint16_t)(take_l ? e : x);
        if (emit) ol += LN_W;
        if (take_l) { pl += LN_W; ++il; st >>= 1; } else { pp += LN_W; ++ip; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// One trade of one lane.  Staged lists: inA / inB as they arrive from the group state, outA / outB the new lists.  The
// bit-packed row of a bit-packed x list trade is staged in X and updated there; the two rows of a bit-packed x
// bit-packed trade are read and written in place in `state` (the lane's column of the row in the group state).
// Scratch, all at stride 32 and already offset to the lane: PF4 ((nw + 3) / 4 + 1 u16); RK / PK (TS u16 each) live in
// the list buffers the bit-packed row does not use (its input buffer, its output buffer); C = nw + 2 u32 words of the
// group state at c_off.  Pointers that only the rare kinds need are formed inside their branches (registers).
// Phase 1 only reads the rows (and writes the list sentinels), phase 2 writes; the kernel waits for the previous bulk
// store between the two.
// ---------------------------------------------------------------------------------------------------------------
template <bool M64>
struct LnTrade {
    LnDeal deal;
    LnMask<M64> toA;
    bool small_is_a;
};

template <bool M64>
LN_HD void ln_trade_phase1(LnTrade<M64> &tr, uint2 dA, uint2 dB, uint16_t *inA, uint16_t *inB, uint16_t *outA, uint16_t *outB, uint8_t *state,
                           uint32_t c_off, int lane, const uint32_t *X, uint16_t *PF, int nw, const LnRng &g, uint32_t &kmax) {
    const int sa = (int)ln_size(dA), sb = (int)ln_size(dB);
    const bool ba = ln_is_bits(dA), bb = ln_is_bits(dB);
    const int kind = ba == bb ? (ba ? 2 : 0) : 1;        // 0: list x list, 1: bit-packed x list, 2: bit-packed x bit-packed
    kmax = 0;
    tr.deal.k = 0; tr.deal.cmn = 0; tr.deal.go = false; tr.small_is_a = true;
    if (sa == 0 || sb == 0) return;                     // an empty row never trades (warp-uniform)
    // |a \ b|, |b \ a| by the kind of the rows
    uint32_t La, Lb;
    if (kind == 0) {
        inA[sa * LN_W] = (uint16_t)LN_SENT;
        inB[sb * LN_W] = (uint16_t)LN_SENT;
        const uint32_t c = ln_ll_common(inA, sa, inB, sb);
        La = (uint32_t)sa - c; Lb = (uint32_t)sb - c;
    } else if (kind == 2) {
        ln_bb_count(reinterpret_cast<const uint32_t *>(state + dA.x) + lane, reinterpret_cast<const uint32_t *>(state + dB.x) + lane, nw, La, Lb);
    } else {                                            // RK: the list buffer the bit-packed row does not use
        const uint32_t c = ln_bl_prepare(X, nw, ba ? inB : inA, ba ? sb : sa, PF, ba ? inA : inB, tr.deal.cmn);
        La = (uint32_t)sa - c; Lb = (uint32_t)sb - c;
    }
    const uint32_t n = La + Lb;
    tr.deal.go = La != 0 && Lb != 0;
    tr.small_is_a = La <= Lb;
    const uint32_t k = tr.deal.go ? (tr.small_is_a ? La : Lb) : 0u;
    tr.deal.k = k;
    // a lane that does not trade keeps its rows: every symmetric-difference element then comes from one row only
    tr.toA = LnMask<M64>::fill(Lb == 0);
    kmax = LN_WARP_MAX(k);
    if (kmax == 0) return;                              // warp-uniform: nobody trades
    uint32_t *C = reinterpret_cast<uint32_t *>(state + c_off) + lane;      // rank mask of a bit-packed x bit-packed trade
    uint16_t *PK = ba ? outA : outB;                    // picks of a bit-packed x list trade: the output buffer of the bit-packed row
    if (kind == 2) {
#pragma unroll 1
        for (int w = 0; w < nw + 2; ++w) C[w * LN_W] = 0u;
    }
    // Floyd's picks: k ranks out of n (the one place the DEAL stream is drawn)
    LnMask<M64> cm;
#pragma unroll 1
    for (uint32_t blk = 0; blk * 4 < kmax; ++blk) {
        const uint4 xb = stream_block(*g.keys, SDX_STREAM_DEAL, g.null_id, g.t, blk);
        const uint32_t x[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t i = blk * 4 + q;
            if (i < k) {
                const uint32_t m = n - k + i + 1;       // j = m - 1, r uniform in [0, j]; every rank chosen so far is < j
                uint32_t r = ln_pick(g, i, m, x[q]);
                if (kind == 0) {
                    if (cm.test(r)) r = m - 1;
                    cm.set(r);
                } else if (kind == 1) {
                    // PK[0 .. i) ascending: position of r, or the end for the replacement m - 1 of a rank drawn before
                    uint32_t pos = 0;
                    bool dup = false;
#pragma unroll 1
                    for (uint32_t j = 0; j < i; ++j) {
                        const uint32_t v = PK[j * LN_W];
                        if (v < r) ++pos;
                        dup = dup || v == r;
                    }
                    if (dup) { r = m - 1; pos = i; }
#pragma unroll 1
                    for (uint32_t j = i; j > pos; --j) PK[j * LN_W] = PK[(j - 1) * LN_W];
                    PK[pos * LN_W] = (uint16_t)r;
                } else {
                    uint32_t w = C[(r >> 5) * LN_W];
                    if ((w >> (r & 31)) & 1u) { r = m - 1; w = C[(r >> 5) * LN_W]; }
                    C[(r >> 5) * LN_W] = w | (1u << (r & 31));
                }
            }
        }
    }
    if (kind == 0 && tr.deal.go) tr.toA = tr.small_is_a ? cm : cm.inv();
}

template <bool M64>
LN_HD void ln_trade_phase2(const LnTrade<M64> &tr, uint2 dA, uint2 dB, uint16_t *inA, uint16_t *inB, uint16_t *outA, uint16_t *outB,
                           uint8_t *state, uint32_t c_off, int lane, uint32_t *X, const uint16_t *PF, int nw, uint32_t kmax) {
    const int sa = (int)ln_size(dA), sb = (int)ln_size(dB);
    const bool ba = ln_is_bits(dA), bb = ln_is_bits(dB);
    if (!ba && !bb) ln_ll_phase2<M64>(inA, sa, inB, sb, outA, outB, tr.toA);
    else if (ba && bb)
        ln_bb_phase2(reinterpret_cast<uint32_t *>(state + dA.x) + lane, reinterpret_cast<uint32_t *>(state + dB.x) + lane, nw,
                     reinterpret_cast<const uint32_t *>(state + c_off) + lane, tr.deal, tr.small_is_a);
    else ln_bl_phase2(X, nw, ba ? inB : inA, ba ? sb : sa, ba ? outB : outA, PF, ba ? inA : inB, ba ? outA : outB, tr.deal, kmax);
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
LN_HD double ln_score(const uint8_t *state, int lane, const uint2 *desc, const int32_t *mod, int k,
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
    for (int wi = 0; wi < nw; ++wi) {
        const uint32_t m = mask[wi];
        uint32_t r[SDX_MAX_K];
        uint32_t U = 0;
#pragma unroll
        for (int j = 0; j < SDX_MAX_K; ++j) {
            uint32_t x = 0;
            if (bp[j]) x = bp[j][(size_t)wi * LN_W];
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
            for (int w = 0; w < wprL; ++w) o[w] = q[(size_t)w * LN_W];
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

// start state of the lane's chain from a compact state (the group layout with one lane: list entries / words contiguous)
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
    uint32_t rows_bytes = 0;     // the rows of the 32 nulls (what a start state holds)
    uint32_t c_off = 0;          // rank mask of bit-packed x bit-packed trades: (nw + 2) x 32 u32
    uint32_t pp_off = 0;         // scorer partials: 2 x 32 x 32 fp64
    uint32_t gsb = 0;            // bytes of one group state: rows, then the scratch above
    int x_bytes = 0;             // shared-memory buffers of one warp for a bit-packed row: the row itself (X) ...
    int pf_bytes = 0;            // ... and its prefix popcounts, one per 4 words (PF4)
};

inline int ln_default_ts(int nw) {
    // a list trade costs ~30 instructions per element of the two rows, a trade with a bit-packed row ~5 per word + ~100 per
    // list entry: short rows are lists
    int ts = nw / 2 + 4;
    if (ts < 16) ts = 16;
    if (ts > LN_TS_MAX) ts = LN_TS_MAX;
    return ts;
}

inline LnLayout ln_build_layout(const int *size, int R, int nw, int ts, uint2 *desc) {
    LnLayout lay;
    if (ts > LN_TS_MAX) ts = LN_TS_MAX;
    lay.ts = ts;
    uint32_t off = 0;
    for (int r = 0; r < R; ++r)
        if (size[r] > ts) { desc[r] = make_uint2(off, (uint32_t)(size[r] > 0xffff ? 0xffff : size[r]) | 0x80000000u); off += (uint32_t)nw * 128u; ++lay.n_bits; }
    for (int r = 0; r < R; ++r)
        if (size[r] <= ts) { desc[r] = make_uint2(off, (uint32_t)size[r]); off += (uint32_t)size[r] * 64u; if (size[r] > lay.max_list) lay.max_list = size[r]; }
    lay.rows_bytes = ((off + 127u) / 128u) * 128u;
    lay.c_off = lay.rows_bytes;
    lay.pp_off = lay.c_off + (lay.n_bits ? (uint32_t)(nw + 2) * 128u : 0u);
    lay.gsb = lay.pp_off + 2u * 32u * 32u * 8u;
    // one staging buffer: the longest list + its sentinel; in a trade with a bit-packed row the unused buffers hold RK / PK (ts u16 each)
    lay.slot_bytes = (lay.max_list + 1) * 64;
    if (lay.n_bits && lay.slot_bytes < ts * 64) lay.slot_bytes = ts * 64;
    lay.x_bytes = lay.n_bits ? nw * 128 : 0;
    lay.pf_bytes = lay.n_bits ? (((nw + 3) / 4 + 1) * 64 + 127) / 128 * 128 : 0;
    return lay;
}
