// tests/tools/lane_host.cu — TEST INFRASTRUCTURE.  Runs the per-lane code of k_trade_lanes (csrc/sdx_lane_trade.cuh, the
// very functions the kernel calls) on the CPU, lane by lane, with the TMA staging replaced by memcpy.  It lets the
// `-m "not gpu"` suite check the one-thread-per-null trade logic, the lane layout, the schedule flags and the scorer
// bit for bit against the oracle without a GPU.  Built by tests/test_lane_host.py with `nvcc -x cu` as host code.
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../superdendrix_b200/csrc/sdx_lane_trade.cuh"

static int popc_row(const uint32_t *r, int nw) { int c = 0; for (int i = 0; i < nw; ++i) c += __builtin_popcount(r[i]); return c; }

// bit-packed state (R x nw) -> compact state, as k_ln_compact does
static void compact_state(const uint32_t *rows, int R, int nw, const uint2 *desc, uint8_t *out) {
    for (int r = 0; r < R; ++r) {
        const uint2 d = desc[r];
        uint8_t *dst = out + d.x / LN_W;
        const uint32_t *src = rows + (size_t)r * nw;
        if (ln_is_bits(d)) { memcpy(dst, src, sizeof(uint32_t) * nw); continue; }
        uint16_t *list = reinterpret_cast<uint16_t *>(dst);
        int pos = 0;
        for (int w = 0; w < nw; ++w)
            for (uint32_t x = src[w]; x; x &= x - 1) list[pos++] = (uint16_t)(w * 32 + __builtin_ctz(x));
    }
}

// One group of 32 chains (chain0 .. chain0 + 31, chain0 a multiple of 32), `rounds` rounds of T trades from the given
// start states (n_start states of R x nw words, chain c starts from state c % n_start).
//   out      rounds x 32 x R x nw   the null after every round
//   applied  rounds x 32
//   W        rounds x 32            SuperW of `mod` (k features) on every null, or NULL
//   stats    [0] trades that shared a row with the trade before (deferred loads), [1] with the 2nd / 3rd before, [2] trades
extern "C" int lane_host_run(const uint32_t *start, int n_start, int R, int nw, int ts, uint64_t seed, uint64_t chain0,
                             int64_t chain_len, int64_t round0, int rounds, int64_t T, uint32_t *out, int64_t *applied,
                             const int32_t *mod, int k, const double *w, const uint32_t *mask, int S, double *W, int64_t *stats) {
    if (chain0 % LN_W) return -1;
    std::vector<int> size((size_t)R);
    for (int r = 0; r < R; ++r) size[r] = popc_row(start + (size_t)r * nw, nw);
    std::vector<uint2> desc((size_t)R);
    if (ts <= 0) ts = ln_default_ts(nw);
    const LnLayout lay = ln_build_layout(size.data(), R, nw, ts, desc.data());
    const bool m64 = 2 * lay.ts > 32;
    const PhiloxKeys keys = make_keys(seed);
    std::vector<uint8_t> state(lay.gsb, 0), compact(lay.rows_bytes / LN_W, 0);
    for (int lane = 0; lane < LN_W; ++lane) {
        const uint64_t ch = chain0 + (uint64_t)lane;
        std::fill(compact.begin(), compact.end(), 0);
        compact_state(start + (size_t)(ch % (uint64_t)n_start) * R * nw, R, nw, desc.data(), compact.data());
        ln_load_state(state.data(), lane, desc.data(), R, nw, compact.data());
    }
    // staging buffers exactly as the kernel sizes them, followed by canaries: an overrun fails the run instead of passing silently
    const size_t SLOT = (size_t)lay.slot_bytes, PFB = (size_t)lay.pf_bytes, XB = (size_t)lay.x_bytes;
    std::vector<uint8_t> inA(SLOT + 64), inB(SLOT + 64), outA(SLOT + 64), outB(SLOT + 64), pf(PFB + 64), xb(XB + 64);
    auto guard_ok = [&](const std::vector<uint8_t> &v, size_t n, uint8_t fill) { for (size_t q = n; q < v.size(); ++q) if (v[q] != fill) return false; return true; };
    double *PP = reinterpret_cast<double *>(state.data() + lay.pp_off), *PN = PP + 32 * LN_W;
    std::vector<uint32_t> ent((size_t)T + 1);
    for (int j = 0; j < rounds; ++j) {
        const uint64_t round = (uint64_t)round0 + (uint64_t)j;
        const uint64_t pid = chain0 * (uint64_t)chain_len + round;
        for (int64_t t = 0; t < T; ++t) {
            uint32_t a, b;
            trade_pair(keys, pid, (uint32_t)t, (uint32_t)R, a, b);
            uint32_t e = a | (b << 14), fl = 0;
            for (int d = 1; d <= 3; ++d)
                if (t >= d && ln_entry_rows_share(e, ent[t - d])) fl |= 1u << (d - 1);
            ent[t] = e | (fl << 28);
            if (stats) { stats[0] += fl & 1u; stats[1] += (fl & 6u) != 0; stats[2] += 1; }
        }
        std::vector<int64_t> ap(LN_W, 0);
        for (int64_t t = 0; t < T; ++t) {
            const uint32_t a = ent[t] & 0x3fffu, b = (ent[t] >> 14) & 0x3fffu;
            const uint2 dA = desc[a], dB = desc[b];
            const bool ba = ln_is_bits(dA), bb = ln_is_bits(dB);
            // "bulk loads": the lists, exactly their bytes; the bit-packed row of a bit-packed x list trade -> X
            memset(inA.data(), 0xAB, inA.size()); memset(inB.data(), 0xAB, inB.size()); memset(pf.data(), 0xEF, pf.size()); memset(xb.data(), 0x77, xb.size());
            if (ba != bb) memcpy(xb.data(), state.data() + (ba ? dA.x : dB.x), (size_t)nw * 128);
            memcpy(inA.data(), state.data() + dA.x, ln_stage_bytes(dA));
            memcpy(inB.data(), state.data() + dB.x, ln_stage_bytes(dB));
            memset(outA.data(), 0xCD, outA.size()); memset(outB.data(), 0xCD, outB.size());
            std::vector<LnTrade<false>> tr32(LN_W);
            std::vector<LnTrade<true>> tr64(LN_W);
            uint32_t kmax = 0;
            auto ptrs = [&](int lane, uint16_t *&pinA, uint16_t *&pinB, uint16_t *&poA, uint16_t *&poB, uint16_t *&PF, uint32_t *&X) {
                pinA = reinterpret_cast<uint16_t *>(inA.data()) + lane; pinB = reinterpret_cast<uint16_t *>(inB.data()) + lane;
                poA = reinterpret_cast<uint16_t *>(outA.data()) + lane; poB = reinterpret_cast<uint16_t *>(outB.data()) + lane;
                PF = reinterpret_cast<uint16_t *>(pf.data()) + lane;
                X = reinterpret_cast<uint32_t *>(xb.data()) + lane;
            };
            // phase 1 of all lanes (reads only; scratch is per lane: stride 32), the vote, then phase 2 of all lanes
            for (int lane = 0; lane < LN_W; ++lane) {
                const LnRng g{&keys, (chain0 + (uint64_t)lane) * (uint64_t)chain_len + round, (uint32_t)t};
                uint16_t *pinA, *pinB, *poA, *poB, *PF; uint32_t *X;
                ptrs(lane, pinA, pinB, poA, poB, PF, X);
                uint32_t km = 0;
                if (m64) ln_trade_phase1<true>(tr64[lane], dA, dB, pinA, pinB, poA, poB, state.data(), lay.c_off, lane, X, PF, nw, g, km);
                else ln_trade_phase1<false>(tr32[lane], dA, dB, pinA, pinB, poA, poB, state.data(), lay.c_off, lane, X, PF, nw, g, km);
                if (km > kmax) kmax = km;
            }
            if (kmax != 0) {
                for (int lane = 0; lane < LN_W; ++lane) {
                    uint16_t *pinA, *pinB, *poA, *poB, *PF; uint32_t *X;
                    ptrs(lane, pinA, pinB, poA, poB, PF, X);
                    if (m64) ln_trade_phase2<true>(tr64[lane], dA, dB, pinA, pinB, poA, poB, state.data(), lay.c_off, lane, X, PF, nw, kmax);
                    else ln_trade_phase2<false>(tr32[lane], dA, dB, pinA, pinB, poA, poB, state.data(), lay.c_off, lane, X, PF, nw, kmax);
                    ap[lane] += (m64 ? tr64[lane].deal.go : tr32[lane].deal.go) ? 1 : 0;
                }
                if (!guard_ok(inA, SLOT, 0xAB) || !guard_ok(inB, SLOT, 0xAB) || !guard_ok(outA, SLOT, 0xCD) || !guard_ok(outB, SLOT, 0xCD) ||
                    !guard_ok(pf, PFB, 0xEF) || !guard_ok(xb, XB, 0x77))
                    return -3 - (!guard_ok(inA, SLOT, 0xAB) ? 10 : 0) - (!guard_ok(inB, SLOT, 0xAB) ? 20 : 0) - (!guard_ok(outA, SLOT, 0xCD) ? 40 : 0) - (!guard_ok(outB, SLOT, 0xCD) ? 80 : 0) - (!guard_ok(pf, PFB, 0xEF) ? 160 : 0) - (!guard_ok(xb, XB, 0x77) ? 320 : 0);   // staging / scratch overrun
                // "bulk stores": the new lists, the updated bit-packed row
                if (ba != bb) memcpy(state.data() + (ba ? dA.x : dB.x), xb.data(), (size_t)nw * 128);
                if (!ba) memcpy(state.data() + dA.x, outA.data(), ln_stage_bytes(dA));
                if (!bb) memcpy(state.data() + dB.x, outB.data(), ln_stage_bytes(dB));
            }
        }
        for (int lane = 0; lane < LN_W; ++lane) {
            uint32_t *dst = out + ((size_t)j * LN_W + lane) * R * nw;
            memset(dst, 0, sizeof(uint32_t) * (size_t)R * nw);
            ln_emit_rows(state.data(), lane, desc.data(), R, nw, dst);
            if (applied) applied[(size_t)j * LN_W + lane] = ap[lane];
            if (W) W[(size_t)j * LN_W + lane] = ln_score(state.data(), lane, desc.data(), mod, k, w, mask, S, PP + lane, PN + lane);
        }
    }
    return 0;
}

extern "C" int lane_host_layout(const int *size, int R, int nw, int ts, uint32_t *desc_xy, int *slot_bytes, uint32_t *gsb, int *n_bits) {
    std::vector<uint2> desc((size_t)R);
    if (ts <= 0) ts = ln_default_ts(nw);
    const LnLayout lay = ln_build_layout(size, R, nw, ts, desc.data());
    for (int r = 0; r < R; ++r) { desc_xy[2 * r] = desc[r].x; desc_xy[2 * r + 1] = desc[r].y; }
    *slot_bytes = lay.slot_bytes; *gsb = lay.gsb; *n_bits = lay.n_bits;
    return lay.ts;
}
