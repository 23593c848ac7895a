// sdx_trade_sm.cuh — k_trade_sm: experimental persistent work-queue variant of k_trade (SDX_TRADE_PERSIST=1).
// Included by sdx_curveball.cu only.
#pragma once
#include "sdx_trade_kernel.cuh"

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

