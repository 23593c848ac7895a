// sdx_common.cuh — context, error handling and device helpers shared by the
// libsdx.so translation units (sm_100a only; no CPU fallback anywhere).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/sdx.h"

#define SDX_VERSION 100

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct sdx_ctx {
    int device = 0;
    int sm_count = 148;
    size_t smem_optin = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    sdx_timing timing{};
    int64_t total_launches = 0;

    // matrix
    int F = 0, S = 0;
    int wprS = 0;            // words per feature row (over samples)
    int wprF = 0;            // words per sample row (over features)
    uint32_t *d_feat = nullptr;   // F x wprS
    uint32_t *d_samp = nullptr;   // S x wprF   (device-side transpose)
    // trade orientation
    int R = 0, L = 0, wprL = 0, RS = 0;   // RS: padded row stride (words, multiple of 4)
    int rows_are_samples = 0;
    uint32_t *d_trade_obs = nullptr;      // R x RS, observed matrix in trade orientation
    int *d_rowsize = nullptr;             // R: popcount of each trade row (invariant under Curveball)
    // lane layout of the trade rows (sdx_lanes.cu): 0 = not built yet, 1 = built, -1 = not applicable to this matrix
    int ln_state = 0;
    int ln_ts = 0;                        // rows of at most ln_ts elements are index lists, the others stay bit-packed
    int ln_slot_bytes = 0;                // one staging buffer of the lane kernel (largest row of a 32-null group)
    int ln_has_bits = 0;
    uint32_t ln_gsb = 0;                  // bytes of the state of one group of 32 nulls: rows, then scratch (rank mask, scorer partials)
    uint32_t ln_rows_bytes = 0, ln_c_off = 0, ln_pp_off = 0;      // LnLayout: the rows of a group state, offsets of its scratch
    int ln_x_bytes = 0, ln_pf_bytes = 0;  // shared-memory buffers of one warp for a bit-packed row: the row, its prefix popcounts
    uint2 *d_ln_desc = nullptr;           // R row descriptors: (byte offset in the group state, size | kind << 31)
    uint8_t *d_ln_obs = nullptr;          // the observed matrix as the rows of one group (32 copies, lanes interleaved)
    uint8_t *d_ln_pool = nullptr;         // the burn-in pool: pool_size / 32 groups (ln_pool_groups), else pool_size compact states
    bool ln_pool_groups = false;
    alignas(8) unsigned char ln_geo[64] = {};   // cached launch geometry of the lane kernel (LnGeom) for ln_geo_groups groups
    int64_t ln_geo_groups = -1;
    bool ln_geo_valid = false;
    size_t cap_ln_pool = 0, cap_ln_desc = 0, cap_ln_obs = 0;      // grow-only: a re-upload of the same shape never calls the allocator
    bool ln_pool_valid = false;
    // how the PAIR stream is shared (sdx_set_pair_group): 1 = every null draws its own row pairs; 32 = the 32 chains
    // c, c+1, .. of an aligned block draw them from the stream of the block's first chain
    int pair_group = 1;
    // burn-in pool (sdx_set_burn_in): pool_size well-mixed start states, each the observed matrix after `burn` rounds
    int64_t burn = 0;
    int pool_size = 0;
    uint32_t *d_pool = nullptr;
    size_t pool_state_words = 0;
    bool pool_valid = false;
    uint64_t pool_seed = 0;
    int64_t pool_T = 0;
    // Burn orientation (pool_begin): the burn-in rounds trade the rows of the orientation whose largest row
    // is the smaller multiple of the mean row — the trade orientation itself (burn_alt = 0) or its transpose.
    int burn_alt = 0;
    int bR = 0, bwpr = 0, bRS = 0;        // rows, words per row, padded row stride of the burn orientation
    int64_t burn_T = 0;                   // trades per burn-in round: 5 * bR (at most 60000)
    uint32_t *d_burn_obs = nullptr;       // bR x bRS (burn_alt only; otherwise d_trade_obs serves)
    int *d_burn_rowsize = nullptr;        // bR
    size_t cap_burn_obs = 0, cap_burn_rowsize = 0;
    uint16_t *d_top = nullptr;    // the top_K densest rows of the burn orientation (size descending, index ascending): pair set of the
                                  // dense-favouring burn-in trades; top_K = 0: uniform pairs
    size_t cap_top = 0;
    int top_K = 0;

    // a request inside one long chain continues where the previous call stopped (run_nulls_bits)
    struct Resume {
        bool valid = false, use_pool = false;
        uint64_t seed = 0, chain = 0;
        int64_t T = 0, chain_len = 0, round = 0;         // round: rounds of the chain already applied to d_state
        int pair_group = 1;
        size_t words = 0, cap = 0;
        uint32_t *d_state = nullptr;
    } resume;

    // weights
    int P = 0;
    double *d_w = nullptr;        // P x S  (0 where masked out)
    uint32_t *d_mask = nullptr;   // P x wprS (always materialised)
    int *d_nsamp = nullptr;       // P: popcount of the mask

    // reusable scratch
    DevBuf scratch[20];
    // capacities (bytes) of the grow-only matrix / weight / pool buffers: re-uploads of the same shape never touch the allocator
    size_t cap_feat = 0, cap_samp = 0, cap_obs = 0, cap_rowsize = 0, cap_pool = 0, cap_w = 0, cap_mask = 0, cap_nsamp = 0;
    // tuning knobs (benchmarks and tests only), read from the environment ONCE in sdx_create — never on the call path
    struct Knobs {
        int trade_g = 0, trade_warps = 0, trade_lb = -1, plan_cap = 0, grid_cap = 0;
        long nulls_per_launch = 0, pad_smem = 0;
        int lanes = 1;                    // 0: never use the lane kernel (pair_group 32 then runs on k_trade)
        int lane_ts = 0, lane_wpc = 0, lane_min_chains = 0, lane_debug = 0;
    } knobs;
};

inline int sdx_fail(sdx_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}

extern std::string g_sdx_create_error;

#define SDX_CUDA(ctx, call)                                                                  \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            char b_[512];                                                                    \
            snprintf(b_, sizeof b_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),  \
                     __FILE__, __LINE__);                                                    \
            return sdx_fail((ctx), SDX_ERR_CUDA, b_);                                        \
        }                                                                                    \
    } while (0)

#define SDX_REQUIRE(ctx, cond, code, msg)                     \
    do {                                                      \
        if (!(cond)) return sdx_fail((ctx), (code), (msg));   \
    } while (0)

// grow-only scratch buffer
inline int sdx_scratch(sdx_ctx *c, int slot, size_t bytes, void **out) {
    DevBuf &b = c->scratch[slot];
    if (b.bytes < bytes) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr; b.bytes = 0;
        SDX_CUDA(c, cudaMalloc(&b.p, bytes));
        b.bytes = bytes;
    }
    *out = b.p;
    return SDX_OK;
}

// grow-only typed buffer with its capacity kept in the context
template <typename T>
inline int sdx_reserve(sdx_ctx *c, T *&p, size_t &cap, size_t bytes) {
    if (cap < bytes || !p) {
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        SDX_CUDA(c, cudaMalloc(&p, bytes));
        cap = bytes;
    }
    return SDX_OK;
}
#define SDX_TRY(call) do { int rc_ = (call); if (rc_ != SDX_OK) return rc_; } while (0)

struct LaunchTimer {      // CUDA-event bracket on the ctx stream
    sdx_ctx *c;
    explicit LaunchTimer(sdx_ctx *ctx) : c(ctx) {
        memset(&c->timing, 0, sizeof c->timing);
        cudaEventRecord(c->ev[0], c->stream);
    }
    void count(int n = 1) { c->timing.launches += n; c->total_launches += n; }
    int finish() {
        cudaEventRecord(c->ev[1], c->stream);
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return sdx_fail(c, SDX_ERR_CUDA, std::string("kernel execution failed: ") + cudaGetErrorString(e));
        cudaEventElapsedTime(&c->timing.total_ms, c->ev[0], c->ev[1]);
        return SDX_OK;
    }
};

// ---------------------------------------------------------------------------
// Philox4x32-10 and the canonical streams (oracle/sdx_oracle.c)
// ---------------------------------------------------------------------------
#define SDX_STREAM_PAIR 0u
#define SDX_STREAM_DEAL 1u
#define SDX_STREAM_PERM 2u

struct PhiloxKeys {      // the ten round keys, precomputed on the host so they sit in the constant bank
    uint32_t k0[10];
    uint32_t k1[10];
};

inline PhiloxKeys make_keys(uint64_t seed) {
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = a; k.k1[r] = b;
        a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    return k;
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, const PhiloxKeys &k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c.x;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c.z;
        uint4 n;
        n.x = (uint32_t)(p1 >> 32) ^ c.y ^ k.k0[r];
        n.y = (uint32_t)p1;
        n.z = (uint32_t)(p0 >> 32) ^ c.w ^ k.k1[r];
        n.w = (uint32_t)p0;
        c = n;
    }
    return c;
}

__host__ __device__ __forceinline__ uint4 stream_block(const PhiloxKeys &k, uint32_t kind, uint64_t id, uint32_t t, uint32_t block) {
    uint4 c;
    c.x = block; c.y = t; c.z = (uint32_t)id;
    c.w = ((uint32_t)(id >> 32) & 0x3fffffffu) | (kind << 30);
    return philox4x32_10(c, k);
}

__host__ __device__ __forceinline__ uint32_t pick4(const uint4 &v, uint32_t i) {
    uint32_t lo = (i & 1) ? v.y : v.x;
    uint32_t hi = (i & 1) ? v.w : v.z;
    return (i & 2) ? hi : lo;
}

// Sequential per-thread view of one stream (used where a single thread owns a
// whole stream: trade pairs, permutation draws).
struct ThreadStream {
    const PhiloxKeys &k;
    uint32_t kind; uint64_t id; uint32_t t;
    uint32_t block; uint32_t pos; uint4 buf;
    __host__ __device__ __forceinline__ ThreadStream(const PhiloxKeys &keys, uint32_t kind_, uint64_t id_, uint32_t t_)
        : k(keys), kind(kind_), id(id_), t(t_), block(0), pos(4) {}
    __host__ __device__ __forceinline__ uint32_t next() {
        if (pos == 4) { buf = stream_block(k, kind, id, t, block++); pos = 0; }
        return pick4(buf, pos++);
    }
    // exactly uniform in [0, m)  (Lemire, with rejection)
    __host__ __device__ __forceinline__ uint32_t below(uint32_t m) {
        uint64_t prod = (uint64_t)next() * m;
        uint32_t lo = (uint32_t)prod;
        if (lo < m) {
            uint32_t thr = (0u - m) % m;
            while (lo < thr) { prod = (uint64_t)next() * m; lo = (uint32_t)prod; }
        }
        return (uint32_t)(prod >> 32);
    }
};

__host__ __device__ __forceinline__ void trade_pair(const PhiloxKeys &k, uint64_t null_id, uint32_t t, uint32_t R, uint32_t &a, uint32_t &b) {
    ThreadStream s(k, SDX_STREAM_PAIR, null_id, t);
    a = s.below(R);
    b = s.below(R - 1);
    if (b >= a) b++;
}

// Pair of a BURN-IN trade when even the burn orientation has dominant rows (oracle: trade_pair_burn; the
// caller uses the plain trade_pair when K < 2): with probability 1/4 both rows come from the K densest rows (top[]), otherwise
// the pair is uniform.  A fixed pair distribution (row sizes never change), so the stationary distribution stays uniform.
__device__ __forceinline__ void trade_pair_burn(const PhiloxKeys &k, uint64_t null_id, uint32_t t, uint32_t R,
                                                const uint16_t *__restrict__ top, uint32_t K, uint32_t &a, uint32_t &b) {
    ThreadStream s(k, SDX_STREAM_PAIR, null_id, t);
    const uint32_t u = s.next();
    if (K >= 2 && u < 0x40000000u) {
        const uint32_t i = s.below(K);
        uint32_t j = s.below(K - 1);
        if (j >= i) j++;
        a = top[i]; b = top[j];
    } else {
        a = s.below(R);
        b = s.below(R - 1);
        if (b >= a) b++;
    }
}
#define SDX_POOL_BASE (1ull << 61)      // burn-in round r of pool state j has stream id (B + j) * burn + r, B = 2^61 / burn rounded up to a multiple of 32 ...
#define SDX_BURN_ID_MIN (SDX_POOL_BASE - 256)   // ... so (burn <= 256) every id from here on is a burn-in round; null indices stay below
static inline uint64_t sdx_pool_base_chain(int64_t burn) { return ((SDX_POOL_BASE / (uint64_t)burn + 31) / 32) * 32; }

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum_f64(double v) {      // fixed butterfly: deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int warp_sum_i32(int v) { return __reduce_add_sync(0xffffffffu, v); }

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }
