// sdx_scan.cu — exhaustive scan of all feature subsets of size <= 3 under W.
//
// Replaces (certifies / seeds) the integral optimum of the ILP of
// src/superdendrix.py:294-313, 393-401 (optimize_model / build_obj): the ILP's
// objective at an integral point equals SuperW (src/superdendrix.py:387-391).
//
// Algebra, with w+ = max(w, 0):
//   s_g   = 2 sum_{s in g} w+_s - sum_{s in g} |w_s|
//   T_gh  = -2 sum_{s in g∩h} w+_s                       (pair table, upper triangle)
//   P_ghl =    sum_{s in g∩h∩l} w+_s                     (only when g∩h∩{w>0} is non-empty)
//   W{g} = s_g,  W{g,h} = s_g + s_h + T_gh,
//   W{g,h,l} = s_g + s_h + s_l + T_gh + T_gl + T_hl + 2 P_ghl
//
// Pipeline (all on the ctx stream):
//   k_scan_prep     w+ vector, packed {w>0} mask, s_g
//   k_pair_table    one CTA per g: T[g][h] for h > g (bit-AND + weighted bit walk)
//   k_scan12        singles and pairs
//   k_scan3         persistent CTAs pull (g-tile, h-block, l-chunk) items from an atomic
//                   counter; a thread owns NJ values of l with c[gi][j] = s_l + T[g_i][l] in
//                   registers, streams row h of T (coalesced, L2-resident) and evaluates
//                   GT x NJ triples per h; every triple of the enumeration is evaluated.
//   k_scan_collect  per-warp top lists -> candidates above the global bound
// The candidates are re-scored with the canonical scorer (score_module) so the reported W,
// coverage and tie order do not depend on the inclusion-exclusion arithmetic.
#include <float.h>

#include <algorithm>

#include "sdx_trade.cuh"

#define SCAN_NT 256
#define SCAN_GT 4
#define SCAN_NJ 4
#define SCAN_HB 128
#define SCAN_E 2
#define SCAN_NL (32 * SCAN_E)
#define SCAN_MAXW 256          // words per feature row the triple slow path stages in shared memory
#define SCAN_KEY_NONE 0xffffffffffffffffull

int sdx_score_sets_device(sdx_ctx *c, const int32_t *d_sets, const int32_t *d_prof, int profile0, int64_t n, int k,
                          double *d_W, int32_t *d_cov, int32_t *d_ovl, int32_t *d_act);

// order-preserving map double -> u64 (for atomicMax on a lower bound)
__host__ __device__ __forceinline__ unsigned long long enc_f64(double v) {
    unsigned long long u;
    memcpy(&u, &v, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dec_f64(unsigned long long e) {
    unsigned long long u = (e >> 63) ? (e ^ 0x8000000000000000ull) : ~e;
    double v;
    memcpy(&v, &u, 8);
    return v;
}

__host__ __device__ __forceinline__ unsigned long long scan_key(int size, int g, int h, int l) {
    // ascending key == enumeration order (size-major, then lexicographic)
    return ((unsigned long long)size << 60) | ((unsigned long long)g << 40) | ((unsigned long long)h << 20) | (unsigned long long)l;
}

// ---------------------------------------------------------------------------
// Per-warp list of the SCAN_NL best (W, key) seen, one/two entries per lane, in registers.
// All 32 lanes call every method together.
// ---------------------------------------------------------------------------
struct WarpTop {
    double W[SCAN_E];
    unsigned long long key[SCAN_E];
    double worstW;
    unsigned long long worstKey;
    int worstSlot;
    double gate;          // candidates need W >= gate

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int e = 0; e < SCAN_E; ++e) { W[e] = -DBL_MAX; key[e] = SCAN_KEY_NONE; }
        worstW = -DBL_MAX; worstKey = SCAN_KEY_NONE; worstSlot = 0; gate = -DBL_MAX;
    }
    __device__ __forceinline__ void find_worst(int lane) {
        double w = W[0]; unsigned long long k = key[0]; int s = lane * SCAN_E;
#pragma unroll
        for (int e = 1; e < SCAN_E; ++e)
            if (W[e] < w || (W[e] == w && key[e] > k)) { w = W[e]; k = key[e]; s = lane * SCAN_E + e; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            double ow = __shfl_xor_sync(0xffffffffu, w, o);
            unsigned long long ok = __shfl_xor_sync(0xffffffffu, k, o);
            int os = __shfl_xor_sync(0xffffffffu, s, o);
            bool take = ow < w || (ow == w && (ok > k || (ok == k && os < s)));
            if (take) { w = ow; k = ok; s = os; }
        }
        worstW = w; worstKey = k; worstSlot = s;
        if (w > gate) gate = w;
    }
    __device__ __forceinline__ void insert(double Wc, unsigned long long kc, int lane) {     // warp-uniform arguments
        if (Wc > worstW || (Wc == worstW && kc < worstKey)) {
            if (lane == worstSlot / SCAN_E) {
#pragma unroll
                for (int e = 0; e < SCAN_E; ++e)
                    if (e == worstSlot % SCAN_E) { W[e] = Wc; key[e] = kc; }
            }
            find_worst(lane);
        }
    }
    __device__ __forceinline__ void offer(bool cand, double Wc, unsigned long long kc, int lane) {
        unsigned m = __ballot_sync(0xffffffffu, cand);
        while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            insert(__shfl_sync(0xffffffffu, Wc, src), __shfl_sync(0xffffffffu, kc, src), lane);
        }
    }
    __device__ __forceinline__ void refresh(const unsigned long long *g_thr) {
        double g = dec_f64(*(volatile const unsigned long long *)g_thr);
        if (g > gate) gate = g;
    }
    __device__ __forceinline__ void publish(unsigned long long *g_thr, int lane) {
        if (lane == 0 && worstKey != SCAN_KEY_NONE) atomicMax(g_thr, enc_f64(worstW));
    }
    __device__ __forceinline__ void store(double *oW, unsigned long long *oK, int lane) {
#pragma unroll
        for (int e = 0; e < SCAN_E; ++e) { oW[lane * SCAN_E + e] = W[e]; oK[lane * SCAN_E + e] = key[e]; }
    }
};

// ---------------------------------------------------------------------------
// prep: w+ and the packed {w > 0} mask (one thread per sample), then s_g (one warp per feature)
// ---------------------------------------------------------------------------
__global__ void k_scan_prep_w(const double *__restrict__ w, const uint32_t *__restrict__ mask, int S,
                              double *__restrict__ wpos, uint32_t *__restrict__ pos) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    double v = 0.0;
    if (s < S) { v = w[s]; in = ((mask[s >> 5] >> (s & 31)) & 1u) && v > 0.0; wpos[s] = in ? v : 0.0; }
    const uint32_t b = __ballot_sync(0xffffffffu, in);
    if ((threadIdx.x & 31) == 0 && s < S) pos[s >> 5] = b;
}

__global__ void k_scan_prep_s1(const uint32_t *__restrict__ feat, int fstride, int F, int wprS,
                               const double *__restrict__ w, const uint32_t *__restrict__ mask, double *__restrict__ s1,
                               size_t feat_b, size_t s1_b) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= F) return;
    feat += (size_t)blockIdx.y * feat_b;          // blockIdx.y = matrix of the batch
    s1 += (size_t)blockIdx.y * s1_b;
    double pos = 0, neg = 0;
    for (int wi = lane; wi < wprS; wi += 32) {
        uint32_t x = feat[(size_t)g * fstride + wi] & mask[wi];
        while (x) {
            const int bp = __ffs(x) - 1;
            x &= x - 1;
            const double ws = w[wi * 32 + bp];
            if (ws > 0) pos += ws;
            neg += fabs(ws);
        }
    }
    pos = warp_sum_f64(pos); neg = warp_sum_f64(neg);
    if (lane == 0) s1[g] = 2.0 * pos - neg;
}

// ---------------------------------------------------------------------------
// pair table: T[g][h] = -2 sum_{s in g∩h, w_s>0} w_s for h > g.  One CTA per g.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_table(const uint32_t *__restrict__ feat, int fstride, int F, int wprS,
                                                    const uint32_t *__restrict__ pos, const double *__restrict__ wpos,
                                                    double *__restrict__ T, int Fp, size_t feat_b, size_t T_b) {
    extern __shared__ uint32_t sm[];
    feat += (size_t)blockIdx.y * feat_b;          // blockIdx.y = matrix of the batch
    T += (size_t)blockIdx.y * T_b;
    uint32_t *gp = sm;                 // wprS: feat[g] & pos
    int *nz = (int *)(sm + wprS);      // indices of the non-zero words of gp, ascending
    __shared__ int n_nz;
    for (int g = blockIdx.x; g < F - 1; g += gridDim.x) {
        __syncthreads();
        for (int wi = threadIdx.x; wi < wprS; wi += blockDim.x) gp[wi] = feat[(size_t)g * fstride + wi] & pos[wi];
        __syncthreads();
        if (threadIdx.x < 32) {        // ordered compaction by warp 0
            int base = 0;
            for (int w0 = 0; w0 < wprS; w0 += 32) {
                const int wi = w0 + threadIdx.x;
                const bool nzw = wi < wprS && gp[wi] != 0;
                const unsigned m = __ballot_sync(0xffffffffu, nzw);
                if (nzw) nz[base + __popc(m & ((1u << threadIdx.x) - 1u))] = wi;
                base += __popc(m);
            }
            if (threadIdx.x == 0) n_nz = base;
        }
        __syncthreads();
        const int nn = n_nz;
        for (int h = g + 1 + threadIdx.x; h < F; h += blockDim.x) {
            double acc = 0.0;
            const uint32_t *rh = feat + (size_t)h * fstride;
            for (int q = 0; q < nn; ++q) {
                const int wi = nz[q];
                uint32_t x = gp[wi] & rh[wi];
                while (x) {
                    const int bp = __ffs(x) - 1;
                    x &= x - 1;
                    acc += wpos[wi * 32 + bp];
                }
            }
            T[(size_t)g * Fp + h] = -2.0 * acc;
        }
    }
}

// ---------------------------------------------------------------------------
struct ScanArgs {
    const uint32_t *feat; int fstride;     // F rows over samples
    const uint32_t *samp; int sstride;     // S rows over features
    const uint32_t *pos;                   // packed {w > 0}
    const double *wpos;                    // S
    const double *s1;                      // F
    const double *T; int Fp;
    int F, S, wprS, k;
    int g_lo, g_hi;                        // only subsets whose smallest feature index is in [g_lo, g_hi)
    int gt0, n_gt, n_hb, n_lc;             // g-tiles gt0 .. gt0 + n_gt - 1
    int nb;                                // matrices in the batch (1 for sdx_scan)
    size_t feat_b, samp_b, s1_b, T_b;      // batch strides (elements)
    unsigned long long *best;              // nb: order-preserving encoding of the best W per matrix (LISTS == false)
    unsigned int *item_counter;
    unsigned long long *g_thr;
    unsigned long long *n_scored;
    double *outW; unsigned long long *outKey;      // per-warp lists (k_scan12 lists first, then k_scan3)
};

// best W per matrix of a batch (LISTS == false): per-thread maximum, warp maximum, one atomicMax
__device__ __forceinline__ void publish_best(unsigned long long *slot, double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0 && v > -DBL_MAX) atomicMax(slot, enc_f64(v));
}

// singles and pairs of every matrix of a batch, maximum only: grid (x, nb)
__global__ void __launch_bounds__(SCAN_NT) k_scan12_max(const __grid_constant__ ScanArgs p) {
    const int tid = threadIdx.x, F = p.F, b = blockIdx.y;
    const double *s1 = p.s1 + (size_t)b * p.s1_b;
    const double *T = p.T + (size_t)b * p.T_b;
    double best = -DBL_MAX;
    for (int g = blockIdx.x * SCAN_NT + tid; g < F; g += gridDim.x * SCAN_NT) best = fmax(best, s1[g]);
    if (p.k >= 2) {
        for (int g = blockIdx.x; g < F - 1; g += gridDim.x) {
            const double sg = s1[g];
            const double *Tg = T + (size_t)g * p.Fp;
            for (int h = g + 1 + tid; h < F; h += SCAN_NT) best = fmax(best, (sg + s1[h]) + Tg[h]);
        }
    }
    publish_best(p.best + b, best);
}

// singles and pairs: CTA-stride over g, threads over h
__global__ void __launch_bounds__(SCAN_NT) k_scan12(const __grid_constant__ ScanArgs p) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    WarpTop top;
    top.init();
    unsigned long long cnt = 0;
    const int F = p.F;
    // singles
    for (int g0 = blockIdx.x * SCAN_NT; g0 < F; g0 += gridDim.x * SCAN_NT) {
        const int g = g0 + tid;
        const bool v = g < F && g >= p.g_lo && g < p.g_hi;
        const double W = v ? p.s1[g] : -INFINITY;
        cnt += v;
        top.offer(v && W >= top.gate, W, scan_key(1, v ? g : 0, 0, 0), lane);
    }
    if (p.k >= 2) {
        for (int g = p.g_lo + blockIdx.x; g < min(p.g_hi, F - 1); g += gridDim.x) {
            top.refresh(p.g_thr);
            const double sg = p.s1[g];
            const double *Tg = p.T + (size_t)g * p.Fp;
            for (int h0 = g + 1; h0 < F; h0 += SCAN_NT) {
                const int h = h0 + tid;
                const bool v = h < F;
                const double W = v ? (sg + p.s1[h]) + Tg[h] : -INFINITY;
                cnt += v;
                const bool cand = v && W >= top.gate;
                if (__any_sync(0xffffffffu, cand)) top.offer(cand, W, scan_key(2, g, v ? h : 0, 0), lane);
            }
            top.publish(p.g_thr, lane);
        }
    }
    top.publish(p.g_thr, lane);
    const size_t slot = (size_t)(blockIdx.x * (SCAN_NT / 32) + wid) * SCAN_NL;
    top.store(p.outW + slot, p.outKey + slot, lane);
    cnt = __reduce_add_sync(0xffffffffu, (unsigned)cnt);       // per-thread counts are small
    if (lane == 0 && cnt) atomicAdd(p.n_scored, cnt);
}

// LISTS == true: one matrix, per-warp top lists (sdx_scan).  LISTS == false: a batch of nb matrices, only the
// maximum of each is kept (the max_k null statistic); `bp` below then carries the batch offsets.
// NT threads per CTA, NJ values of l per thread: one l-chunk is NT * NJ wide (small NT for small F).
template <bool LISTS, int NT, int NJ>
__global__ void __launch_bounds__(NT, NT >= 256 ? 2 : 5) k_scan3(const __grid_constant__ ScanArgs pa) {
    __shared__ unsigned s_item;
    __shared__ double s_sgh[SCAN_GT][SCAN_HB];
    __shared__ unsigned char s_fl[SCAN_HB];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int F = pa.F, Fp = pa.Fp;
    WarpTop top;
    top.init();
    unsigned long long cnt = 0;
    const int per_gt = pa.n_hb * pa.n_lc;
    const unsigned per_b = (unsigned)pa.n_gt * per_gt;
    const unsigned total = per_b * (unsigned)pa.nb;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_item = atomicAdd(pa.item_counter, 1u);
        __syncthreads();
        unsigned it = s_item;
        if (it >= total) break;
        const int b = it / per_b;
        it -= b * per_b;
        struct { const uint32_t *feat, *samp, *pos; const double *wpos, *s1, *T; int fstride, sstride, wprS, g_lo, g_hi, gt0, n_lc;
                 unsigned long long *g_thr; } p;
        p.feat = pa.feat + (size_t)b * pa.feat_b; p.samp = pa.samp + (size_t)b * pa.samp_b; p.pos = pa.pos; p.wpos = pa.wpos;
        p.s1 = pa.s1 + (size_t)b * pa.s1_b; p.T = pa.T + (size_t)b * pa.T_b; p.fstride = pa.fstride; p.sstride = pa.sstride;
        p.wprS = pa.wprS; p.g_lo = pa.g_lo; p.g_hi = pa.g_hi; p.gt0 = pa.gt0; p.n_lc = pa.n_lc; p.g_thr = pa.g_thr;
        double best = -DBL_MAX;
        const int gt = p.gt0 + it / per_gt, rem = it % per_gt, hb = rem / p.n_lc, lc = rem - hb * p.n_lc;
        const int g0 = gt * SCAN_GT, l0 = lc * (NT * NJ);
        const int hlo = max(hb * SCAN_HB, g0 + 1), hhi = min(hb * SCAN_HB + SCAN_HB, F - 1);
        if (hlo >= hhi || l0 + NT * NJ - 1 <= hlo) continue;
        if (LISTS) top.refresh(p.g_thr);

        double c[SCAN_GT][NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int l = l0 + tid + j * NT;
            const double sl = l < F ? p.s1[l] : 0.0;
#pragma unroll
            for (int gi = 0; gi < SCAN_GT; ++gi) {
                const int g = g0 + gi;
                c[gi][j] = (l < F && g < l && g >= p.g_lo && g < p.g_hi) ? sl + p.T[(size_t)g * Fp + l] : -INFINITY;
            }
        }

        // per-h terms that do not depend on l, once per item: s_g + s_h + T_gh for the tile's 4 g, and which of them
        // have a non-empty g∩h∩{w>0} (bits 0-3) plus the number of valid g (bits 4-6)
        __syncthreads();
        if (tid < SCAN_HB) {
            const int h = hb * SCAN_HB + tid;
            unsigned fl = 0;
            if (h >= hlo && h < hhi) {
                const double sh = p.s1[h];
                int nv = 0;
#pragma unroll
                for (int gi = 0; gi < SCAN_GT; ++gi) {
                    const int g = g0 + gi;
                    const bool v = g < h && g >= p.g_lo && g < p.g_hi;
                    const double tgh = v ? p.T[(size_t)g * Fp + h] : 0.0;
                    s_sgh[gi][tid] = v ? (p.s1[g] + sh) + tgh : -INFINITY;
                    fl |= (v && tgh != 0.0) ? (1u << gi) : 0u;
                    nv += v;
                }
                fl |= (unsigned)nv << 4;
            }
            s_fl[tid] = (unsigned char)fl;
        }
        __syncthreads();

        for (int h = hlo; h < hhi; ++h) {
            double sgh[SCAN_GT];
            bool slow[SCAN_GT];
            const int hi_ = h - hb * SCAN_HB;
            const unsigned fl = s_fl[hi_];
            const bool any_slow = (fl & 15u) != 0;
            const int nvg = (int)(fl >> 4);
#pragma unroll
            for (int gi = 0; gi < SCAN_GT; ++gi) {
                sgh[gi] = s_sgh[gi][hi_];
                slow[gi] = (fl >> gi) & 1u;
            }
            const double *Th = p.T + (size_t)h * Fp;
            double t[NJ];
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int l = l0 + tid + j * NT;
                const bool v = l > h && l < F;
                t[j] = v ? Th[l] : -INFINITY;
                cnt += v ? nvg : 0;
            }
            if (!any_slow) {
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (l0 + (j + 1) * NT - 1 <= h) continue;          // CTA-uniform: every l of this slice is <= h
                    double W[SCAN_GT];
                    bool cand = false;
#pragma unroll
                    for (int gi = 0; gi < SCAN_GT; ++gi) {
                        W[gi] = (sgh[gi] + t[j]) + c[gi][j];
                        if (LISTS) cand |= W[gi] >= top.gate;
                        else best = W[gi] > best ? W[gi] : best;          // no NaNs here: a compare and a select, not the NaN-aware fmax
                    }
                    if (LISTS && __any_sync(0xffffffffu, cand)) {
                        const int l = l0 + tid + j * NT;
#pragma unroll
                        for (int gi = 0; gi < SCAN_GT; ++gi)
                            top.offer(W[gi] >= top.gate, W[gi], scan_key(3, g0 + gi, h, l), lane);
                    }
                }
            } else {
#pragma unroll
                for (int gi = 0; gi < SCAN_GT; ++gi) {
                    if (sgh[gi] == -INFINITY) continue;           // uniform
                    double p3[NJ];
#pragma unroll
                    for (int j = 0; j < NJ; ++j) p3[j] = 0.0;
                    if (slow[gi]) {                               // uniform: g∩h∩{w>0} is non-empty
                        const int g = g0 + gi;
                        // g∩h∩{w>0}: every warp forms the intersection words itself, one per lane and 32 at a time (the rows are L1 / L2
                        // resident), and walks the non-zero ones: no shared-memory staging, no CTA barrier in the slow path, and the
                        // usual handful of samples in one or two words costs one ballot instead of a scan over all wprS words
                        for (int w0 = 0; w0 < p.wprS; w0 += 32) {
                            const int wl = w0 + lane;
                            const uint32_t xi = wl < p.wprS ? p.feat[(size_t)g * p.fstride + wl] & p.feat[(size_t)h * p.fstride + wl] & p.pos[wl] : 0u;
                            for (uint32_t nzm = __ballot_sync(0xffffffffu, xi != 0u); nzm; nzm &= nzm - 1) {
                            const int wsrc = __ffs(nzm) - 1, wi = w0 + wsrc;
                            uint32_t x = __shfl_sync(0xffffffffu, xi, wsrc);
                            while (x) {
                                const int bp = __ffs(x) - 1;
                                x &= x - 1;
                                const int s = wi * 32 + bp;
                                const double wp = p.wpos[s];
                                const uint32_t *sr = p.samp + (size_t)s * p.sstride;
#pragma unroll
                                for (int j = 0; j < NJ; ++j) {
                                    const int l = l0 + tid + j * NT;
                                    if (l < F && ((sr[l >> 5] >> (l & 31)) & 1u)) p3[j] += wp;
                                }
                            }
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        if (l0 + (j + 1) * NT - 1 <= h) continue;      // CTA-uniform
                        const double W = ((sgh[gi] + t[j]) + c[gi][j]) + 2.0 * p3[j];
                        if (LISTS) {
                            const bool cand = W >= top.gate;
                            if (__any_sync(0xffffffffu, cand))
                                top.offer(cand, W, scan_key(3, g0 + gi, h, l0 + tid + j * NT), lane);
                        } else {
                            best = W > best ? W : best;
                        }
                    }
                }
            }
        }
        if (LISTS) top.publish(p.g_thr, lane);
        else publish_best(pa.best + b, best);
    }
    if (LISTS) {
        const size_t slot = (size_t)((gridDim.x + blockIdx.x) * (NT / 32) + wid) * SCAN_NL;      // after the k_scan12 lists
        top.store(pa.outW + slot, pa.outKey + slot, lane);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0 && cnt) atomicAdd(pa.n_scored, cnt);
    }
}

__global__ void k_scan_collect(const double *__restrict__ W, const unsigned long long *__restrict__ key, int n,
                               const unsigned long long *__restrict__ g_thr, unsigned int *__restrict__ n_out,
                               double *__restrict__ oW, unsigned long long *__restrict__ oK) {
    const double thr = dec_f64(*g_thr);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (key[i] != SCAN_KEY_NONE && W[i] >= thr) {
            const unsigned q = atomicAdd(n_out, 1u);
            oW[q] = W[i]; oK[q] = key[i];
        }
    }
}

// ---------------------------------------------------------------------------
extern "C" int sdx_scan_range(sdx_ctx *c, int profile, int k, int topn, int g_lo, int g_hi, sdx_hit *hits, int64_t *n_scored);

extern "C" int sdx_scan(sdx_ctx *c, int profile, int k, int topn, sdx_hit *hits, int64_t *n_scored) {
    if (!c) return SDX_ERR_INVALID;
    return sdx_scan_range(c, profile, k, topn, 0, c->F, hits, n_scored);
}

extern "C" int sdx_scan_range(sdx_ctx *c, int profile, int k, int topn, int g_lo, int g_hi, sdx_hit *hits, int64_t *n_scored) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, profile >= 0 && profile < c->P, SDX_ERR_INVALID, "profile index out of range");
    SDX_REQUIRE(c, k >= 1 && k <= 3, SDX_ERR_UNSUPPORTED, "the exhaustive scan covers k = 1..3");
    SDX_REQUIRE(c, topn >= 1 && topn <= SCAN_NL - 16 && hits, SDX_ERR_INVALID, "topn must be in 1..48");
    SDX_REQUIRE(c, c->wprS <= SCAN_MAXW, SDX_ERR_UNSUPPORTED, "the scan supports at most 8192 samples");
    SDX_REQUIRE(c, c->F < (1 << 20), SDX_ERR_UNSUPPORTED, "the scan supports fewer than 2^20 features");
    SDX_REQUIRE(c, g_lo >= 0 && g_lo <= g_hi && g_hi <= c->F, SDX_ERR_INVALID, "feature range must satisfy 0 <= g_lo <= g_hi <= n_features");
    const int F = c->F, S = c->S, wprS = c->wprS;
    const int Fp = ((F + 3) / 4) * 4;
    const int grid12 = c->sm_count * 2, grid3 = c->sm_count * 2;
    const int n_lists = (grid12 + grid3) * (SCAN_NT / 32);
    const int n_entries = n_lists * SCAN_NL;
    int rc;
    void *p_small, *p_T, *p_lists;
    // small: wpos[S] | s1[F] | pos[wprS] | counters
    const size_t off_s1 = sizeof(double) * S, off_pos = off_s1 + sizeof(double) * F;
    const size_t off_cnt = ((off_pos + sizeof(uint32_t) * wprS + 15) / 16) * 16;
    if ((rc = sdx_scratch(c, 0, off_cnt + 64, &p_small))) return rc;
    if ((rc = sdx_scratch(c, 10, sizeof(double) * (size_t)F * Fp + 64, &p_T))) return rc;
    // lists: W[n_entries] | key[n_entries] | candW[n_entries] | candKey[n_entries] | rescoring buffers
    const int n_resc = SCAN_NL;
    const size_t list_bytes = (sizeof(double) + sizeof(unsigned long long)) * 2 * (size_t)n_entries;
    const size_t resc_bytes = n_resc * (3 * sizeof(int32_t) + sizeof(double) + 3 * sizeof(int32_t));
    if ((rc = sdx_scratch(c, 11, list_bytes + resc_bytes + 64, &p_lists))) return rc;
    double *d_wpos = (double *)p_small, *d_s1 = (double *)((char *)p_small + off_s1);
    uint32_t *d_pos = (uint32_t *)((char *)p_small + off_pos);
    unsigned long long *d_thr = (unsigned long long *)((char *)p_small + off_cnt);
    unsigned long long *d_nscored = d_thr + 1;
    unsigned int *d_item = (unsigned int *)(d_thr + 2), *d_ncand = d_item + 1;
    double *d_T = (double *)p_T;
    double *d_lW = (double *)p_lists;
    unsigned long long *d_lK = (unsigned long long *)(d_lW + n_entries);
    double *d_cW = (double *)(d_lK + n_entries);
    unsigned long long *d_cK = (unsigned long long *)(d_cW + n_entries);
    double *d_rW = (double *)(d_cK + n_entries);
    int32_t *d_rsets = (int32_t *)(d_rW + n_resc), *d_rcov = d_rsets + 3 * n_resc, *d_rovl = d_rcov + n_resc, *d_ract = d_rovl + n_resc;

    LaunchTimer timer(c);
    const double *d_w = c->d_w + (size_t)profile * S;
    const uint32_t *d_mask = c->d_mask + (size_t)profile * wprS;
    struct { unsigned long long thr, nscored; unsigned int item, ncand; } init = {enc_f64(-DBL_MAX), 0ull, 0u, 0u};
    SDX_CUDA(c, cudaMemcpyAsync(d_thr, &init, sizeof init, cudaMemcpyHostToDevice, c->stream));
    cudaEventRecord(c->ev[2], c->stream);
    k_scan_prep_w<<<ceil_div(wprS * 32, 256), 256, 0, c->stream>>>(d_w, d_mask, S, d_wpos, d_pos);
    k_scan_prep_s1<<<ceil_div(F, 8), 256, 0, c->stream>>>(c->d_feat, wprS, F, wprS, d_w, d_mask, d_s1, 0, 0);
    timer.count(2);
    if (k >= 2 && F >= 2) {
        k_pair_table<<<std::min(F - 1, c->sm_count * 8), 256, sizeof(uint32_t) * 2 * wprS, c->stream>>>(
            c->d_feat, wprS, F, wprS, d_pos, d_wpos, d_T, Fp, 0, 0);
        timer.count();
    }
    ScanArgs a{};
    a.feat = c->d_feat; a.fstride = wprS; a.samp = c->d_samp; a.sstride = c->wprF; a.pos = d_pos; a.wpos = d_wpos;
    a.s1 = d_s1; a.T = d_T; a.Fp = Fp; a.F = F; a.S = S; a.wprS = wprS; a.k = k; a.nb = 1;
    a.g_lo = g_lo; a.g_hi = g_hi; a.gt0 = g_lo / SCAN_GT; a.n_gt = g_hi > g_lo ? ceil_div(g_hi, SCAN_GT) - a.gt0 : 0;
    a.n_hb = ceil_div(F, SCAN_HB); a.n_lc = ceil_div(F, SCAN_NT * SCAN_NJ);
    a.item_counter = d_item; a.g_thr = d_thr; a.n_scored = d_nscored; a.outW = d_lW; a.outKey = d_lK;
    SDX_REQUIRE(c, (int64_t)a.n_gt * a.n_hb * a.n_lc < (1ll << 31), SDX_ERR_UNSUPPORTED, "too many scan work items");
    k_scan12<<<grid12, SCAN_NT, 0, c->stream>>>(a);
    timer.count();
    int used_lists = grid12 * (SCAN_NT / 32);
    if (k >= 3 && F >= 3) {
        k_scan3<true, SCAN_NT, SCAN_NJ><<<grid3, SCAN_NT, 0, c->stream>>>(a);
        timer.count();
        used_lists = n_lists;
    }
    k_scan_collect<<<c->sm_count, 256, 0, c->stream>>>(d_lW, d_lK, used_lists * SCAN_NL, d_thr, d_ncand, d_cW, d_cK);
    timer.count();
    cudaEventRecord(c->ev[3], c->stream);
    struct { unsigned long long thr, nscored; unsigned int item, ncand; } fin;
    SDX_CUDA(c, cudaMemcpyAsync(&fin, d_thr, sizeof fin, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t nc = fin.ncand;
    std::vector<double> cW(nc);
    std::vector<unsigned long long> cK(nc);
    if (nc) {
        SDX_CUDA(c, cudaMemcpyAsync(cW.data(), d_cW, sizeof(double) * nc, cudaMemcpyDeviceToHost, c->stream));
        SDX_CUDA(c, cudaMemcpyAsync(cK.data(), d_cK, sizeof(unsigned long long) * nc, cudaMemcpyDeviceToHost, c->stream));
        SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    std::vector<int> order(nc);
    for (size_t i = 0; i < nc; ++i) order[i] = (int)i;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return cW[x] > cW[y] || (cW[x] == cW[y] && cK[x] < cK[y]); });
    const int nr = (int)std::min<size_t>(nc, (size_t)std::min(n_resc, topn + 16));
    std::vector<int32_t> sets(3 * (size_t)n_resc, -1);
    std::vector<unsigned long long> rkey(nr);
    for (int i = 0; i < nr; ++i) {
        const unsigned long long key = cK[order[i]];
        rkey[i] = key;
        const int size = (int)(key >> 60);
        sets[3 * i] = (int)((key >> 40) & 0xfffff);
        if (size >= 2) sets[3 * i + 1] = (int)((key >> 20) & 0xfffff);
        if (size >= 3) sets[3 * i + 2] = (int)(key & 0xfffff);
    }
    std::vector<double> rW(nr);
    std::vector<int32_t> rcov(nr);
    if (nr) {
        SDX_CUDA(c, cudaMemcpyAsync(d_rsets, sets.data(), sizeof(int32_t) * 3 * nr, cudaMemcpyHostToDevice, c->stream));
        if ((rc = sdx_score_sets_device(c, d_rsets, nullptr, profile, nr, 3, d_rW, d_rcov, d_rovl, d_ract))) return rc;
        timer.count();
        SDX_CUDA(c, cudaMemcpyAsync(rW.data(), d_rW, sizeof(double) * nr, cudaMemcpyDeviceToHost, c->stream));
        SDX_CUDA(c, cudaMemcpyAsync(rcov.data(), d_rcov, sizeof(int32_t) * nr, cudaMemcpyDeviceToHost, c->stream));
    }
    rc = timer.finish();
    if (rc) return rc;
    cudaEventElapsedTime(&c->timing.score_ms, c->ev[2], c->ev[3]);
    std::vector<int> ro(nr);
    for (int i = 0; i < nr; ++i) ro[i] = i;
    std::sort(ro.begin(), ro.end(), [&](int x, int y) { return rW[x] > rW[y] || (rW[x] == rW[y] && rkey[x] < rkey[y]); });
    for (int i = 0; i < topn; ++i) {
        if (i < nr) {
            const int q = ro[i];
            hits[i].W = rW[q]; hits[i].coverage = rcov[q];
            for (int j = 0; j < 3; ++j) hits[i].features[j] = sets[3 * q + j];
        } else {
            hits[i].W = -INFINITY; hits[i].coverage = 0;
            hits[i].features[0] = hits[i].features[1] = hits[i].features[2] = -1;
        }
    }
    if (n_scored) *n_scored = (int64_t)fin.nscored;
    return SDX_OK;
}

// ---------------------------------------------------------------------------
// The reference-faithful null statistic (src/superdendrix.py:191): for every null the
// optimum over all modules of size <= k — what the reference obtains by re-running the ILP on the null.
// Batches of nulls are generated into device memory, transposed to the other orientation, and scanned
// with the batched kernels above; the empty module is feasible for the ILP, so the statistic is >= 0.
// ---------------------------------------------------------------------------
int sdx_generate_to_device(sdx_ctx *ctx, uint64_t seed, uint64_t null0, int64_t n, int64_t n_trades, int64_t chain_len,
                           uint32_t *d_out);
int sdx_transpose_batch(sdx_ctx *c, const uint32_t *d_in, int n_rows, int wpr_in, int n_cols, uint32_t *d_out, int wpr_out, int nb);

__global__ void k_maxk_finalize(const unsigned long long *__restrict__ best, int nb, const double *__restrict__ z_obs,
                                unsigned long long *__restrict__ n_ge, double *__restrict__ scores) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool ge = false;
    if (b < nb) {
        double v = dec_f64(best[b]);
        v = v > 0.0 ? v : 0.0;                 // M = {} is feasible: z >= 0
        if (scores) scores[b] = v;
        ge = z_obs && v >= *z_obs;
    }
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_ge, (unsigned long long)__popc(m));
}

// rows_src == nullptr: the nulls are generated (sdx_null_pvalue); otherwise they are the caller's n_nulls x F x wprS packed
// feature rows, host or device memory (sdx_null_pvalue_rows: the reference's loop over null files, src/superdendrix.py:175-192)
static int maxk_impl(sdx_ctx *c, const uint32_t *rows_src, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                     const int32_t *modules, int k, const double *z_obs, int64_t *n_ge, double *null_scores, double *z_obs_out) {
    SDX_REQUIRE(c, c->d_trade_obs && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, modules && n_ge && k >= 1 && k <= SDX_MAX_K, SDX_ERR_INVALID, "modules / n_ge / k invalid");
    SDX_REQUIRE(c, n_nulls >= 0 && chain_len >= 1, SDX_ERR_INVALID, "n_nulls must be >= 0 and chain_len >= 1");
    SDX_REQUIRE(c, c->wprS <= SCAN_MAXW && c->F < (1 << 20), SDX_ERR_UNSUPPORTED, "matrix too large for the scan");
    const int F = c->F, S = c->S, wprS = c->wprS, wprF = c->wprF, P = c->P;
    std::vector<int> kp(P);
    for (int p = 0; p < P; ++p) {               // k of profile p = size of its module (src/superdendrix.py:173)
        int cnt = 0;
        for (int j = 0; j < k; ++j) cnt += modules[(size_t)p * k + j] >= 0;
        SDX_REQUIRE(c, cnt >= 1 && cnt <= 3, SDX_ERR_UNSUPPORTED, "max_k re-optimises modules of 1..3 features");
        kp[p] = cnt;
    }
    const int Fp = ((F + 3) / 4) * 4;
    const size_t feat_words = (size_t)F * wprS, samp_words = (size_t)S * wprF;
    const size_t per_null = sizeof(uint32_t) * (feat_words + samp_words) + sizeof(double) * ((size_t)F * Fp + F) + 16;
    int64_t B = (int64_t)(((size_t)3 << 29) / per_null);          // ~1.5 GB of scratch
    if (B < 1) B = 1;
    if (B > 2048) B = 2048;
    if (B > n_nulls && n_nulls > 0) B = n_nulls;
    const int grid3 = c->sm_count * 2;
    int rc;
    void *p_small, *p_mat, *p_T, *p_sc = nullptr;
    // small: wpos[S] | pos[wprS] | z_obs[P] | n_ge[P] | best[B] | item counter
    const size_t off_pos = sizeof(double) * S, off_z = ((off_pos + sizeof(uint32_t) * wprS + 15) / 16) * 16;
    const size_t off_ge = off_z + sizeof(double) * P, off_best = off_ge + sizeof(unsigned long long) * P;
    const size_t off_item = off_best + sizeof(unsigned long long) * B;
    if ((rc = sdx_scratch(c, 14, off_item + 64, &p_small))) return rc;
    if ((rc = sdx_scratch(c, 11, sizeof(uint32_t) * (feat_words + samp_words) * B + sizeof(double) * (size_t)F * B + 64, &p_mat))) return rc;
    if ((rc = sdx_scratch(c, 10, sizeof(double) * (size_t)F * Fp * B + 64, &p_T))) return rc;
    if (null_scores && n_nulls > 0 && (rc = sdx_scratch(c, 15, sizeof(double) * (size_t)P * n_nulls + 64, &p_sc))) return rc;
    double *d_wpos = (double *)p_small;
    uint32_t *d_pos = (uint32_t *)((char *)p_small + off_pos);
    double *d_z = (double *)((char *)p_small + off_z);
    unsigned long long *d_ge = (unsigned long long *)((char *)p_small + off_ge);
    unsigned long long *d_best = (unsigned long long *)((char *)p_small + off_best);
    unsigned int *d_item = (unsigned int *)((char *)p_small + off_item);
    // trade-orientation rows arrive dense; the other orientation is produced by the batched transpose
    uint32_t *d_a = (uint32_t *)p_mat;                              // as generated: B x R x wprL
    uint32_t *d_b = d_a + (c->rows_are_samples ? samp_words : feat_words) * B;
    double *d_s1 = (double *)(d_b + (c->rows_are_samples ? feat_words : samp_words) * B);
    uint32_t *d_featB = c->rows_are_samples ? d_b : d_a, *d_sampB = c->rows_are_samples ? d_a : d_b;
    double *d_scores = (double *)p_sc;

    // the scan of a batch of nb matrices for profile p: best[b] = max over |M| <= kp[p]
    auto scan_batch = [&](int p, const uint32_t *feat, size_t feat_b, const uint32_t *samp, size_t samp_b, int nb) -> int {
        const double *d_w = c->d_w + (size_t)p * S;
        const uint32_t *d_mask = c->d_mask + (size_t)p * wprS;
        std::vector<unsigned long long> init((size_t)nb, enc_f64(-DBL_MAX));
        SDX_CUDA(c, cudaMemcpyAsync(d_best, init.data(), sizeof(unsigned long long) * nb, cudaMemcpyHostToDevice, c->stream));
        SDX_CUDA(c, cudaMemsetAsync(d_item, 0, sizeof(unsigned int), c->stream));
        k_scan_prep_w<<<ceil_div(wprS * 32, 256), 256, 0, c->stream>>>(d_w, d_mask, S, d_wpos, d_pos);
        k_scan_prep_s1<<<dim3(ceil_div(F, 8), nb), 256, 0, c->stream>>>(feat, wprS, F, wprS, d_w, d_mask, d_s1, feat_b, (size_t)F);
        c->total_launches += 2; c->timing.launches += 2;
        ScanArgs a{};
        a.feat = feat; a.fstride = wprS; a.samp = samp; a.sstride = wprF; a.pos = d_pos; a.wpos = d_wpos; a.s1 = d_s1;
        a.T = (const double *)p_T; a.Fp = Fp; a.F = F; a.S = S; a.wprS = wprS; a.k = kp[p];
        // few features: 128-thread CTAs and ONE l-chunk of 128 * nj >= F (nj = 1 .. 4 values of l per thread): a chunk wider than F
        // would spend a slice of every h on l >= F (F = 384 with 512-wide chunks: a third of the issued slices)
        const bool narrow = F <= 512;
        const int nj = narrow ? std::max(1, ceil_div(F, 128)) : SCAN_NJ;
        a.g_lo = 0; a.g_hi = F; a.gt0 = 0; a.n_gt = ceil_div(F, SCAN_GT); a.n_hb = ceil_div(F, SCAN_HB);
        a.n_lc = ceil_div(F, narrow ? 128 * nj : SCAN_NT * SCAN_NJ);
        a.nb = nb; a.feat_b = feat_b; a.samp_b = samp_b; a.s1_b = (size_t)F; a.T_b = (size_t)F * Fp;
        a.best = d_best; a.item_counter = d_item;
        if ((int64_t)a.n_gt * a.n_hb * a.n_lc * nb >= (1ll << 31)) return sdx_fail(c, SDX_ERR_UNSUPPORTED, "too many scan work items");
        if (kp[p] >= 2 && F >= 2) {
            k_pair_table<<<dim3(std::min(F - 1, c->sm_count * 4), nb), 256, sizeof(uint32_t) * 2 * wprS, c->stream>>>(
                feat, wprS, F, wprS, d_pos, d_wpos, (double *)p_T, Fp, feat_b, (size_t)F * Fp);
            c->total_launches++; c->timing.launches++;
        }
        k_scan12_max<<<dim3(std::max(1, std::min(F, c->sm_count * 2 / std::max(1, nb) + 1)), nb), SCAN_NT, 0, c->stream>>>(a);
        c->total_launches++; c->timing.launches++;
        if (kp[p] >= 3 && F >= 3) {
            if (narrow && nj == 1) k_scan3<false, 128, 1><<<c->sm_count * 5, 128, 0, c->stream>>>(a);
            else if (narrow && nj == 2) k_scan3<false, 128, 2><<<c->sm_count * 5, 128, 0, c->stream>>>(a);
            else if (narrow && nj == 3) k_scan3<false, 128, 3><<<c->sm_count * 5, 128, 0, c->stream>>>(a);
            else if (narrow) k_scan3<false, 128, 4><<<c->sm_count * 5, 128, 0, c->stream>>>(a);
            else k_scan3<false, SCAN_NT, SCAN_NJ><<<grid3, SCAN_NT, 0, c->stream>>>(a);
            c->total_launches++; c->timing.launches++;
        }
        SDX_CUDA(c, cudaGetLastError());
        return SDX_OK;
    };

    memset(&c->timing, 0, sizeof c->timing);
    cudaEventRecord(c->ev[5], c->stream);
    // observed statistic
    SDX_CUDA(c, cudaMemsetAsync(d_ge, 0, sizeof(unsigned long long) * P, c->stream));
    if (z_obs) {
        SDX_CUDA(c, cudaMemcpyAsync(d_z, z_obs, sizeof(double) * P, cudaMemcpyHostToDevice, c->stream));
    } else {
        for (int p = 0; p < P; ++p) {
            if ((rc = scan_batch(p, c->d_feat, 0, c->d_samp, 0, 1))) return rc;
            k_maxk_finalize<<<1, 32, 0, c->stream>>>(d_best, 1, nullptr, d_ge, d_z + p);
            c->total_launches++; c->timing.launches++;
        }
    }
    float gen_ms = 0, plan_ms = 0;
    int launches = c->timing.launches;
    for (int64_t b0 = 0; b0 < n_nulls; b0 += B) {
        const int nb = (int)std::min<int64_t>(B, n_nulls - b0);
        if (rows_src) {
            SDX_CUDA(c, cudaMemcpyAsync(d_featB, rows_src + (size_t)b0 * feat_words, sizeof(uint32_t) * feat_words * nb, cudaMemcpyDefault, c->stream));
            rc = sdx_transpose_batch(c, d_featB, F, wprS, S, d_sampB, wprF, nb);                  // feature rows -> sample rows
        } else {
            if ((rc = sdx_generate_to_device(c, seed, null0 + (uint64_t)b0, nb, n_trades, chain_len, d_a))) return rc;
            gen_ms += c->timing.trade_ms; plan_ms += c->timing.plan_ms; launches += c->timing.launches;
            c->timing.launches = 0;
            if (c->rows_are_samples) rc = sdx_transpose_batch(c, d_a, S, wprF, F, d_b, wprS, nb);     // sample rows -> feature rows
            else rc = sdx_transpose_batch(c, d_a, F, wprS, S, d_b, wprF, nb);                         // feature rows -> sample rows
        }
        if (rc) return rc;
        ++launches;
        for (int p = 0; p < P; ++p) {
            if ((rc = scan_batch(p, d_featB, feat_words, d_sampB, samp_words, nb))) return rc;
            k_maxk_finalize<<<ceil_div(nb, 128), 128, 0, c->stream>>>(d_best, nb, d_z + p, d_ge + p,
                                                                     d_scores ? d_scores + (size_t)p * n_nulls + b0 : nullptr);
            c->total_launches++; c->timing.launches++;
        }
        launches += c->timing.launches;
        c->timing.launches = 0;
    }
    cudaEventRecord(c->ev[1], c->stream);
    std::vector<unsigned long long> ge((size_t)P);
    SDX_CUDA(c, cudaMemcpyAsync(ge.data(), d_ge, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
    if (z_obs_out) SDX_CUDA(c, cudaMemcpyAsync(z_obs_out, d_z, sizeof(double) * P, cudaMemcpyDeviceToHost, c->stream));
    if (null_scores && n_nulls > 0)
        SDX_CUDA(c, cudaMemcpyAsync(null_scores, d_scores, sizeof(double) * (size_t)P * n_nulls, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    SDX_CUDA(c, cudaGetLastError());
    for (int p = 0; p < P; ++p) n_ge[p] = (int64_t)ge[p];
    cudaEventElapsedTime(&c->timing.total_ms, c->ev[5], c->ev[1]);
    c->timing.trade_ms = gen_ms; c->timing.plan_ms = plan_ms; c->timing.launches = launches;
    c->timing.score_ms = c->timing.total_ms - gen_ms - plan_ms;
    return SDX_OK;
}

int sdx_null_pvalue_maxk(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                         const int32_t *modules, int k, const double *z_obs, int64_t *n_ge, double *null_scores,
                         double *z_obs_out) {
    return maxk_impl(c, nullptr, seed, null0, n_nulls, n_trades, chain_len, modules, k, z_obs, n_ge, null_scores, z_obs_out);
}

// ---------------------------------------------------------------------------
// sdx_null_pvalue_rows: the null loop of src/superdendrix.py:175-192 on nulls the CALLER supplies (the reference reads
// them from <nm><i>.tsv; here they are packed feature rows, e.g. from sdx_read_nulls_tsv): every null scored for every
// profile in one call.  stat = fixed: one warp per (null, profile) runs score_module on the null's rows.
// ---------------------------------------------------------------------------
__global__ void k_score_null_rows(const uint32_t *__restrict__ feat, size_t feat_b, int wprS, int S, int P, int k, int64_t nb,
                                  const int32_t *__restrict__ modules, const double *__restrict__ w, const uint32_t *__restrict__ mask,
                                  const double *__restrict__ z_obs, unsigned long long *__restrict__ n_ge, double *__restrict__ scores,
                                  int64_t n_total, int64_t b0) {
    const int lane = threadIdx.x & 31;
    const int64_t job = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (job >= nb * P) return;
    const int64_t b = job / P;
    const int p = (int)(job - b * P);
    const SetScore sc = score_module(feat + (size_t)b * feat_b, wprS, false, modules + (size_t)p * k, k, w + (size_t)p * S, mask + (size_t)p * wprS, S, lane);
    if (lane == 0) {
        if (scores) scores[(size_t)p * n_total + b0 + b] = sc.W;
        if (sc.W >= z_obs[p]) atomicAdd(&n_ge[p], 1ull);
    }
}

__global__ void k_score_observed_rows(const uint32_t *__restrict__ feat, int wprS, int S, int P, int k, const int32_t *__restrict__ modules,
                                      const double *__restrict__ w, const uint32_t *__restrict__ mask, double *__restrict__ z) {
    const int lane = threadIdx.x & 31, p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= P) return;
    const SetScore sc = score_module(feat, wprS, false, modules + (size_t)p * k, k, w + (size_t)p * S, mask + (size_t)p * wprS, S, lane);
    if (lane == 0) z[p] = sc.W;
}

extern "C" int sdx_null_pvalue_rows(sdx_ctx *c, const uint32_t *feature_rows, int64_t n_nulls, int stat, const int32_t *modules, int k,
                                    const double *z_obs, int64_t *n_ge, double *null_scores, double *z_obs_out) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, modules && n_ge && k >= 1 && k <= SDX_MAX_K && n_nulls >= 0, SDX_ERR_INVALID, "modules / n_ge / k / n_nulls invalid");
    SDX_REQUIRE(c, feature_rows || n_nulls == 0, SDX_ERR_INVALID, "feature_rows is NULL");
    SDX_REQUIRE(c, stat == SDX_STAT_FIXED || stat == SDX_STAT_MAXK, SDX_ERR_INVALID, "unknown statistic");
    const int F = c->F, S = c->S, wprS = c->wprS, P = c->P;
    for (int i = 0; i < P * k; ++i) SDX_REQUIRE(c, modules[i] >= -1 && modules[i] < F, SDX_ERR_INVALID, "module feature index out of range");
    if (stat == SDX_STAT_MAXK) return maxk_impl(c, feature_rows, 0, 0, n_nulls, 0, 1, modules, k, z_obs, n_ge, null_scores, z_obs_out);
    LaunchTimer timer(c);
    const size_t feat_words = (size_t)F * wprS;
    int64_t B = (int64_t)(((size_t)1 << 30) / (sizeof(uint32_t) * feat_words + 1));      // ~1 GB of staging per batch
    if (B < 1) B = 1;
    if (B > n_nulls && n_nulls > 0) B = n_nulls;
    cudaPointerAttributes at{};
    const bool on_device = n_nulls > 0 && cudaPointerGetAttributes(&at, feature_rows) == cudaSuccess && at.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    void *p_small, *p_rows = nullptr, *p_sc = nullptr;
    const size_t off_z = ((sizeof(int32_t) * (size_t)P * k + 15) / 16) * 16, off_ge = off_z + sizeof(double) * P;
    SDX_TRY(sdx_scratch(c, 14, off_ge + sizeof(unsigned long long) * P + 64, &p_small));
    if (!on_device && n_nulls > 0) SDX_TRY(sdx_scratch(c, 11, sizeof(uint32_t) * feat_words * B + 64, &p_rows));
    if (null_scores && n_nulls > 0) SDX_TRY(sdx_scratch(c, 15, sizeof(double) * (size_t)P * n_nulls + 64, &p_sc));
    int32_t *d_mod = (int32_t *)p_small;
    double *d_z = (double *)((char *)p_small + off_z);
    unsigned long long *d_ge = (unsigned long long *)((char *)p_small + off_ge);
    SDX_CUDA(c, cudaMemcpyAsync(d_mod, modules, sizeof(int32_t) * (size_t)P * k, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaMemsetAsync(d_ge, 0, sizeof(unsigned long long) * P, c->stream));
    if (z_obs) SDX_CUDA(c, cudaMemcpyAsync(d_z, z_obs, sizeof(double) * P, cudaMemcpyHostToDevice, c->stream));
    else { k_score_observed_rows<<<ceil_div(P, 4), 128, 0, c->stream>>>(c->d_feat, wprS, S, P, k, d_mod, c->d_w, c->d_mask, d_z); timer.count(); }
    for (int64_t b0 = 0; b0 < n_nulls; b0 += B) {
        const int64_t nb = std::min<int64_t>(B, n_nulls - b0);
        const uint32_t *d_rows = feature_rows + (size_t)b0 * feat_words;
        if (!on_device) {
            SDX_CUDA(c, cudaMemcpyAsync(p_rows, d_rows, sizeof(uint32_t) * feat_words * nb, cudaMemcpyHostToDevice, c->stream));
            d_rows = (const uint32_t *)p_rows;
        }
        k_score_null_rows<<<(unsigned)ceil_div(nb * P, 8), 256, 0, c->stream>>>(d_rows, feat_words, wprS, S, P, k, nb, d_mod, c->d_w, c->d_mask, d_z, d_ge,
                                                                                (double *)p_sc, n_nulls, b0);
        timer.count();
        SDX_CUDA(c, cudaGetLastError());
        if (!on_device) SDX_CUDA(c, cudaStreamSynchronize(c->stream));      // the staging buffer is reused by the next batch
    }
    std::vector<unsigned long long> ge((size_t)P);
    SDX_CUDA(c, cudaMemcpyAsync(ge.data(), d_ge, sizeof(unsigned long long) * P, cudaMemcpyDeviceToHost, c->stream));
    if (z_obs_out) SDX_CUDA(c, cudaMemcpyAsync(z_obs_out, d_z, sizeof(double) * P, cudaMemcpyDeviceToHost, c->stream));
    if (null_scores && n_nulls > 0)
        SDX_CUDA(c, cudaMemcpyAsync(null_scores, p_sc, sizeof(double) * (size_t)P * n_nulls, cudaMemcpyDeviceToHost, c->stream));
    int rc = timer.finish();
    for (int p = 0; p < P; ++p) n_ge[p] = (int64_t)ge[p];
    c->timing.score_ms = c->timing.total_ms;
    return rc;
}

// ---------------------------------------------------------------------------
// sdx_curveball_features: sdx_curveball with the nulls returned as FEATURE rows over samples whatever the trade
// orientation — the layout the driver's file format wants (one line per sample, src/generate_null_matrices.py:75-86).  When
// the trade rows are samples (F >= S) the batch is transposed on the device.
// ---------------------------------------------------------------------------
extern "C" int sdx_curveball_features(sdx_ctx *c, uint64_t seed, uint64_t null0, int64_t n_nulls, int64_t n_trades, int64_t chain_len,
                                      uint32_t *out_feature_rows) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_trade_obs, SDX_ERR_STATE, "sdx_upload_matrix has not been called");
    SDX_REQUIRE(c, out_feature_rows || n_nulls == 0, SDX_ERR_INVALID, "out_feature_rows is NULL");
    if (!c->rows_are_samples) return sdx_curveball(c, seed, null0, n_nulls, n_trades, chain_len, out_feature_rows, nullptr);
    const int F = c->F, S = c->S, wprS = c->wprS, wprF = c->wprF;
    const size_t feat_words = (size_t)F * wprS, samp_words = (size_t)S * wprF;
    int64_t B = (int64_t)(((size_t)1 << 30) / (sizeof(uint32_t) * (feat_words + samp_words) + 1));
    if (B < 1) B = 1;
    if (B > n_nulls) B = n_nulls;
    void *p_mat = nullptr;
    if (n_nulls > 0) SDX_TRY(sdx_scratch(c, 11, sizeof(uint32_t) * (feat_words + samp_words) * B + 64, &p_mat));
    uint32_t *d_samp = (uint32_t *)p_mat, *d_feat = d_samp + samp_words * B;
    float trade_ms = 0, plan_ms = 0;
    int launches = 0, tl = 0;
    for (int64_t b0 = 0; b0 < n_nulls; b0 += B) {
        const int nb = (int)std::min<int64_t>(B, n_nulls - b0);
        SDX_TRY(sdx_generate_to_device(c, seed, null0 + (uint64_t)b0, nb, n_trades, chain_len, d_samp));
        trade_ms += c->timing.trade_ms; plan_ms += c->timing.plan_ms; launches += c->timing.launches; tl += c->timing.trade_launches;
        SDX_CUDA(c, cudaMemsetAsync(d_feat, 0, sizeof(uint32_t) * feat_words * nb, c->stream));
        SDX_TRY(sdx_transpose_batch(c, d_samp, S, wprF, F, d_feat, wprS, nb));
        SDX_CUDA(c, cudaMemcpyAsync(out_feature_rows + (size_t)b0 * feat_words, d_feat, sizeof(uint32_t) * feat_words * nb, cudaMemcpyDefault, c->stream));
        SDX_CUDA(c, cudaStreamSynchronize(c->stream));
        ++launches;
    }
    c->timing.trade_ms = trade_ms; c->timing.plan_ms = plan_ms; c->timing.launches = launches; c->timing.trade_launches = tl;
    return SDX_OK;
}
