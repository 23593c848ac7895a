// sdx_score.cu — stand-alone set scorer, per-feature scores, self-test hooks.
//   src/superdendrix.py:387-391 (SuperW), :281-286, :219-236 (per-feature outputs)
#include "sdx_trade.cuh"

__global__ void k_score_sets(const uint32_t *__restrict__ feat, int wprS, int S, const int32_t *__restrict__ sets,
                             const int32_t *__restrict__ prof, int64_t n_sets, int k, const double *__restrict__ w,
                             const uint32_t *__restrict__ mask, double *W, int32_t *cov, int32_t *ovl, int32_t *act) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_sets) return;
    const int p = prof ? prof[i] : 0;
    SetScore s = score_module(feat, wprS, false, sets + i * k, k, w + (size_t)p * S, mask + (size_t)p * wprS, S, lane);
    if (lane == 0) {
        W[i] = s.W; cov[i] = s.cov; ovl[i] = s.tot - s.cov; act[i] = s.act;
    }
}

// device-pointer variant used inside the library (scan re-scoring): no copies, no sync
int sdx_score_sets_device(sdx_ctx *c, const int32_t *d_sets, const int32_t *d_prof, int profile0, int64_t n, int k,
                          double *d_W, int32_t *d_cov, int32_t *d_ovl, int32_t *d_act) {
    if (n <= 0) return SDX_OK;
    k_score_sets<<<ceil_div(n, 8), 256, 0, c->stream>>>(c->d_feat, c->wprS, c->S, d_sets, d_prof, n, k,
                                                        c->d_w + (size_t)profile0 * c->S, c->d_mask + (size_t)profile0 * c->wprS,
                                                        d_W, d_cov, d_ovl, d_act);
    c->total_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sdx_fail(c, SDX_ERR_CUDA, std::string("k_score_sets launch failed: ") + cudaGetErrorString(e));
    return SDX_OK;
}

extern "C" int sdx_score_sets(sdx_ctx *c, const int32_t *sets, const int32_t *prof, int64_t n, int k,
                              double *W, int32_t *cov, int32_t *ovl, int32_t *act) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, n >= 0 && k >= 1 && k <= SDX_MAX_K && (n == 0 || sets), SDX_ERR_INVALID, "bad sets / k");
    for (int64_t i = 0; i < n * k; ++i)
        SDX_REQUIRE(c, sets[i] >= -1 && sets[i] < c->F, SDX_ERR_INVALID, "feature index out of range");
    if (prof) for (int64_t i = 0; i < n; ++i)
        SDX_REQUIRE(c, prof[i] >= 0 && prof[i] < c->P, SDX_ERR_INVALID, "profile index out of range");
    LaunchTimer timer(c);
    if (n == 0) return timer.finish();
    void *p_in, *p_out;
    int rc;
    const size_t in_bytes = sizeof(int32_t) * n * (k + 1);
    const size_t out_bytes = (sizeof(double) + 3 * sizeof(int32_t)) * n;
    if ((rc = sdx_scratch(c, 0, in_bytes + 64, &p_in))) return rc;
    if ((rc = sdx_scratch(c, 1, out_bytes + 64, &p_out))) return rc;
    int32_t *d_sets = (int32_t *)p_in, *d_prof = prof ? d_sets + n * k : nullptr;
    double *d_W = (double *)p_out;
    int32_t *d_cov = (int32_t *)(d_W + n), *d_ovl = d_cov + n, *d_act = d_ovl + n;
    SDX_CUDA(c, cudaMemcpyAsync(d_sets, sets, sizeof(int32_t) * n * k, cudaMemcpyHostToDevice, c->stream));
    if (prof) SDX_CUDA(c, cudaMemcpyAsync(d_prof, prof, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    cudaEventRecord(c->ev[2], c->stream);
    k_score_sets<<<ceil_div(n, 8), 256, 0, c->stream>>>(c->d_feat, c->wprS, c->S, d_sets, d_prof, n, k, c->d_w, c->d_mask,
                                                        d_W, d_cov, d_ovl, d_act);
    cudaEventRecord(c->ev[3], c->stream);
    timer.count();
    if (W) SDX_CUDA(c, cudaMemcpyAsync(W, d_W, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    if (cov) SDX_CUDA(c, cudaMemcpyAsync(cov, d_cov, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (ovl) SDX_CUDA(c, cudaMemcpyAsync(ovl, d_ovl, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    if (act) SDX_CUDA(c, cudaMemcpyAsync(act, d_act, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c->stream));
    rc = timer.finish();
    cudaEventElapsedTime(&c->timing.score_ms, c->ev[2], c->ev[3]);
    return rc;
}

// sum of w over the union of each set (the "total_score" of utils/compute_individual_contribution_SELECT.py:75)
__global__ void k_union_sums(const uint32_t *__restrict__ feat, int wprS, int S, const int32_t *__restrict__ sets,
                             const int32_t *__restrict__ prof, int64_t n_sets, int k, const double *__restrict__ w,
                             const uint32_t *__restrict__ mask, double *sum_w) {
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_sets) return;
    const int p = prof ? prof[i] : 0;
    SetScore s = score_module(feat, wprS, false, sets + i * k, k, w + (size_t)p * S, mask + (size_t)p * wprS, S, lane);
    if (lane == 0) sum_w[i] = s.sw;
}

extern "C" int sdx_union_sums(sdx_ctx *c, const int32_t *sets, const int32_t *prof, int64_t n, int k, double *sum_w) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, n >= 0 && k >= 1 && k <= SDX_MAX_K && (n == 0 || (sets && sum_w)), SDX_ERR_INVALID, "bad sets / k");
    for (int64_t i = 0; i < n * k; ++i)
        SDX_REQUIRE(c, sets[i] >= -1 && sets[i] < c->F, SDX_ERR_INVALID, "feature index out of range");
    if (prof) for (int64_t i = 0; i < n; ++i)
        SDX_REQUIRE(c, prof[i] >= 0 && prof[i] < c->P, SDX_ERR_INVALID, "profile index out of range");
    LaunchTimer timer(c);
    if (n == 0) return timer.finish();
    void *p_in, *p_out;
    int rc;
    if ((rc = sdx_scratch(c, 0, sizeof(int32_t) * n * (k + 1) + 64, &p_in))) return rc;
    if ((rc = sdx_scratch(c, 1, sizeof(double) * n + 64, &p_out))) return rc;
    int32_t *d_sets = (int32_t *)p_in, *d_prof = prof ? d_sets + n * k : nullptr;
    SDX_CUDA(c, cudaMemcpyAsync(d_sets, sets, sizeof(int32_t) * n * k, cudaMemcpyHostToDevice, c->stream));
    if (prof) SDX_CUDA(c, cudaMemcpyAsync(d_prof, prof, sizeof(int32_t) * n, cudaMemcpyHostToDevice, c->stream));
    k_union_sums<<<ceil_div(n, 8), 256, 0, c->stream>>>(c->d_feat, c->wprS, c->S, d_sets, d_prof, n, k, c->d_w, c->d_mask, (double *)p_out);
    timer.count();
    SDX_CUDA(c, cudaMemcpyAsync(sum_w, p_out, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    return timer.finish();
}

// per-feature: one warp per feature
__global__ void k_feature_scores(const uint32_t *__restrict__ feat, int F, int wprS, int S, const double *__restrict__ w,
                                 const uint32_t *__restrict__ mask, double *sum_w, double *W1, int32_t *count) {
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= F) return;
    double sw = 0, pos = 0, neg = 0;
    int cnt = 0;
    for (int wi = lane; wi < wprS; wi += 32) {
        uint32_t x = feat[(size_t)g * wprS + wi] & mask[wi];
        cnt += __popc(x);
        while (x) {
            int bp = __ffs(x) - 1;
            x &= x - 1;
            double ws = w[wi * 32 + bp];
            sw += ws;
            if (ws > 0) pos += ws;
            neg += fabs(ws);
        }
    }
    sw = warp_sum_f64(sw); pos = warp_sum_f64(pos); neg = warp_sum_f64(neg);
    cnt = warp_sum_i32(cnt);
    if (lane == 0) { sum_w[g] = sw; W1[g] = 2.0 * pos - neg; count[g] = cnt; }
}

// argmax with "first maximum wins" (strict > in src/superdendrix.py:284)
__global__ void k_argmax_first(const double *v, int n, int32_t *out) {
    __shared__ double sv[256];
    __shared__ int si[256];
    double best = -INFINITY; int bi = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (v[i] > best) { best = v[i]; bi = i; }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            double ov = sv[threadIdx.x + o]; int oi = si[threadIdx.x + o];
            bool take = oi >= 0 && (si[threadIdx.x] < 0 || ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x]));
            if (take) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = si[0];
}

extern "C" int sdx_feature_scores(sdx_ctx *c, int profile, double *sum_w, double *W1, int32_t *count, int32_t *best) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, profile >= 0 && profile < c->P, SDX_ERR_INVALID, "profile index out of range");
    LaunchTimer timer(c);
    const int F = c->F;
    void *p;
    int rc;
    if ((rc = sdx_scratch(c, 1, (2 * sizeof(double) + sizeof(int32_t)) * F + 64, &p))) return rc;
    double *d_sw = (double *)p, *d_W1 = d_sw + F;
    int32_t *d_cnt = (int32_t *)(d_W1 + F), *d_best = d_cnt + F;
    cudaEventRecord(c->ev[2], c->stream);
    k_feature_scores<<<ceil_div(F, 8), 256, 0, c->stream>>>(c->d_feat, F, c->wprS, c->S, c->d_w + (size_t)profile * c->S,
                                                            c->d_mask + (size_t)profile * c->wprS, d_sw, d_W1, d_cnt);
    k_argmax_first<<<1, 256, 0, c->stream>>>(d_sw, F, d_best);
    cudaEventRecord(c->ev[3], c->stream);
    timer.count(2);
    if (sum_w) SDX_CUDA(c, cudaMemcpyAsync(sum_w, d_sw, sizeof(double) * F, cudaMemcpyDeviceToHost, c->stream));
    if (W1) SDX_CUDA(c, cudaMemcpyAsync(W1, d_W1, sizeof(double) * F, cudaMemcpyDeviceToHost, c->stream));
    if (count) SDX_CUDA(c, cudaMemcpyAsync(count, d_cnt, sizeof(int32_t) * F, cudaMemcpyDeviceToHost, c->stream));
    if (best) SDX_CUDA(c, cudaMemcpyAsync(best, d_best, sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
    rc = timer.finish();
    cudaEventElapsedTime(&c->timing.score_ms, c->ev[2], c->ev[3]);
    return rc;
}

// ---------------------------------------------------------------------------
// self-test hooks: known-answer checks of the device random streams
// ---------------------------------------------------------------------------
__global__ void k_test_philox(const uint32_t *ctr, const uint32_t *key, int n, uint32_t *out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PhiloxKeys k;
    uint32_t a = key[2 * i], b = key[2 * i + 1];
    for (int r = 0; r < 10; ++r) { k.k0[r] = a; k.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    uint4 c = make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]);
    uint4 o = philox4x32_10(c, k);
    out[4 * i] = o.x; out[4 * i + 1] = o.y; out[4 * i + 2] = o.z; out[4 * i + 3] = o.w;
}

extern "C" int sdx_test_philox(sdx_ctx *c, const uint32_t *ctr, const uint32_t *key, int n, uint32_t *out) {
    if (!c || !ctr || !key || !out || n < 1) return SDX_ERR_INVALID;
    LaunchTimer timer(c);
    void *p;
    int rc;
    if ((rc = sdx_scratch(c, 0, sizeof(uint32_t) * 10 * (size_t)n, &p))) return rc;
    uint32_t *d_ctr = (uint32_t *)p, *d_key = d_ctr + 4 * n, *d_out = d_key + 2 * n;
    SDX_CUDA(c, cudaMemcpyAsync(d_ctr, ctr, sizeof(uint32_t) * 4 * n, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(d_key, key, sizeof(uint32_t) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    k_test_philox<<<ceil_div(n, 128), 128, 0, c->stream>>>(d_ctr, d_key, n, d_out);
    timer.count();
    SDX_CUDA(c, cudaMemcpyAsync(out, d_out, sizeof(uint32_t) * 4 * n, cudaMemcpyDeviceToHost, c->stream));
    return timer.finish();
}

__global__ void k_test_pairs(const __grid_constant__ PhiloxKeys keys, uint64_t null_id, int T, int R, int32_t *a, int32_t *b) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    uint32_t x, y;
    trade_pair(keys, null_id, (uint32_t)t, (uint32_t)R, x, y);
    a[t] = (int32_t)x; b[t] = (int32_t)y;
}

extern "C" int sdx_test_trade_pairs(sdx_ctx *c, uint64_t seed, uint64_t null_id, int T, int R, int32_t *a, int32_t *b) {
    if (!c || !a || !b || T < 1 || R < 2) return SDX_ERR_INVALID;
    LaunchTimer timer(c);
    void *p;
    int rc;
    if ((rc = sdx_scratch(c, 0, sizeof(int32_t) * 2 * (size_t)T, &p))) return rc;
    int32_t *d_a = (int32_t *)p, *d_b = d_a + T;
    k_test_pairs<<<ceil_div(T, 128), 128, 0, c->stream>>>(make_keys(seed), null_id, T, R, d_a, d_b);
    timer.count();
    SDX_CUDA(c, cudaMemcpyAsync(a, d_a, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(b, d_b, sizeof(int32_t) * T, cudaMemcpyDeviceToHost, c->stream));
    return timer.finish();
}

// ---------------------------------------------------------------------------
// Integer-pipe peaks measured on the box (bench.py): register-only loops of eight independent chains
// per thread, enough CTAs to fill every SM; the rates are what the XU (POPC) and ALU (LOP3) pipes sustain.
// ---------------------------------------------------------------------------
#define PEAK_CHAINS 8
template <bool POPC>
__global__ void __launch_bounds__(256) k_int_peak(uint32_t *out, int iters, uint32_t salt) {
    uint32_t a[PEAK_CHAINS], s[PEAK_CHAINS];
#pragma unroll
    for (int j = 0; j < PEAK_CHAINS; ++j) { a[j] = (threadIdx.x + 1u) * 2654435761u + j * salt; s[j] = j + salt; }
    const uint32_t m0 = salt | 0x55555555u, m1 = ~salt ^ 0x0f0f0f0fu;
#pragma unroll 8
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < PEAK_CHAINS; ++j) {
            if (POPC) s[j] += __popc(a[j] ^ s[j]);               // one LOP3 + one POPC + one IADD per step: POPC is the slow pipe
            else asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(a[j]) : "r"(m0), "r"(m1));      // one LOP3 (majority) per step
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < PEAK_CHAINS; ++j) r ^= a[j] + s[j];
    if (r == 0x12345678u) out[0] = r;                            // keeps the loops alive without a store per thread
}

extern "C" int sdx_measure_int_peaks(sdx_ctx *c, double *popc_per_s, double *lop3_per_s) {
    if (!c || !popc_per_s || !lop3_per_s) return SDX_ERR_INVALID;
    SDX_CUDA(c, cudaSetDevice(c->device));
    void *p;
    int rc;
    if ((rc = sdx_scratch(c, 0, 64, &p))) return rc;
    const int grid = c->sm_count * 8, threads = 256, iters = 4096;
    float ms[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 3; ++rep) {                       // two warm-up launches, the third is timed
            cudaEventRecord(c->ev[2], c->stream);
            if (which == 0) k_int_peak<true><<<grid, threads, 0, c->stream>>>((uint32_t *)p, iters, 0x9e3779b9u + rep);
            else k_int_peak<false><<<grid, threads, 0, c->stream>>>((uint32_t *)p, iters, 0x9e3779b9u + rep);
            cudaEventRecord(c->ev[3], c->stream);
            c->total_launches++;
        }
        SDX_CUDA(c, cudaStreamSynchronize(c->stream));
        SDX_CUDA(c, cudaGetLastError());
        cudaEventElapsedTime(&ms[which], c->ev[2], c->ev[3]);
    }
    const double ops = (double)grid * threads * iters * PEAK_CHAINS;
    *popc_per_s = ops / (ms[0] * 1e-3);
    *lop3_per_s = ops / (ms[1] * 1e-3);
    return SDX_OK;
}
