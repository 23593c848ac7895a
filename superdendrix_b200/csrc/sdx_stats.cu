// sdx_stats.cu — conditional permutation test and rank-sum.
//   src/superdendrix.py:361-384   conditional_permutation_test
//   utils/compute_p_value.py:306-318, utils/compute_ranksum_pval.py:116-128   scipy.stats.ranksums call sites
#include <algorithm>

#include "sdx_trade.cuh"

// ---------------------------------------------------------------------------
// Conditional permutation.  Per (job, slot i): with U = union of the other
// features of the module and base = W(M \ g_i),
//     W(M with g_i := R) = base + sum_{s in R} v_s,   v_s = 2 w_s [w_s > 0, s not in U] - |w_s|
// (SURVEY.md §8 a7), so one permutation = a sum of n_i draws without
// replacement from v.  The set-up kernel walks the samples in ascending order
// with one thread per (job, slot) so base, v and real_W are bit-identical to
// the sequential restatement (oracle/sdx_oracle.c: sdxo_condperm); the draws
// are Floyd's algorithm on stream (PERM, perm_id0 + slot, t).
// ---------------------------------------------------------------------------
struct PermSlot {
    double base, real_W;
    int n_i, n_s;
};

__global__ void k_condperm_setup(const uint32_t *__restrict__ feat, int wprS, int S, const double *__restrict__ w_all,
                                 const uint32_t *__restrict__ mask_all, const int32_t *__restrict__ prof,
                                 const int32_t *__restrict__ modules, int n_jobs, int k,
                                 double *__restrict__ v_all, PermSlot *__restrict__ slots) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_jobs * k) return;
    const int job = idx / k, i = idx - job * k;
    const int32_t *mod = modules + (size_t)job * k;
    const int p = prof ? prof[job] : 0;
    const double *w = w_all + (size_t)p * S;
    const uint32_t *mask = mask_all + (size_t)p * wprS;
    double *v = v_all + (size_t)idx * S;
    PermSlot out;
    out.n_i = 0; out.n_s = 0; out.base = 0; out.real_W = 0;
    if (mod[i] < 0) { out.n_i = -1; slots[idx] = out; return; }      // empty slot of a padded module
    double pos = 0.0, neg = 0.0, real_sum = 0.0;
    int q = 0;
    for (int s = 0; s < S; ++s) {
        const int wi = s >> 5;
        const uint32_t bit = 1u << (s & 31);
        if (!(mask[wi] & bit)) continue;
        int mult = 0;
        for (int j = 0; j < k; ++j)
            if (j != i && mod[j] >= 0 && (feat[(size_t)mod[j] * wprS + wi] & bit)) ++mult;
        const double ws = w[s];
        if (mult) {
            if (ws > 0) pos += ws;
            neg += mult * fabs(ws);
        }
        const double vs = ((ws > 0 && !mult) ? 2.0 * ws : 0.0) - fabs(ws);
        v[q] = vs;
        if (feat[(size_t)mod[i] * wprS + wi] & bit) { ++out.n_i; real_sum += vs; }
        ++q;
    }
    out.n_s = q;
    out.base = 2.0 * pos - neg;
    out.real_W = out.base + real_sum;
    slots[idx] = out;
}

__global__ void k_condperm(const __grid_constant__ PhiloxKeys keys, const double *__restrict__ v_all, int S,
                           const PermSlot *__restrict__ slots, int n_slots, int k, uint64_t perm_id0, int64_t n_perm,
                           int words, unsigned long long *__restrict__ counts) {
    extern __shared__ uint32_t chosen[];          // [words][blockDim.x]
    const int slot = blockIdx.y;
    const PermSlot sl = slots[slot];
    if (sl.n_i < 0) return;
    const double *v = v_all + (size_t)slot * S;
    const int nt = blockDim.x, tid = threadIdx.x;
    const uint64_t stream_id = perm_id0 + (uint64_t)(slot % k);
    unsigned int hits = 0;
    for (int64_t t = (int64_t)blockIdx.x * nt + tid; t < n_perm; t += (int64_t)gridDim.x * nt) {
        for (int wi = 0; wi < words; ++wi) chosen[wi * nt + tid] = 0u;
        ThreadStream st(keys, SDX_STREAM_PERM, stream_id, (uint32_t)t);
        double sum = 0.0;
        for (int j = sl.n_s - sl.n_i; j < sl.n_s; ++j) {
            uint32_t x = st.below((uint32_t)j + 1u);
            uint32_t cw = chosen[(x >> 5) * nt + tid];
            if ((cw >> (x & 31)) & 1u) { x = (uint32_t)j; cw = chosen[(x >> 5) * nt + tid]; }
            chosen[(x >> 5) * nt + tid] = cw | (1u << (x & 31));
            sum += v[x];
        }
        hits += (sl.base + sum >= sl.real_W) ? 1u : 0u;
    }
    hits = __reduce_add_sync(0xffffffffu, hits);
    if ((tid & 31) == 0 && hits) atomicAdd(&counts[slot], (unsigned long long)hits);
}

static int condperm_run(sdx_ctx *c, const int32_t *prof, const int32_t *modules, int n_jobs, int k, uint64_t seed,
                        uint64_t perm_id0, int64_t n_perm, int64_t *counts, double *real_W) {
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, n_jobs >= 1 && modules && counts, SDX_ERR_INVALID, "modules / counts missing");
    SDX_REQUIRE(c, k >= 1 && k <= SDX_MAX_K, SDX_ERR_INVALID, "module size k must be in 1..SDX_MAX_K");
    SDX_REQUIRE(c, n_perm >= 0 && n_perm < (1ll << 32), SDX_ERR_INVALID, "n_perm out of range");
    SDX_REQUIRE(c, perm_id0 < (1ull << 61), SDX_ERR_INVALID, "perm_id0 out of range");
    for (int64_t i = 0; i < (int64_t)n_jobs * k; ++i)
        SDX_REQUIRE(c, modules[i] >= -1 && modules[i] < c->F, SDX_ERR_INVALID, "module feature index out of range");
    if (prof) for (int j = 0; j < n_jobs; ++j)
        SDX_REQUIRE(c, prof[j] >= 0 && prof[j] < c->P, SDX_ERR_INVALID, "profile index out of range");
    const int S = c->S, n_slots = n_jobs * k;
    SDX_REQUIRE(c, n_slots <= 65535, SDX_ERR_UNSUPPORTED, "too many (job, slot) pairs in one call (max 65535)");
    const int words = c->wprS;
    int threads = 256;
    while (threads > 32 && (size_t)threads * words * 4 > (size_t)160 * 1024) threads >>= 1;
    const size_t smem = (size_t)threads * words * 4;
    SDX_REQUIRE(c, smem <= c->smem_optin, SDX_ERR_UNSUPPORTED, "too many samples for the permutation kernel");
    LaunchTimer timer(c);
    int rc;
    void *p_in, *p_v, *p_out;
    const size_t in_bytes = sizeof(int32_t) * ((size_t)n_slots + n_jobs) + 64;
    if ((rc = sdx_scratch(c, 0, in_bytes, &p_in))) return rc;
    if ((rc = sdx_scratch(c, 10, sizeof(double) * (size_t)n_slots * S + 64, &p_v))) return rc;
    if ((rc = sdx_scratch(c, 1, (sizeof(PermSlot) + sizeof(unsigned long long)) * (size_t)n_slots + 64, &p_out))) return rc;
    int32_t *d_mod = (int32_t *)p_in, *d_prof = prof ? d_mod + n_slots : nullptr;
    PermSlot *d_slots = (PermSlot *)p_out;
    unsigned long long *d_counts = (unsigned long long *)(d_slots + n_slots);
    SDX_CUDA(c, cudaMemcpyAsync(d_mod, modules, sizeof(int32_t) * n_slots, cudaMemcpyHostToDevice, c->stream));
    if (prof) SDX_CUDA(c, cudaMemcpyAsync(d_prof, prof, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaMemsetAsync(d_counts, 0, sizeof(unsigned long long) * n_slots, c->stream));
    cudaEventRecord(c->ev[2], c->stream);
    k_condperm_setup<<<ceil_div(n_slots, 64), 64, 0, c->stream>>>(c->d_feat, c->wprS, S, c->d_w, c->d_mask, d_prof, d_mod,
                                                                  n_jobs, k, (double *)p_v, d_slots);
    timer.count();
    if (n_perm > 0) {
        if (smem > 48 * 1024)
            SDX_CUDA(c, cudaFuncSetAttribute(k_condperm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int gx = ceil_div(n_perm, threads);
        const int cap = std::max(1, (c->sm_count * 16) / std::max(1, n_slots));
        if (gx > cap) gx = cap;
        k_condperm<<<dim3(gx, n_slots), threads, smem, c->stream>>>(make_keys(seed), (const double *)p_v, S, d_slots, n_slots, k,
                                                                    perm_id0, n_perm, words, d_counts);
        timer.count();
    }
    cudaEventRecord(c->ev[3], c->stream);
    std::vector<PermSlot> hs((size_t)n_slots);
    std::vector<unsigned long long> hc((size_t)n_slots);
    SDX_CUDA(c, cudaMemcpyAsync(hs.data(), d_slots, sizeof(PermSlot) * n_slots, cudaMemcpyDeviceToHost, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(hc.data(), d_counts, sizeof(unsigned long long) * n_slots, cudaMemcpyDeviceToHost, c->stream));
    rc = timer.finish();
    if (rc) return rc;
    cudaEventElapsedTime(&c->timing.score_ms, c->ev[2], c->ev[3]);
    for (int i = 0; i < n_slots; ++i) {
        counts[i] = hs[i].n_i < 0 ? -1 : (int64_t)hc[i];
        if (real_W) real_W[i] = hs[i].real_W;
    }
    return SDX_OK;
}

extern "C" int sdx_condperm(sdx_ctx *c, int profile, const int32_t *module, int k, uint64_t seed, uint64_t perm_id0,
                            int64_t n_perm, int64_t *counts, double *real_W) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, profile >= 0 && profile < c->P, SDX_ERR_INVALID, "profile index out of range");
    const int32_t pj = profile;
    return condperm_run(c, &pj, module, 1, k, seed, perm_id0, n_perm, counts, real_W);
}

extern "C" int sdx_condperm_batch(sdx_ctx *c, const int32_t *profile_of_job, const int32_t *modules, int64_t n_jobs, int k,
                                  uint64_t seed, uint64_t perm_id0, int64_t n_perm, int64_t *counts, double *real_W) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, n_jobs >= 0 && n_jobs < (1 << 24) && k >= 1 && k <= SDX_MAX_K, SDX_ERR_INVALID, "n_jobs / k out of range");
    // chunks keep (job, slot) pairs within the grid.y limit and the v buffer small
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(65535 / k, ((int64_t)512 << 20) / ((int64_t)sizeof(double) * k * std::max(1, c->S))));
    float ms = 0;
    int launches = 0;
    for (int64_t j0 = 0; j0 < n_jobs; j0 += chunk) {
        const int nj = (int)std::min(chunk, n_jobs - j0);
        int rc = condperm_run(c, profile_of_job ? profile_of_job + j0 : nullptr, modules + j0 * k, nj, k, seed, perm_id0, n_perm,
                              counts + j0 * k, real_W ? real_W + j0 * k : nullptr);
        if (rc) return rc;
        ms += c->timing.score_ms; launches += c->timing.launches;
    }
    c->timing.score_ms = ms; c->timing.total_ms = ms; c->timing.launches = launches;
    return SDX_OK;
}

// ---------------------------------------------------------------------------
// Rank-sum: one CTA per (profile, module) job.  The profile's values over its
// samples are staged in shared memory; for every carrier x of the union the
// CTA counts less = #{v < x} and equal = #{v == x}: twice its mid-rank is
// 2*less + equal + 1 (an exact integer), summed exactly in 64-bit.
// z and p follow scipy.stats.ranksums: no tie or continuity correction.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ranksum(const uint32_t *__restrict__ feat, int wprS, int S,
                                                 const double *__restrict__ w_all, const uint32_t *__restrict__ mask_all,
                                                 const int32_t *__restrict__ prof, const int32_t *__restrict__ modules, int k,
                                                 double *__restrict__ z_out, double *__restrict__ p_out,
                                                 long long *__restrict__ s2_out, int32_t *__restrict__ n1_out) {
    extern __shared__ double vals[];              // compacted values of the profile's samples, then carrier flags
    __shared__ int s_ns;
    __shared__ unsigned long long s_sum;
    __shared__ int s_n1;
    const int job = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int p = prof ? prof[job] : 0;
    const double *w = w_all + (size_t)p * S;
    const uint32_t *mask = mask_all + (size_t)p * wprS;
    const int32_t *mod = modules + (size_t)job * k;
    uint8_t *carrier = (uint8_t *)(vals + S);
    if (tid == 0) { s_sum = 0ull; s_n1 = 0; }
    // ordered compaction of the masked samples (warp 0), so vals[] is in ascending sample order
    if (tid < 32) {
        int base = 0;
        for (int s0 = 0; s0 < S; s0 += 32) {
            const int s = s0 + lane;
            const bool in = s < S && ((mask[s >> 5] >> (s & 31)) & 1u);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) {
                const int q = base + __popc(m & ((1u << lane) - 1u));
                vals[q] = w[s];
                bool car = false;
                for (int j = 0; j < k; ++j)
                    if (mod[j] >= 0 && ((feat[(size_t)mod[j] * wprS + (s >> 5)] >> (s & 31)) & 1u)) car = true;
                carrier[q] = car ? 1 : 0;
            }
            base += __popc(m);
        }
        if (lane == 0) s_ns = base;
    }
    __syncthreads();
    const int ns = s_ns;
    unsigned long long my = 0;
    int my_n1 = 0;
    for (int q = tid; q < ns; q += blockDim.x) {
        if (!carrier[q]) continue;
        const double x = vals[q];
        int less = 0, equal = 0;
        for (int r = 0; r < ns; ++r) {
            const double y = vals[r];
            less += y < x;
            equal += y == x;
        }
        my += 2ull * less + equal + 1ull;
        ++my_n1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my += __shfl_xor_sync(0xffffffffu, my, o);
        my_n1 += __shfl_xor_sync(0xffffffffu, my_n1, o);
    }
    if (lane == 0) { atomicAdd(&s_sum, my); atomicAdd(&s_n1, my_n1); }
    __syncthreads();
    if (tid == 0) {
        const long long s2 = (long long)s_sum;
        const int n1 = s_n1, n2 = ns - n1;
        const double expected = n1 * (double)(n1 + n2 + 1) / 2.0;
        const double zz = (0.5 * (double)s2 - expected) / sqrt((double)n1 * n2 * (n1 + n2 + 1) / 12.0);
        z_out[job] = zz;
        p_out[job] = erfc(fabs(zz) / sqrt(2.0));
        s2_out[job] = s2;
        n1_out[job] = n1;
    }
}

extern "C" int sdx_ranksum(sdx_ctx *c, const int32_t *prof, const int32_t *modules, int64_t n_jobs, int k, double *z,
                           double *p, int64_t *twice_rank_sum, int32_t *n1) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat && c->d_w, SDX_ERR_STATE, "matrix and weights must be uploaded first");
    SDX_REQUIRE(c, n_jobs >= 0 && n_jobs < (1ll << 31) && (n_jobs == 0 || modules), SDX_ERR_INVALID, "bad job list");
    SDX_REQUIRE(c, k >= 1 && k <= SDX_MAX_K, SDX_ERR_INVALID, "module size k must be in 1..SDX_MAX_K");
    for (int64_t i = 0; i < n_jobs * k; ++i)
        SDX_REQUIRE(c, modules[i] >= -1 && modules[i] < c->F, SDX_ERR_INVALID, "module feature index out of range");
    if (prof) for (int64_t j = 0; j < n_jobs; ++j)
        SDX_REQUIRE(c, prof[j] >= 0 && prof[j] < c->P, SDX_ERR_INVALID, "profile index out of range");
    const size_t smem = (size_t)c->S * 9 + 16;
    SDX_REQUIRE(c, smem <= c->smem_optin, SDX_ERR_UNSUPPORTED, "too many samples for the rank-sum kernel");
    LaunchTimer timer(c);
    if (n_jobs == 0) return timer.finish();
    int rc;
    void *p_in, *p_out;
    if ((rc = sdx_scratch(c, 0, sizeof(int32_t) * (size_t)n_jobs * (k + 1) + 64, &p_in))) return rc;
    if ((rc = sdx_scratch(c, 1, (2 * sizeof(double) + sizeof(long long) + sizeof(int32_t)) * (size_t)n_jobs + 64, &p_out))) return rc;
    int32_t *d_mod = (int32_t *)p_in, *d_prof = prof ? d_mod + n_jobs * k : nullptr;
    double *d_z = (double *)p_out, *d_p = d_z + n_jobs;
    long long *d_s2 = (long long *)(d_p + n_jobs);
    int32_t *d_n1 = (int32_t *)(d_s2 + n_jobs);
    SDX_CUDA(c, cudaMemcpyAsync(d_mod, modules, sizeof(int32_t) * n_jobs * k, cudaMemcpyHostToDevice, c->stream));
    if (prof) SDX_CUDA(c, cudaMemcpyAsync(d_prof, prof, sizeof(int32_t) * n_jobs, cudaMemcpyHostToDevice, c->stream));
    if (smem > 48 * 1024)
        SDX_CUDA(c, cudaFuncSetAttribute(k_ranksum, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEventRecord(c->ev[2], c->stream);
    k_ranksum<<<(unsigned)n_jobs, 256, smem, c->stream>>>(c->d_feat, c->wprS, c->S, c->d_w, c->d_mask, d_prof, d_mod, k, d_z, d_p,
                                                         d_s2, d_n1);
    cudaEventRecord(c->ev[3], c->stream);
    timer.count();
    if (z) SDX_CUDA(c, cudaMemcpyAsync(z, d_z, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost, c->stream));
    if (p) SDX_CUDA(c, cudaMemcpyAsync(p, d_p, sizeof(double) * n_jobs, cudaMemcpyDeviceToHost, c->stream));
    if (twice_rank_sum) SDX_CUDA(c, cudaMemcpyAsync(twice_rank_sum, d_s2, sizeof(long long) * n_jobs, cudaMemcpyDeviceToHost, c->stream));
    if (n1) SDX_CUDA(c, cudaMemcpyAsync(n1, d_n1, sizeof(int32_t) * n_jobs, cudaMemcpyDeviceToHost, c->stream));
    rc = timer.finish();
    cudaEventElapsedTime(&c->timing.score_ms, c->ev[2], c->ev[3]);
    return rc;
}
