// sdx_api.cu — context lifecycle, uploads and layout kernels of libsdx.so.
#include <stdlib.h>

#include "sdx_common.cuh"

std::string g_sdx_create_error;

extern "C" int sdx_version(void) { return SDX_VERSION; }

extern "C" int sdx_device_count(int *n) {
    if (!n) return SDX_ERR_INVALID;
    int k = 0;
    cudaError_t e = cudaGetDeviceCount(&k);
    if (e != cudaSuccess) { *n = 0; g_sdx_create_error = cudaGetErrorString(e); return SDX_ERR_CUDA; }
    *n = k;
    return SDX_OK;
}

extern "C" const char *sdx_last_error(const sdx_ctx *ctx) {
    return ctx ? ctx->err.c_str() : g_sdx_create_error.c_str();
}

extern "C" int sdx_create(int device, sdx_ctx **out) {
    if (!out) return SDX_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_sdx_create_error = std::string("no CUDA device available (libsdx has no CPU fallback): ") +
                             (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return SDX_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_sdx_create_error = "device index out of range"; return SDX_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_sdx_create_error = cudaGetErrorString(e);
        return SDX_ERR_CUDA;
    }
    if (prop.major != 10) {
        char b[256];
        snprintf(b, sizeof b, "libsdx is built for sm_100a only; device %d is sm_%d%d (%s)", device, prop.major, prop.minor, prop.name);
        g_sdx_create_error = b;
        return SDX_ERR_UNSUPPORTED;
    }
    sdx_ctx *c = new sdx_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_sdx_create_error = cudaGetErrorString(e);
        delete c;
        return SDX_ERR_CUDA;
    }
    c->stream = c->own_stream;
    for (auto &ev : c->ev) cudaEventCreate(&ev);
    {   // tuning knobs (benchmarks and tests only): read here, once, never on the call path
        auto geti = [](const char *n, long dflt) { const char *e = getenv(n); return e && *e ? atol(e) : dflt; };
        c->knobs.trade_g = (int)geti("SDX_TRADE_G", 0);
        c->knobs.trade_warps = (int)geti("SDX_TRADE_WARPS", 0);
        c->knobs.trade_lb = (int)geti("SDX_TRADE_LB", -1);
        c->knobs.plan_cap = (int)geti("SDX_PLAN_CAP", 0);
        c->knobs.grid_cap = (int)geti("SDX_TRADE_GRID_CAP", 0);
        c->knobs.nulls_per_launch = geti("SDX_NULLS_PER_LAUNCH", 0);
        c->knobs.pad_smem = geti("SDX_TRADE_PAD_SMEM", 0);
        c->knobs.lanes = (int)geti("SDX_LANES", 1);
        c->knobs.lane_ts = (int)geti("SDX_LANE_TS", 0);
        c->knobs.lane_wpc = (int)geti("SDX_LANE_WPC", 0);
        c->knobs.lane_min_chains = (int)geti("SDX_LANE_MIN_CHAINS", 0);
        c->knobs.lane_debug = (int)geti("SDX_LANE_DEBUG", 0);
    }
    *out = c;
    return SDX_OK;
}

static void free_matrix(sdx_ctx *c) {
    if (c->d_feat) cudaFree(c->d_feat);
    if (c->d_samp) cudaFree(c->d_samp);
    if (c->d_trade_obs) cudaFree(c->d_trade_obs);
    if (c->d_rowsize) cudaFree(c->d_rowsize);
    if (c->d_ln_desc) cudaFree(c->d_ln_desc);
    if (c->d_ln_obs) cudaFree(c->d_ln_obs);
    if (c->d_ln_pool) cudaFree(c->d_ln_pool);
    c->d_feat = c->d_samp = c->d_trade_obs = nullptr;
    c->d_rowsize = nullptr;
    c->d_ln_desc = nullptr; c->d_ln_obs = nullptr; c->d_ln_pool = nullptr; c->cap_ln_pool = 0; c->cap_ln_desc = 0; c->cap_ln_obs = 0;
    c->ln_state = 0; c->ln_geo_valid = false; c->ln_pool_valid = false;
    if (c->d_pool) cudaFree(c->d_pool);
    c->d_pool = nullptr;
    c->pool_valid = false;
    if (c->resume.d_state) cudaFree(c->resume.d_state);
    c->resume.d_state = nullptr; c->resume.cap = 0; c->resume.valid = false;
    if (c->d_top) cudaFree(c->d_top);
    c->d_top = nullptr; c->cap_top = 0; c->top_K = 0;
    if (c->d_burn_obs) cudaFree(c->d_burn_obs);
    if (c->d_burn_rowsize) cudaFree(c->d_burn_rowsize);
    c->d_burn_obs = nullptr; c->d_burn_rowsize = nullptr; c->cap_burn_obs = c->cap_burn_rowsize = 0;
    c->cap_feat = c->cap_samp = c->cap_obs = c->cap_rowsize = c->cap_pool = 0;
}

// a new matrix invalidates everything derived from the old one; the buffers themselves are kept and reused
static void invalidate_matrix(sdx_ctx *c) {
    c->ln_state = 0; c->ln_geo_valid = false;                  // the lane layout (row kinds and offsets) is rebuilt on first use
    c->ln_pool_valid = false;
    c->pool_valid = false;
    c->resume.valid = false;
    c->top_K = 0;
}

static void free_weights(sdx_ctx *c) {
    if (c->d_w) cudaFree(c->d_w);
    if (c->d_mask) cudaFree(c->d_mask);
    if (c->d_nsamp) cudaFree(c->d_nsamp);
    c->d_w = nullptr; c->d_mask = nullptr; c->d_nsamp = nullptr;
    c->cap_w = c->cap_mask = c->cap_nsamp = 0;
    c->P = 0;
}

extern "C" int sdx_destroy(sdx_ctx *c) {
    if (!c) return SDX_ERR_INVALID;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    free_matrix(c);
    free_weights(c);
    for (auto &b : c->scratch) if (b.p) cudaFree(b.p);
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
    return SDX_OK;
}

extern "C" int sdx_set_stream(sdx_ctx *c, void *s) {
    if (!c) return SDX_ERR_INVALID;
    cudaStreamSynchronize(c->stream);
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return SDX_OK;
}

extern "C" int sdx_set_burn_in(sdx_ctx *c, int64_t burn_rounds, int pool_size) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, burn_rounds >= 0 && burn_rounds <= 256, SDX_ERR_INVALID, "burn_rounds must be in 0..256");
    SDX_REQUIRE(c, burn_rounds == 0 || (pool_size >= 1 && pool_size <= 65536), SDX_ERR_INVALID, "pool_size must be in 1..65536");
    if (burn_rounds != c->burn || pool_size != c->pool_size) { c->pool_valid = false; c->ln_pool_valid = false; c->resume.valid = false; }
    c->burn = burn_rounds;
    c->pool_size = burn_rounds > 0 ? pool_size : 0;
    if (burn_rounds == 0) c->top_K = 0;          // no burn-in trades: the dense-row table of the pool is not in use
    return SDX_OK;
}

extern "C" int sdx_set_pair_group(sdx_ctx *c, int pair_group) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, pair_group == 1 || pair_group == 32, SDX_ERR_INVALID, "pair_group must be 1 or 32");
    c->pair_group = pair_group;
    return SDX_OK;
}

extern "C" int sdx_get_timing(const sdx_ctx *c, sdx_timing *out) {
    if (!c || !out) return SDX_ERR_INVALID;
    *out = c->timing;
    return SDX_OK;
}

extern "C" int sdx_total_launches(const sdx_ctx *c, int64_t *n) {
    if (!c || !n) return SDX_ERR_INVALID;
    *n = c->total_launches;
    return SDX_OK;
}

// ---------------------------------------------------------------------------
// bit transpose: n_rows x wpr_in (rows over n_cols bits) -> n_cols x wpr_out.
// One warp per 32 x 32 bit tile (row word rw of the output, column word cw of the input): lane l loads word cw of row
// 32 rw + l, five shuffle-and-mask steps swap ever smaller off-diagonal blocks (the recursive transpose), after which
// lane l holds word rw of output row 32 cw + l.
// ---------------------------------------------------------------------------
__global__ void k_transpose_bits(const uint32_t *__restrict__ in, int n_rows, int wpr_in, int n_cols,
                                 uint32_t *__restrict__ out, int wpr_out) {
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    in += (size_t)blockIdx.y * n_rows * wpr_in;              // blockIdx.y = matrix of a batch
    out += (size_t)blockIdx.y * n_cols * wpr_out;
    const int tiles_c = wpr_in;
    const int rw = tile / tiles_c, cw = tile - rw * tiles_c;     // row word (of out), col word (of in)
    if (rw >= wpr_out) return;
    const int r = rw * 32 + lane;
    uint32_t x = (r < n_rows) ? in[(size_t)r * wpr_in + cw] : 0u;
    uint32_t m = 0x0000ffffu;
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t p = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? (((p >> j) & m) | (x & ~m)) : ((x & m) | ((p & m) << j));
        m ^= m << (j >> 1);
    }
    const int col = cw * 32 + lane;
    if (col < n_cols) out[(size_t)col * wpr_out + rw] = x;
}

// popcount of every trade row: one warp per row
__global__ void k_row_sizes(const uint32_t *__restrict__ rows, int R, int RS, int *__restrict__ size) {
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= R) return;
    int c = 0;
    for (int i = lane; i < RS; i += 32) c += __popc(rows[(size_t)r * RS + i]);
    c = warp_sum_i32(c);
    if (lane == 0) size[r] = c;
}

int sdx_row_sizes(sdx_ctx *c, const uint32_t *d_rows, int R, int RS, int *d_size) {
    k_row_sizes<<<ceil_div(R, 8), 256, 0, c->stream>>>(d_rows, R, RS, d_size);
    c->total_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sdx_fail(c, SDX_ERR_CUDA, std::string("k_row_sizes launch failed: ") + cudaGetErrorString(e));
    return SDX_OK;
}

// batch of nb bit matrices (n_rows x wpr_in words each, dense) -> their transposes (n_cols x wpr_out words each).  Every word of
// the output is written (n_cols <= 32 * wpr_in), so it needs no clearing; row strides padded with zero words are fine.
int sdx_transpose_batch(sdx_ctx *c, const uint32_t *d_in, int n_rows, int wpr_in, int n_cols, uint32_t *d_out, int wpr_out, int nb) {
    const int tiles = wpr_out * wpr_in;
    k_transpose_bits<<<dim3(ceil_div(tiles, 8), nb), 256, 0, c->stream>>>(d_in, n_rows, wpr_in, n_cols, d_out, wpr_out);
    c->total_launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sdx_fail(c, SDX_ERR_CUDA, std::string("k_transpose_bits launch failed: ") + cudaGetErrorString(e));
    return SDX_OK;
}

extern "C" int sdx_upload_matrix(sdx_ctx *c, const uint32_t *rows, int F, int S, int wpr) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, rows && F >= 1 && S >= 1, SDX_ERR_INVALID, "matrix pointer / shape invalid");
    SDX_REQUIRE(c, wpr == (S + 31) / 32, SDX_ERR_INVALID, "words_per_row must be ceil(n_samples/32)");
    if (S % 32) {        // padding bits must be zero
        const uint32_t pad = ~0u << (S % 32);
        for (int g = 0; g < F; ++g)
            SDX_REQUIRE(c, (rows[(size_t)g * wpr + wpr - 1] & pad) == 0, SDX_ERR_INVALID, "padding bits of a feature row are not zero");
    }
    SDX_CUDA(c, cudaSetDevice(c->device));
    invalidate_matrix(c);
    if (S != c->S) free_weights(c);          // weights are per sample: a matrix over other samples needs new ones
    c->F = F; c->S = S; c->wprS = wpr; c->wprF = (F + 31) / 32;
    SDX_TRY(sdx_reserve(c, c->d_feat, c->cap_feat, sizeof(uint32_t) * (size_t)F * c->wprS));
    SDX_TRY(sdx_reserve(c, c->d_samp, c->cap_samp, sizeof(uint32_t) * (size_t)S * c->wprF));
    SDX_CUDA(c, cudaMemcpyAsync(c->d_feat, rows, sizeof(uint32_t) * (size_t)F * c->wprS, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaMemsetAsync(c->d_samp, 0, sizeof(uint32_t) * (size_t)S * c->wprF, c->stream));
    const int tiles = c->wprF * c->wprS;
    k_transpose_bits<<<ceil_div(tiles, 8), 256, 0, c->stream>>>(c->d_feat, F, c->wprS, S, c->d_samp, c->wprF);
    c->total_launches++;
    // trade orientation (src/curveball.py:10-11): rows over the shorter dimension, samples when F >= S
    c->rows_are_samples = F >= S;
    c->R = c->rows_are_samples ? S : F;
    c->L = c->rows_are_samples ? F : S;
    c->wprL = c->rows_are_samples ? c->wprF : c->wprS;
    c->RS = ((c->wprL + 3) / 4) * 4;
    SDX_TRY(sdx_reserve(c, c->d_trade_obs, c->cap_obs, sizeof(uint32_t) * (size_t)c->R * c->RS));
    SDX_CUDA(c, cudaMemsetAsync(c->d_trade_obs, 0, sizeof(uint32_t) * (size_t)c->R * c->RS, c->stream));
    SDX_CUDA(c, cudaMemcpy2DAsync(c->d_trade_obs, (size_t)c->RS * 4, c->rows_are_samples ? c->d_samp : c->d_feat,
                                  (size_t)c->wprL * 4, (size_t)c->wprL * 4, c->R, cudaMemcpyDeviceToDevice, c->stream));
    SDX_TRY(sdx_reserve(c, c->d_rowsize, c->cap_rowsize, sizeof(int) * (size_t)c->R));
    k_row_sizes<<<ceil_div(c->R, 8), 256, 0, c->stream>>>(c->d_trade_obs, c->R, c->RS, c->d_rowsize);
    c->total_launches++;
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    SDX_CUDA(c, cudaGetLastError());
    return SDX_OK;
}

extern "C" int sdx_trade_shape(const sdx_ctx *c, int *R, int *L, int *wpr, int *ras) {
    if (!c) return SDX_ERR_INVALID;
    if (!c->d_trade_obs) return SDX_ERR_STATE;
    if (R) *R = c->R;
    if (L) *L = c->L;
    if (wpr) *wpr = c->wprL;
    if (ras) *ras = c->rows_are_samples;
    return SDX_OK;
}

extern "C" int sdx_upload_weights(sdx_ctx *c, const double *w, const uint32_t *mask, int P, int S) {
    if (!c) return SDX_ERR_INVALID;
    SDX_REQUIRE(c, c->d_feat, SDX_ERR_STATE, "sdx_upload_matrix must be called before sdx_upload_weights");
    SDX_REQUIRE(c, w && P >= 1 && S == c->S, SDX_ERR_INVALID, "weights pointer / shape invalid (n_samples must match the matrix)");
    const int wpr = c->wprS;
    std::vector<double> hw((size_t)P * S);
    std::vector<uint32_t> hm((size_t)P * wpr);
    std::vector<int> hn((size_t)P);
    for (int p = 0; p < P; ++p) {
        int cnt = 0;
        for (int s = 0; s < S; ++s) {
            bool in = mask ? ((mask[(size_t)p * wpr + (s >> 5)] >> (s & 31)) & 1u) : true;
            double v = w[(size_t)p * S + s];
            SDX_REQUIRE(c, !in || v == v, SDX_ERR_INVALID, "NaN weight for a sample inside the profile mask");
            hw[(size_t)p * S + s] = in ? v : 0.0;
            if (in) { hm[(size_t)p * wpr + (s >> 5)] |= 1u << (s & 31); ++cnt; }
        }
        hn[p] = cnt;
    }
    SDX_CUDA(c, cudaSetDevice(c->device));
    c->P = 0;
    SDX_TRY(sdx_reserve(c, c->d_w, c->cap_w, sizeof(double) * (size_t)P * S));
    SDX_TRY(sdx_reserve(c, c->d_mask, c->cap_mask, sizeof(uint32_t) * (size_t)P * wpr));
    SDX_TRY(sdx_reserve(c, c->d_nsamp, c->cap_nsamp, sizeof(int) * (size_t)P));
    SDX_CUDA(c, cudaMemcpyAsync(c->d_w, hw.data(), sizeof(double) * (size_t)P * S, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(c->d_mask, hm.data(), sizeof(uint32_t) * (size_t)P * wpr, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaMemcpyAsync(c->d_nsamp, hn.data(), sizeof(int) * (size_t)P, cudaMemcpyHostToDevice, c->stream));
    SDX_CUDA(c, cudaStreamSynchronize(c->stream));
    c->P = P;
    return SDX_OK;
}
