// Conformal calibration kernels.
//
// Reference behaviour: /root/reference/1D/inference/guidance.py:9-66 (calculate_guidance, get_weight,
// normalize_weights), /root/reference/1D/inference/conformal.py:74-85 (nonconformity score) and :95-118
// (calculate_quantile = rank-th order statistic via torch.sort).
// All of these are latency-bound at the reference's sizes (n = 1,000 ... 50,000 scalars); the point of the
// device versions is that calibration never leaves the GPU or synchronises the host.
#include "common.cuh"
#include <math.h>

namespace sdc {

// block-wide reduction of the safety statistic of one sample: red(scaler * x[b,2,:nt,:])
__device__ float block_safety_stat(const float* xb, int use_mean, float scaler, int nt, int H, int W, float* red) {
    const float4* p = reinterpret_cast<const float4*>(xb + 2 * H * W);
    const int n4 = nt * W / 4;
    float acc = use_mean ? 0.f : -INFINITY;
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 v = p[i];
        float a = __fmul_rn(v.x, scaler), b = __fmul_rn(v.y, scaler), c = __fmul_rn(v.z, scaler), d = __fmul_rn(v.w, scaler);
        acc = use_mean ? acc + ((a + b) + (c + d)) : fmaxf(acc, fmaxf(fmaxf(a, b), fmaxf(c, d)));
    }
    acc = use_mean ? warp_sum(acc) : warp_max(acc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    float t = red[0];
    for (int w = 1; w < (blockDim.x >> 5); ++w) t = use_mean ? t + red[w] : fmaxf(t, red[w]);
    return use_mean ? t / (float)(nt * W) : t;
}

__global__ void __launch_bounds__(128) safety_stat_kernel(const float* x, float* stat, int use_mean, float scaler, int nt,
                                                          int H, int W) {
    __shared__ float red[4];
    const int64_t b = blockIdx.x;
    float s = block_safety_stat(x + b * 3 * H * W, use_mean, scaler, nt, H, W, red);
    if (threadIdx.x == 0) stat[b] = s;
}

__device__ __forceinline__ float weight_of(float stat, float Q, const sdc_guidance& g) {
    float margin = __fsub_rn(__fadd_rn(stat, Q), g.u_bound_sq);
    float guid = __fmul_rn(fmaxf(margin, 0.f), g.w_score);
    if (margin != margin) guid = margin;  // torch.maximum propagates NaN
    return expf(-guid);
}

__global__ void __launch_bounds__(128) conformal_scores_kernel(const float* pred, const float* state, float* score,
                                                               float* weight, sdc_guidance g, float Q2, int H, int W) {
    __shared__ float red[4];
    const int64_t b = blockIdx.x;
    const int use_mean = g.mode == 1;
    float st = block_safety_stat(state + b * 3 * H * W, use_mean, g.scaler, g.nt, H, W, red);
    float sp = pred ? block_safety_stat(pred + b * 3 * H * W, use_mean, g.scaler, g.nt, H, W, red) : 0.f;
    if (threadIdx.x == 0) {
        if (score && pred) score[b] = fabsf(__fsub_rn(sp, st));
        if (weight) {
            float w = weight_of(st, g.Q, g);
            if (!isinf(Q2) && Q2 == Q2) w = __fmul_rn(w, weight_of(st, Q2, g));
            weight[b] = w;
        }
    }
}

// Single-CTA: n is small (<= a few 100k) and the result must not depend on the launch geometry.
__global__ void __launch_bounds__(1024) normalize_weights_kernel(float* w, float* out, float* scores, int64_t n) {
    __shared__ float red[32];
    __shared__ int flag;
    __shared__ float s_val;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // 1. any inf?  largest non-inf value
    int has_inf = 0;
    float mx = -INFINITY;
    bool any_fin = false;
    for (int64_t i = tid; i < n; i += blockDim.x) {
        float v = w[i];
        if (isinf(v)) has_inf = 1; else { mx = any_fin ? fmaxf(mx, v) : v; any_fin = true; }
    }
    has_inf = __syncthreads_or(has_inf);
    if (has_inf) {
        mx = warp_max(mx);
        if (lane == 0) red[wid] = mx;
        __syncthreads();
        if (tid == 0) { float t = red[0]; for (int k = 1; k < 32; ++k) t = fmaxf(t, red[k]); s_val = t; }
        __syncthreads();
        const float rep = s_val;
        for (int64_t i = tid; i < n; i += blockDim.x) if (isinf(w[i])) w[i] = rep;
        __syncthreads();
    }
    // 2. sum (deterministic: strided partials -> warp tree -> serial over warps)
    float acc = 0.f;
    for (int64_t i = tid; i < n; i += blockDim.x) acc += w[i];
    acc = warp_sum(acc);
    __syncthreads();
    if (lane == 0) red[wid] = acc;
    __syncthreads();
    if (tid == 0) { float t = 0.f; for (int k = 0; k < 32; ++k) t += red[k]; s_val = t; flag = (t == 0.f); }
    __syncthreads();
    const float sum = s_val, nf = (float)n;
    const int zero = flag;
    for (int64_t i = tid; i < n; i += blockDim.x) {
        float o = zero ? 1.0f : __fdiv_rn(__fmul_rn(nf, w[i]), sum);
        out[i] = o;
        if (scores) scores[i] = __fmul_rn(o, scores[i]);
    }
}

// order-preserving key: ascending float order (torch.sort semantics: -0 == +0, NaN last) -> ascending uint32
__device__ __forceinline__ uint32_t sort_key(float v) {
    if (v != v) return 0xFFFFFFFFu;
    uint32_t b = __float_as_uint(v);
    if (b == 0x80000000u) b = 0u;
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(1024) kth_select_kernel(const float* s, int64_t n, int64_t rank, float* value_out,
                                                          int64_t* index_out) {
    __shared__ unsigned int hist[256];
    __shared__ uint32_t s_prefix;
    __shared__ int64_t s_rank;
    __shared__ int64_t scan[1024];
    const int tid = threadIdx.x;
    uint32_t prefix = 0, mask = 0;
    int64_t r = rank;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) hist[tid] = 0;
        __syncthreads();
        for (int64_t i = tid; i < n; i += blockDim.x) {
            uint32_t k = sort_key(s[i]);
            if ((k & mask) == prefix) atomicAdd(&hist[(k >> shift) & 0xFF], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int64_t rr = r;
            int bin = 0;
            for (; bin < 255; ++bin) {
                if (rr < (int64_t)hist[bin]) break;
                rr -= hist[bin];
            }
            s_prefix = prefix | ((uint32_t)bin << shift);
            s_rank = rr;
        }
        __syncthreads();
        prefix = s_prefix;
        r = s_rank;
        mask |= 0xFFu << shift;
        __syncthreads();
    }
    // r-th (0-based) occurrence of `prefix` in index order == what a stable sort selects
    const int64_t chunk = (n + blockDim.x - 1) / blockDim.x;
    const int64_t lo = (int64_t)tid * chunk, hi = min(n, lo + chunk);
    int64_t cnt = 0;
    for (int64_t i = lo; i < hi; ++i) cnt += (sort_key(s[i]) == prefix) ? 1 : 0;
    scan[tid] = cnt;
    __syncthreads();
    if (tid == 0) {
        int64_t run = 0;
        for (int k = 0; k < (int)blockDim.x; ++k) { int64_t c = scan[k]; scan[k] = run; run += c; }
    }
    __syncthreads();
    const int64_t before = scan[tid];
    if (r >= before && r < before + cnt) {
        int64_t want = r - before;
        for (int64_t i = lo; i < hi; ++i) {
            if (sort_key(s[i]) == prefix) {
                if (want == 0) {
                    if (value_out) *value_out = s[i];
                    if (index_out) *index_out = i;
                    break;
                }
                --want;
            }
        }
    }
}

}  // namespace sdc

using namespace sdc;

extern "C" int sdc_safety_stat(const float* x, float* stat, int use_mean, float scaler, int nt, int64_t B, int H, int W,
                               void* stream) {
    SDC_REQUIRE(B >= 0 && B < (1LL << 31) && H > 0 && W > 0 && nt > 0 && nt <= H && (nt * W) % 4 == 0 && (H * W) % 4 == 0,
                "safety_stat: bad sizes");
    if (B == 0) return SDC_OK;
    SDC_REQUIRE(x && stat, "safety_stat: null pointer");
    safety_stat_kernel<<<(unsigned)B, 128, 0, as_stream(stream)>>>(x, stat, use_mean, scaler, nt, H, W);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_conformal_scores(const float* pred, const float* state, float* score, float* weight,
                                    const sdc_guidance* g, float Q2, int64_t B, int H, int W, void* stream) {
    SDC_REQUIRE(g != nullptr && (g->mode == 1 || g->mode == 2), "conformal_scores: guidance mode must be 1 (mean) or 2 (amax)");
    SDC_REQUIRE(B >= 0 && B < (1LL << 31) && H > 0 && W > 0 && g->nt > 0 && g->nt <= H && (g->nt * W) % 4 == 0 &&
                    (H * W) % 4 == 0, "conformal_scores: bad sizes");
    if (B == 0) return SDC_OK;
    SDC_REQUIRE(state != nullptr, "conformal_scores: null state");
    conformal_scores_kernel<<<(unsigned)B, 128, 0, as_stream(stream)>>>(pred, state, score, weight, *g, Q2, H, W);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_normalize_weights(float* w, float* out, float* scores, int64_t n, void* stream) {
    SDC_REQUIRE(n >= 0, "normalize_weights: n < 0");
    if (n == 0) return SDC_OK;
    SDC_REQUIRE(w && out, "normalize_weights: null pointer");
    normalize_weights_kernel<<<1, 1024, 0, as_stream(stream)>>>(w, out, scores, n);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int64_t sdc_kth_select_workspace(int64_t n) { (void)n; return 256; }

extern "C" int sdc_kth_select(const float* scores, int64_t n, int64_t rank, float* value_out, int64_t* index_out,
                              void* workspace, void* stream) {
    (void)workspace;
    SDC_REQUIRE(n > 0 && rank >= 0 && rank < n, "kth_select: need 0 <= rank < n (rank=%lld n=%lld)", (long long)rank, (long long)n);
    SDC_REQUIRE(scores != nullptr, "kth_select: null pointer");
    kth_select_kernel<<<1, 1024, 0, as_stream(stream)>>>(scores, n, rank, value_out, index_out);
    SDC_LAUNCHED();
    return SDC_OK;
}
