// Backward-data (VJP with respect to the denoiser INPUT) building blocks: SURVEY.md section 8 rows A1/A7 — the pass that
// autograd runs through Unet2D when a guidance callable or the `enable_grad` last step differentiates eps_theta(x_t, t)
// with respect to x_t (/root/reference/1D/model/diffusion.py:254-262,524-551).  Parameter gradients are NOT produced.
//
// The dense part (dgrad of every convolution) reuses the tcgen05 implicit-GEMM kernels of conv_gemm.cu / conv_row.cu in
// TF32 mode with weights re-packed by sdc_pack_conv_weight_dgrad (transposed, taps flipped): the data gradient of a
// zero-padded 3x3 convolution is a zero-padded 3x3 convolution of dY.  Gradients live in fp32 containers (fp16 would
// underflow) and are rounded to TF32 by the kernel that produces a dgrad operand.  This file holds the bandwidth-bound
// backward kernels: GroupNorm+FiLM+SiLU, channel LayerNorm, linear / full attention, pixel-(un)shuffle, nearest-upsample,
// the 3-channel head and the col2im of the 7x7 stem.  All tensors are NHWC pixel rows [B*H*W, C].
#include "common.cuh"
#include <stdlib.h>
#include "../../include/safediffcon_b200_unet.h"
#include <math.h>

namespace sdc {

// MUFU.EX2 + MUFU.RCP (relative error ~1e-6; every consumer rounds to TF32 afterwards).  With expf and an IEEE division the two
// GroupNorm backward passes were issue bound, not HBM bound (ncu: 3.4 TB/s at 46 % warp occupancy, ~70 instructions per element).
__device__ __forceinline__ float sigmoidf_(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }
// d silu(z) / dz
__device__ __forceinline__ float dsilu(float z) {
    const float s = sigmoidf_(z);
    return s * (1.0f + z * (1.0f - s));
}

// ---------------------------------------------------------------------------------------------- dgrad weight packing
// kind 1: Wt[ci, tap'*Cout + co] = W[co, ci, 8 - tap']   (spatially flipped taps)
// kind 0: Wt[ci, co] = W[co, ci]
// kind 2: Wt[p*C + c, co] = W[co, c*4 + p]               (rows ordered for the pixel-shuffle that follows)
__global__ void pack_dgrad_weight_kernel(int kind, const float* __restrict__ w, float* __restrict__ wt, int Cout, int Cin) {
    const int taps = kind == 1 ? 9 : 1;
    const int K = taps * Cout;
    const int64_t total = (int64_t)Cin * K;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / K), k = (int)(i % K);
        float v;
        if (kind == 1) {
            const int tp = k / Cout, co = k % Cout;
            v = w[((int64_t)co * Cin + r) * 9 + (8 - tp)];
        } else if (kind == 2) {
            const int C = Cin / 4, p = r / C, c = r % C;
            v = w[(int64_t)k * Cin + c * 4 + p];
        } else {
            v = w[(int64_t)k * Cin + r];
        }
        wt[i] = to_tf32(v);
    }
}

// ---------------------------------------------------------------------------------------------- GroupNorm(1)+FiLM+SiLU backward
// forward: z = x * a_c + b_c, y = silu(z) with a_c = gamma_c rstd (sc_c + 1); xhat = (x - mean) rstd.
// dxhat = dy silu'(z) a_c / rstd;  dx = rstd (dxhat - S1/N - xhat S2/N), S1 = sum dxhat, S2 = sum dxhat xhat per sample.
struct GnCoef {
    float mean, rstd;
};
__device__ __forceinline__ GnCoef gn_prepare(float* coef, const double* stats, const float* gamma, const float* beta,
                                             const float* scale_shift, const int32_t* t_index, int64_t ss_stride, int b, int HW, int C) {
    const double cnt = (double)HW * (double)C;
    const double mean_d = stats[2 * b] / cnt;
    double var_d = stats[2 * b + 1] / cnt - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    GnCoef g;
    g.mean = (float)mean_d;
    g.rstd = (float)(1.0 / sqrt(var_d + 1e-5));
    const float* ss = scale_shift ? scale_shift + (int64_t)(t_index ? t_index[b] : 0) * ss_stride : nullptr;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float gm = gamma[c] * g.rstd;
        float a = gm, bb = beta[c] - g.mean * gm;
        if (ss) {
            const float sc = ss[c] + 1.0f;
            a *= sc;
            bb = bb * sc + ss[C + c];
        }
        coef[c] = a;
        coef[C + c] = bb;
    }
    __syncthreads();
    return g;
}

// pass 1: sums[b] += (S1, S2)
__global__ void __launch_bounds__(256) gn_silu_bwd_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                 const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                                 const int32_t* __restrict__ t_index, int64_t ss_stride,
                                                                 double* __restrict__ sums, int HW, int C, int pix_per_cta) {
    extern __shared__ float coef[];
    __shared__ float red[2][8];
    const int b = blockIdx.x;
    const GnCoef g = gn_prepare(coef, stats, gamma, beta, scale_shift, t_index, ss_stride, b, HW, C);
    const int c4n = C / 4;
    const int64_t row0 = (int64_t)b * HW + (int64_t)blockIdx.y * pix_per_cta;
    const int rows = min(pix_per_cta, HW - (int)blockIdx.y * pix_per_cta);
    const float4* x4 = reinterpret_cast<const float4*>(x + row0 * C);
    const float4* d4 = reinterpret_cast<const float4*>(dy + row0 * C);
    const int total = rows * c4n;
    const float inv_rstd = 1.0f / g.rstd;
    float s1 = 0.f, s2 = 0.f;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        const float4 xv = x4[i], dv = d4[i];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = coef[c + j];
            const float z = fmaf(xs[j], a, coef[C + c + j]);
            const float dxh = ds[j] * dsilu(z) * a * inv_rstd;
            s1 += dxh;
            s2 += dxh * (xs[j] - g.mean) * g.rstd;
        }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < 8; ++w) { a += (double)red[0][w]; c += (double)red[1][w]; }
        atomicAdd(sums + 2 * b, a);
        atomicAdd(sums + 2 * b + 1, c);
    }
}

// pass 2: dx (TF32-rounded: it feeds the dgrad convolution)
__global__ void __launch_bounds__(256) gn_silu_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                                const int32_t* __restrict__ t_index, int64_t ss_stride,
                                                                const double* __restrict__ sums, float* __restrict__ dx, int HW, int C,
                                                                int pix_per_cta) {
    extern __shared__ float coef[];
    const int b = blockIdx.x;
    const GnCoef g = gn_prepare(coef, stats, gamma, beta, scale_shift, t_index, ss_stride, b, HW, C);
    const double cnt = (double)HW * (double)C;
    const float m1 = (float)(sums[2 * b] / cnt) * g.rstd, m2 = (float)(sums[2 * b + 1] / cnt) * g.rstd;
    const int c4n = C / 4;
    const int64_t row0 = (int64_t)b * HW + (int64_t)blockIdx.y * pix_per_cta;
    const int rows = min(pix_per_cta, HW - (int)blockIdx.y * pix_per_cta);
    const float4* x4 = reinterpret_cast<const float4*>(x + row0 * C);
    const float4* d4 = reinterpret_cast<const float4*>(dy + row0 * C);
    float* o = dx + row0 * C;
    const int total = rows * c4n;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        const float4 xv = x4[i], dv = d4[i];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
        float r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = coef[c + j];
            const float z = fmaf(xs[j], a, coef[C + c + j]);
            r[j] = ds[j] * dsilu(z) * a - m1 - (xs[j] - g.mean) * g.rstd * m2;
        }
        store_operand4(o + 4 * (int64_t)i, make_float4(r[0], r[1], r[2], r[3]));
    }
}

// Both passes in ONE kernel (round 2): a cluster of CL CTAs owns one sample.  Pass 1 reads dy and x from HBM and leaves them in L2;
// the per-CTA partial sums meet through distributed shared memory (fixed summation order: deterministic, no atomics); pass 2
// re-reads the same lines while they are still L2 resident (2 CTAs of 512 threads per SM: at most 37 samples x 2 MB in flight at
// the 16x128 level, against 126 MB of L2) and writes dx.  HBM traffic per element: 8 B read + 4 B written instead of 16 + 4 for the
// two-kernel version above (the backward-data pass spent 16.5 ms of 118 ms in those 76 launches at B = 1024).
__device__ __forceinline__ double ld_cluster_f64(const double* local, uint32_t rank) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(local);
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    double v;
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CL>
__global__ void __launch_bounds__(512, 2) gn_silu_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                   const double* __restrict__ stats, const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                                   const int32_t* __restrict__ t_index, int64_t ss_stride,
                                                                   double* __restrict__ sums, float* __restrict__ dx, int HW, int C) {
    extern __shared__ float coef[];
    __shared__ float red[2][16];
    __shared__ double part[2];   // this CTA's (S1, S2): read by every CTA of the cluster
    __shared__ float tot[2];
    const int b = blockIdx.x / CL, rank = blockIdx.x % CL;   // 1-D clusters of CL consecutive CTAs
    const GnCoef g = gn_prepare(coef, stats, gamma, beta, scale_shift, t_index, ss_stride, b, HW, C);
    const int c4n = C / 4;
    const int ppc = (HW + CL - 1) / CL;
    const int rows = max(0, min(ppc, HW - rank * ppc));
    const int64_t row0 = (int64_t)b * HW + (int64_t)rank * ppc;
    const float4* x4 = reinterpret_cast<const float4*>(x + row0 * C);
    const float4* d4 = reinterpret_cast<const float4*>(dy + row0 * C);
    const int total = rows * c4n;
    const float inv_rstd = 1.0f / g.rstd;
    // 512 % c4n == 0 for every channel count of the network: a thread always meets the same four channels
    const bool fixed_c = (512 % c4n) == 0;
    float a4[4], b4[4];
    {
        const int c = (threadIdx.x % c4n) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) { a4[j] = coef[c + j]; b4[j] = coef[C + c + j]; }
    }
    float s1 = 0.f, s2 = 0.f;
    for (int i0 = threadIdx.x; i0 < total; i0 += 2 * 512) {
        float4 xv[2], dv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * 512;
            if (i < total) { xv[u] = x4[i]; dv[u] = d4[i]; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * 512;
            if (i < total) {
                if (!fixed_c) {
                    const int c = (i % c4n) * 4;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { a4[j] = coef[c + j]; b4[j] = coef[C + c + j]; }
                }
                const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float z = fmaf(xs[j], a4[j], b4[j]);
                    const float dxh = ds[j] * dsilu(z) * a4[j] * inv_rstd;
                    s1 += dxh;
                    s2 += dxh * (xs[j] - g.mean) * g.rstd;
                }
            }
        }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < 16; ++w) { a += (double)red[0][w]; c += (double)red[1][w]; }
        part[0] = a;
        part[1] = c;
    }
    if constexpr (CL > 1) cluster_barrier(); else __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        if constexpr (CL > 1) {
            for (uint32_t r = 0; r < (uint32_t)CL; ++r) { a += ld_cluster_f64(&part[0], r); c += ld_cluster_f64(&part[1], r); }
        } else { a = part[0]; c = part[1]; }
        const double cnt = (double)HW * (double)C;
        tot[0] = (float)(a / cnt) * g.rstd;
        tot[1] = (float)(c / cnt) * g.rstd;
        if (rank == 0) { sums[2 * b] = a; sums[2 * b + 1] = c; }
    }
    __syncthreads();
    const float m1 = tot[0], m2 = tot[1];
    float* o = dx + row0 * C;
    for (int i0 = threadIdx.x; i0 < total; i0 += 2 * 512) {
        float4 xv[2], dv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * 512;
            if (i < total) { xv[u] = __ldcs(x4 + i); dv[u] = __ldcs(d4 + i); }   // last use: evict first
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * 512;
            if (i < total) {
                if (!fixed_c) {
                    const int c = (i % c4n) * 4;
#pragma unroll
                    for (int j = 0; j < 4; ++j) { a4[j] = coef[c + j]; b4[j] = coef[C + c + j]; }
                }
                const float xs[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w}, ds[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
                float r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float z = fmaf(xs[j], a4[j], b4[j]);
                    r[j] = ds[j] * dsilu(z) * a4[j] - m1 - (xs[j] - g.mean) * g.rstd * m2;
                }
                store_operand4(o + 4 * (int64_t)i, make_float4(r[0], r[1], r[2], r[3]));
            }
        }
    }
    if constexpr (CL > 1) cluster_barrier();   // no CTA leaves while a partner may still read its shared memory
}

// ---------------------------------------------------------------------------------------------- channel LayerNorm backward
// forward: y = xhat * g, xhat = (x - mean) rsqrt(var + eps) per pixel row.  dx = rstd (dxhat - mean(dxhat) - xhat mean(dxhat xhat)) (+ add)
template <int NV, typename TX>
__global__ void __launch_bounds__(256) channel_layernorm_bwd_kernel(const float* __restrict__ dy, const TX* __restrict__ x,
                                                                    const float* __restrict__ g, const float* __restrict__ add,
                                                                    float* __restrict__ dx, int64_t M, int C, int operand_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= M) return;
    const int n4 = C / 4;
    float4 v[NV], d[NV];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = lane + 32 * j;
        if (i < n4) {
            v[j] = load4(x + row * C + 4 * i);
            const float4 gv = *reinterpret_cast<const float4*>(g + 4 * i);
            const float4 dv = *reinterpret_cast<const float4*>(dy + row * C + 4 * i);
            d[j] = make_float4(dv.x * gv.x, dv.y * gv.y, dv.z * gv.z, dv.w * gv.w);
        } else {
            v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            d[j] = v[j];
        }
        s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        if (lane + 32 * j < n4) {
            v[j].x -= mean; v[j].y -= mean; v[j].z -= mean; v[j].w -= mean;
            q += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
        }
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        v[j].x *= rstd; v[j].y *= rstd; v[j].z *= rstd; v[j].w *= rstd;   // xhat (zero in the padded slots)
        s1 += (d[j].x + d[j].y) + (d[j].z + d[j].w);
        s2 += (d[j].x * v[j].x + d[j].y * v[j].y) + (d[j].z * v[j].z + d[j].w * v[j].w);
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = lane + 32 * j;
        if (i < n4) {
            float4 o = make_float4(rstd * (d[j].x - s1 - v[j].x * s2), rstd * (d[j].y - s1 - v[j].y * s2),
                                   rstd * (d[j].z - s1 - v[j].z * s2), rstd * (d[j].w - s1 - v[j].w * s2));
            if (add) { const float4 a = *reinterpret_cast<const float4*>(add + row * C + 4 * i); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
            if (operand_out) store_operand4(dx + row * C + 4 * i, o); else store4(dx + row * C + 4 * i, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------- linear attention backward
constexpr int LB_HEADS = 4, LB_D = 32, LB_QKV = 3 * LB_HEADS * LB_D, LB_HID = LB_HEADS * LB_D;
constexpr int LB_CTX = LB_D * LB_D + 2 * LB_D;   // floats per (b, head) in the forward workspace: ctx | kmax | ksum
constexpr int LB_CHUNK = 128;

// dctx[b,h,d,e] = sum_n qs[d,n] dO[e,n],  qs = softmax_d(q) * 32^-0.5        grid = B*heads, 256 threads
__global__ void __launch_bounds__(256) linattn_bwd_reduce_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                                 float* __restrict__ dctx, int n) {
    __shared__ __align__(16) float qs[LB_CHUNK][LB_D];
    __shared__ __align__(16) float ds[LB_CHUNK][LB_D];
    const int b = blockIdx.x / LB_HEADS, h = blockIdx.x % LB_HEADS;
    const int tid = threadIdx.x;
    const float* qp = qkv + (int64_t)b * n * LB_QKV + h * LB_D;
    const float* dp = dout + (int64_t)b * n * LB_HID + h * LB_D;
    const int ng = tid >> 6, t64 = tid & 63;
    const int d0 = (t64 >> 3) * 4, e0 = (t64 & 7) * 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lc4 = (tid & 7) * 4, lr = tid >> 3;   // 8 consecutive lanes share one pixel
    for (int n0 = 0; n0 < n; n0 += LB_CHUNK) {
        const int cnt = min(LB_CHUNK, n - n0);
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < LB_CHUNK; rr += 32) {
            const int r = rr + lr;
            float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), dv = qv;
            const bool ok = r < cnt;
            if (ok) {
                qv = *reinterpret_cast<const float4*>(qp + (int64_t)(n0 + r) * LB_QKV + lc4);
                dv = *reinterpret_cast<const float4*>(dp + (int64_t)(n0 + r) * LB_HID + lc4);
            }
            float mx = fmaxf(fmaxf(qv.x, qv.y), fmaxf(qv.z, qv.w));
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            qv = make_float4(expf(qv.x - mx), expf(qv.y - mx), expf(qv.z - mx), expf(qv.w - mx));
            float sm = (qv.x + qv.y) + (qv.z + qv.w);
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
            const float sc = ok ? 0.17677669529663687f / sm : 0.f;
            *reinterpret_cast<float4*>(&qs[r][lc4]) = make_float4(qv.x * sc, qv.y * sc, qv.z * sc, qv.w * sc);
            *reinterpret_cast<float4*>(&ds[r][lc4]) = dv;
        }
        __syncthreads();
#pragma unroll 8
        for (int r = ng * (LB_CHUNK / 4); r < (ng + 1) * (LB_CHUNK / 4); ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&qs[r][d0]);
            const float4 vv = *reinterpret_cast<const float4*>(&ds[r][e0]);
            const float av[4] = {a.x, a.y, a.z, a.w}, vv4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], vv4[j], acc[i][j]);
        }
    }
    __syncthreads();
    float* racc = &qs[0][0];   // [4][64][16]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) racc[(ng * 64 + t64) * 16 + i * 4 + j] = acc[i][j];
    __syncthreads();
    if (ng == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int gq = 0; gq < 4; ++gq)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] += racc[(gq * 64 + t64) * 16 + i * 4 + j];
            *reinterpret_cast<float4*>(dctx + ((int64_t)blockIdx.x * LB_D + d0 + i) * LB_D + e0) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

// per pixel: dq, dk, dv -> dqkv[B*n, 384] (TF32-rounded operand of the qkv dgrad)      grid = (B*heads, ceil(n/128)), 128 threads
//   dq = p (dp - <p, dp>),  p = softmax_d(q),  dp = 32^-0.5 dO C^T          (C = ctx[d][e] of the forward)
//   dk = ks (S - cdot),     ks = exp(k - kmax) / ksum,  S = v D^T            (D = dctx[d][e], cdot[d] = <C[d], D[d]>)
//   dv = ks D
// Round-2 history: (1) one thread per pixel with three 32 x 32 matrix-vector products in registers and per-thread global accesses:
// 5.0 ms per 16x128-level call at B = 1024; (2) coalesced shared-memory staging: 4.1 ms -- the profile then showed the kernel bound by
// the broadcast LDS.128 that feed the matrix rows (one per 4-8 FMA: 512 B of register-file return traffic each).  (3) This version:
// the three products are [128 px x 32] x [32 x 32] GEMMs on mma.sync.m16n8k8 TF32 (operands rounded to nearest like every other
// tensor-core operand of the backward pass, fp32 accumulate); a warp owns 32 pixel rows (two M tiles); softmax / Jacobian algebra runs
// on the accumulator fragments (a pixel's 32 channels sit in the 4 lanes of a quad: two shuffles per reduction).  All global traffic
// goes through two shared-memory tiles with coalesced 16-byte accesses; 36-float pitch = conflict-free fragment loads.
constexpr int LB_TP = LB_D + 4;
__device__ __forceinline__ uint32_t lb_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void lb_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// acc[nt] (16 x 8 each, nt = 0..3) = A[16 x 32] * B with A = rows r0 .. r0 + 15 of `tile` and B[k][n] = m[n][k] (both pitch LB_TP)
__device__ __forceinline__ void lb_gemm(float (&acc)[4][4], const float* tile, int r0, const float* m, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = 0.f; acc[nt][1] = 0.f; acc[nt][2] = 0.f; acc[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const float* ar = tile + (r0 + g) * LB_TP + 8 * ks + t;
        const uint32_t a[4] = {lb_tf32(ar[0]), lb_tf32(ar[8 * LB_TP]), lb_tf32(ar[4]), lb_tf32(ar[8 * LB_TP + 4])};
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float* br = m + (8 * nt + g) * LB_TP + 8 * ks + t;
            lb_mma(acc[nt], a, lb_tf32(br[0]), lb_tf32(br[4]));
        }
    }
}
// tile[r][0..31] <- f(src[r * row_stride + 0..31]), r < rows (zero beyond); 128 threads: 16 rows per pass, 8 lanes per row
template <bool KS>
__device__ __forceinline__ void lb_tile_load(float* tile, const float* __restrict__ src, int64_t row_stride, int rows, int tid,
                                             const float* kmax, const float* kinv) {
    const int c4 = (tid & 7) * 4, r0 = tid >> 3;
    float m4[4] = {0.f, 0.f, 0.f, 0.f}, i4[4] = {1.f, 1.f, 1.f, 1.f};
    if constexpr (KS) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { m4[k] = kmax[c4 + k]; i4[k] = kinv[c4 + k]; }
    }
#pragma unroll
    for (int rr = 0; rr < 128; rr += 16) {
        const int r = rr + r0;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) {
            v = *reinterpret_cast<const float4*>(src + (int64_t)r * row_stride + c4);
            if constexpr (KS) v = make_float4(__expf(v.x - m4[0]) * i4[0], __expf(v.y - m4[1]) * i4[1], __expf(v.z - m4[2]) * i4[2], __expf(v.w - m4[3]) * i4[3]);
        }
        *reinterpret_cast<float4*>(tile + r * LB_TP + c4) = v;
    }
}
__device__ __forceinline__ void lb_tile_store(const float* tile, float* __restrict__ dst, int64_t row_stride, int rows, int tid) {
    const int c4 = (tid & 7) * 4, r0 = tid >> 3;
#pragma unroll
    for (int rr = 0; rr < 128; rr += 16) {
        const int r = rr + r0;
        if (r < rows) store_operand4(dst + (int64_t)r * row_stride + c4, *reinterpret_cast<const float4*>(tile + r * LB_TP + c4));
    }
}
__global__ void __launch_bounds__(128) linattn_bwd_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                                const float* __restrict__ fwd_ws, const float* __restrict__ dctx,
                                                                float* __restrict__ dqkv, int n) {
    __shared__ __align__(16) float m0[LB_D * LB_TP];   // phase A: ctx[d][e];  phase B: dctx transposed, [e][d]
    __shared__ __align__(16) float m1[LB_D * LB_TP];   // dctx[d][e]
    __shared__ float kmax[LB_D], kinv[LB_D], cdot[LB_D];
    __shared__ __align__(16) float t0[128 * LB_TP], t1[128 * LB_TP];
    const int b = blockIdx.x / LB_HEADS, h = blockIdx.x % LB_HEADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const float* ws = fwd_ws + (int64_t)blockIdx.x * LB_CTX;
    const int p0 = blockIdx.y * 128;
    const int rows = min(128, n - p0);
    const int64_t pix0 = (int64_t)b * n + p0;
    const float* qbase = qkv + pix0 * LB_QKV + h * LB_D;
    float* obase = dqkv + pix0 * LB_QKV + h * LB_D;
    lb_tile_load<false>(t0, qbase, LB_QKV, rows, tid, nullptr, nullptr);                                   // q
    lb_tile_load<false>(t1, dout + pix0 * LB_HID + h * LB_D, LB_HID, rows, tid, nullptr, nullptr);         // dO
    for (int i = tid; i < LB_D * LB_D; i += 128) {
        m0[(i >> 5) * LB_TP + (i & 31)] = ws[i];
        m1[(i >> 5) * LB_TP + (i & 31)] = dctx[(int64_t)blockIdx.x * LB_D * LB_D + i];
    }
    if (tid < LB_D) {
        kmax[tid] = ws[LB_D * LB_D + tid];
        kinv[tid] = 1.0f / ws[LB_D * LB_D + LB_D + tid];
    }
    __syncthreads();
    if (tid < LB_D) {
        float a = 0.f;
        for (int e = 0; e < LB_D; ++e) a = fmaf(m1[tid * LB_TP + e], m0[tid * LB_TP + e], a);
        cdot[tid] = a;   // = sum_n ks[d,n] dks[d,n]
    }
    // ---- phase A: dq (this warp's 32 rows, two M tiles) ----
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int r0 = warp * 32 + mt * 16;
        float dp[4][4];
        lb_gemm(dp, t1, r0, m0, lane);   // dO C^T: columns = d
        float* q0 = t0 + (r0 + g) * LB_TP + 2 * t;   // accumulator layout: rows r0 + g, r0 + g + 8; columns 8 nt + 2 t, + 1
        float* q1 = q0 + 8 * LB_TP;
        float p[4][4];
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float2 u = *reinterpret_cast<const float2*>(q0 + 8 * nt), w = *reinterpret_cast<const float2*>(q1 + 8 * nt);
            p[nt][0] = u.x; p[nt][1] = u.y; p[nt][2] = w.x; p[nt][3] = w.y;
            mx0 = fmaxf(mx0, fmaxf(u.x, u.y));
            mx1 = fmaxf(mx1, fmaxf(w.x, w.y));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        float d0 = 0.f, d1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            p[nt][0] = __expf(p[nt][0] - mx0); p[nt][1] = __expf(p[nt][1] - mx0);
            p[nt][2] = __expf(p[nt][2] - mx1); p[nt][3] = __expf(p[nt][3] - mx1);
            d0 += p[nt][0] + p[nt][1];
            d1 += p[nt][2] + p[nt][3];
        }
        d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
        d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
        const float i0 = 1.0f / d0, i1 = 1.0f / d1;
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            p[nt][0] *= i0; p[nt][1] *= i0; p[nt][2] *= i1; p[nt][3] *= i1;
#pragma unroll
            for (int k = 0; k < 4; ++k) dp[nt][k] *= 0.17677669529663687f;
            s0 += p[nt][0] * dp[nt][0] + p[nt][1] * dp[nt][1];
            s1 += p[nt][2] * dp[nt][2] + p[nt][3] * dp[nt][3];
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {   // own rows, own elements: each q element was read by exactly this lane
            *reinterpret_cast<float2*>(q0 + 8 * nt) = make_float2(p[nt][0] * (dp[nt][0] - s0), p[nt][1] * (dp[nt][1] - s0));
            *reinterpret_cast<float2*>(q1 + 8 * nt) = make_float2(p[nt][2] * (dp[nt][2] - s1), p[nt][3] * (dp[nt][3] - s1));
        }
    }
    __syncthreads();   // dq rows complete, dO rows and ctx consumed, cdot visible
    lb_tile_store(t0, obase, LB_QKV, rows, tid);                                           // dq out
    lb_tile_load<false>(t0, qbase + 2 * LB_HID, LB_QKV, rows, tid, nullptr, nullptr);       // v  (same thread <-> element mapping as the store above)
    lb_tile_load<true>(t1, qbase + LB_HID, LB_QKV, rows, tid, kmax, kinv);                  // ks = exp(k - kmax) / ksum
    for (int i = tid; i < LB_D * LB_D; i += 128) m0[(i & 31) * LB_TP + (i >> 5)] = m1[(i >> 5) * LB_TP + (i & 31)];   // m0[e][d] = dctx[d][e]
    __syncthreads();
    // ---- phase B: dk, dv ----
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int r0 = warp * 32 + mt * 16;
        float sv[4][4], dv[4][4];
        lb_gemm(sv, t0, r0, m1, lane);   // v D^T: columns = d
        lb_gemm(dv, t1, r0, m0, lane);   // ks D:  columns = e
        float* k0 = t1 + (r0 + g) * LB_TP + 2 * t;
        float* k1 = k0 + 8 * LB_TP;
        float* v0 = t0 + (r0 + g) * LB_TP + 2 * t;
        float* v1 = v0 + 8 * LB_TP;
        float dk[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const float2 u = *reinterpret_cast<const float2*>(k0 + 8 * nt), w = *reinterpret_cast<const float2*>(k1 + 8 * nt);
            const float c0 = cdot[8 * nt + 2 * t], c1 = cdot[8 * nt + 2 * t + 1];
            dk[nt][0] = u.x * (sv[nt][0] - c0); dk[nt][1] = u.y * (sv[nt][1] - c1);
            dk[nt][2] = w.x * (sv[nt][2] - c0); dk[nt][3] = w.y * (sv[nt][3] - c1);
        }
        __syncwarp();   // every lane's fragment reads of these 16 rows of ks / v are done before they are overwritten
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<float2*>(k0 + 8 * nt) = make_float2(dk[nt][0], dk[nt][1]);
            *reinterpret_cast<float2*>(k1 + 8 * nt) = make_float2(dk[nt][2], dk[nt][3]);
            *reinterpret_cast<float2*>(v0 + 8 * nt) = make_float2(dv[nt][0], dv[nt][1]);
            *reinterpret_cast<float2*>(v1 + 8 * nt) = make_float2(dv[nt][2], dv[nt][3]);
        }
    }
    __syncthreads();
    lb_tile_store(t1, obase + LB_HID, LB_QKV, rows, tid);
    lb_tile_store(t0, obase + 2 * LB_HID, LB_QKV, rows, tid);
}

// ---------------------------------------------------------------------------------------------- full attention backward (n <= 32)
// one warp per (b, head); lane = query i for the row quantities, lane = key j for dk / dv
__global__ void __launch_bounds__(32) attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                           float* __restrict__ dqkv, int n) {
    __shared__ float qs[1][32][LB_D + 1], ks[1][32][LB_D + 1], vs[1][32][LB_D + 1], gs[1][32][LB_D + 1];
    __shared__ float Ps[1][32][33], Ss[1][32][33];
    const int b = blockIdx.x / LB_HEADS, hh = blockIdx.x % LB_HEADS, lane = threadIdx.x;
    constexpr int h = 0;   // one warp (= one head of one sample) per CTA
    const float* base = qkv + (int64_t)b * n * LB_QKV;
    const float scale = 0.17677669529663687f;
    for (int j = 0; j < 32; ++j) {
        const bool ok = j < n;
        qs[h][j][lane] = ok ? base[(int64_t)j * LB_QKV + hh * LB_D + lane] : 0.f;
        ks[h][j][lane] = ok ? base[(int64_t)j * LB_QKV + LB_HID + hh * LB_D + lane] : 0.f;
        vs[h][j][lane] = ok ? base[(int64_t)j * LB_QKV + 2 * LB_HID + hh * LB_D + lane] : 0.f;
        gs[h][j][lane] = ok ? dout[((int64_t)b * n + j) * LB_HID + hh * LB_D + lane] : 0.f;
    }
    __syncwarp();
    {   // row i = lane: P[i,:], dS[i,:], dq_i
        const int i = lane;
        float sim[32];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float s = -INFINITY;
            if (j < n) {
                s = 0.f;
#pragma unroll
                for (int d = 0; d < LB_D; ++d) s = fmaf(qs[h][i][d] * scale, ks[h][j][d], s);
            }
            sim[j] = s;
            mx = fmaxf(mx, s);
        }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { sim[j] = (j < n) ? expf(sim[j] - mx) : 0.f; den += sim[j]; }
        const float inv = 1.0f / den;
        float dot = 0.f;
        float dP[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            sim[j] *= inv;
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < LB_D; ++d) s = fmaf(gs[h][i][d], vs[h][j][d], s);
            dP[j] = s;
            dot = fmaf(sim[j], s, dot);
        }
        float dq[LB_D];
#pragma unroll
        for (int d = 0; d < LB_D; ++d) dq[d] = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float dS = sim[j] * (dP[j] - dot);
            Ps[h][i][j] = sim[j];
            Ss[h][i][j] = dS;
#pragma unroll
            for (int d = 0; d < LB_D; ++d) dq[d] = fmaf(dS * scale, ks[h][j][d], dq[d]);
        }
        if (i < n) {
            float* o = dqkv + ((int64_t)b * n + i) * LB_QKV + hh * LB_D;
#pragma unroll
            for (int d = 0; d < LB_D; ++d) o[d] = to_tf32(dq[d]);
        }
    }
    __syncwarp();
    {   // key j = lane: dk_j = scale sum_i dS[i,j] q_i,  dv_j = sum_i P[i,j] dO_i
        const int j = lane;
        float dk[LB_D], dv[LB_D];
#pragma unroll
        for (int d = 0; d < LB_D; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
        for (int i = 0; i < n; ++i) {
            const float dS = Ss[h][i][j] * scale, P = Ps[h][i][j];
#pragma unroll
            for (int d = 0; d < LB_D; ++d) {
                dk[d] = fmaf(dS, qs[h][i][d], dk[d]);
                dv[d] = fmaf(P, gs[h][i][d], dv[d]);
            }
        }
        if (j < n) {
            float* o = dqkv + ((int64_t)b * n + j) * LB_QKV + hh * LB_D;
#pragma unroll
            for (int d = 0; d < LB_D; ++d) { o[LB_HID + d] = to_tf32(dk[d]); o[2 * LB_HID + d] = to_tf32(dv[d]); }
        }
    }
}

// ---------------------------------------------------------------------------------------------- layout / small ops
// dx[b, 2h+p1, 2w+p2, c] = t[b, h, w, (2 p1 + p2) C + c] (+ add)          (backward of the pixel-unshuffle view)
__global__ void pixel_shuffle_bwd_kernel(const float4* __restrict__ t, const float4* __restrict__ add, float4* __restrict__ dx,
                                         int64_t total4, int H, int W, int c4, int operand_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        int64_t pix = i / c4;
        const int wo = (int)(pix % (2 * W)); pix /= (2 * W);
        const int ho = (int)(pix % (2 * H));
        const int64_t b = pix / (2 * H);
        const int p = (ho & 1) * 2 + (wo & 1);
        float4 v = t[(((b * H + (ho >> 1)) * W + (wo >> 1)) * 4 + p) * c4 + c];
        if (add) { const float4 a = add[i]; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
        if (operand_out) v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        dx[i] = v;
    }
}

// dx[b, h, w, c] = sum of the 2x2 block of dy[b, 2h.., 2w.., c]                (backward of nearest x2 upsample)
__global__ void upsample2x_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, int64_t total4, int H, int W, int c4,
                                      int operand_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        int64_t pix = i / c4;
        const int w = (int)(pix % W); pix /= W;
        const int h = (int)(pix % H);
        const int64_t b = pix / H;
        const int64_t r0 = ((b * 2 * H + 2 * h) * 2 * W + 2 * w) * c4 + c;
        const int64_t r1 = r0 + (int64_t)2 * W * c4;
        const float4 a = dy[r0], bq = dy[r0 + c4], cq = dy[r1], d = dy[r1 + c4];
        float4 v = make_float4((a.x + bq.x) + (cq.x + d.x), (a.y + bq.y) + (cq.y + d.y), (a.z + bq.z) + (cq.z + d.z), (a.w + bq.w) + (cq.w + d.w));
        if (operand_out) v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        dx[i] = v;
    }
}

// a += b (optionally rounding the sum to TF32)
__global__ void add_inplace_kernel(float4* __restrict__ a, const float4* __restrict__ b, int64_t n4, int operand_out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = a[i];
        const float4 u = b[i];
        v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
        if (operand_out) v = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
        a[i] = v;
    }
}

// head 1x1 conv backward: dx[m, c] = sum_o g[b, o, p] w[o, c]       (g NCHW fp32, dx NHWC fp32)
__global__ void __launch_bounds__(256) head_conv1_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                             float* __restrict__ dx, int64_t M, int HW, int Cin, int Cout, int operand_out) {
    const int c4n = Cin / 4;
    const int64_t total = M * c4n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        const int64_t m = i / c4n, b = m / HW, p = m % HW;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int o = 0; o < Cout; ++o) {
            const float gv = g[(b * Cout + o) * HW + p];
            const float4 wv = *reinterpret_cast<const float4*>(w + (int64_t)o * Cin + c);
            acc.x = fmaf(gv, wv.x, acc.x); acc.y = fmaf(gv, wv.y, acc.y); acc.z = fmaf(gv, wv.z, acc.z); acc.w = fmaf(gv, wv.w, acc.w);
        }
        if (operand_out) store_operand4(dx + m * Cin + c, acc); else store4(dx + m * Cin + c, acc);
    }
}

// stem 7x7 backward, col2im half: dx[b, ci, h, w] = sum_{ky,kx} t[(b, h+3-ky, w+3-kx), ci*49 + ky*7 + kx]   (t: [B*H*W, ld])
__global__ void __launch_bounds__(256) stem_col2im_kernel(const float* __restrict__ t, float* __restrict__ dx, int B, int Cin, int H,
                                                          int W, int ld) {
    const int64_t total = (int64_t)B * Cin * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(i % W);
        int64_t r = i / W;
        const int h = (int)(r % H); r /= H;
        const int ci = (int)(r % Cin);
        const int64_t b = r / Cin;
        float acc = 0.f;
        for (int ky = 0; ky < 7; ++ky) {
            const int hh = h + 3 - ky;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const int ww = w + 3 - kx;
                if (ww < 0 || ww >= W) continue;
                acc += t[((b * H + hh) * W + ww) * ld + ci * 49 + ky * 7 + kx];
            }
        }
        dx[i] = acc;
    }
}

}  // namespace sdc

using namespace sdc;

static inline unsigned grid_for(int64_t n, int per) { int64_t b = (n + per - 1) / per; return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

extern "C" int sdc_pack_conv_weight_dgrad(int kind, const float* w, float* wt, int Cout, int Cin, void* stream) {
    SDC_REQUIRE(kind >= 0 && kind <= 2 && w && wt && Cout > 0 && Cin > 0, "pack_conv_weight_dgrad: bad arguments");
    SDC_REQUIRE(kind != 2 || Cin % 4 == 0, "pack_conv_weight_dgrad: unshuffle conv needs Cin %% 4 == 0");
    const int64_t total = (int64_t)Cout * Cin * (kind == 1 ? 9 : 1);
    pack_dgrad_weight_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(kind, w, wt, Cout, Cin);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_gn_silu_bwd(const float* dy, const float* x, const double* stats, const float* gamma, const float* beta,
                               const float* scale_shift, const int32_t* t_index, int64_t ss_stride, double* sums, float* dx, int B,
                               int HW, int C, void* stream) {
    SDC_REQUIRE(dy && x && stats && gamma && beta && sums && dx && B > 0 && HW > 0, "gn_silu_bwd: bad arguments");
    SDC_REQUIRE(C % 4 == 0 && C <= 4096, "gn_silu_bwd: C=%d unsupported", C);
    const size_t sm = 2 * C * sizeof(float);
    cudaStream_t st = as_stream(stream);
    static const bool fused = []() { const char* e = getenv("SDC_GN_BWD_FUSED"); return !(e && e[0] == '0'); }();
    if (fused) {
        // one kernel: a cluster of 16 / 8 CTAs per sample when a sample is large enough to feed them (>= 256 K / 128 K elements: at most
        // 18 x 2 MB / 37 x 1 MB of dy + x in flight between the passes), else one CTA per sample
        static const int max_cl = []() { const char* e = getenv("SDC_GN_BWD_CL"); return e ? atoi(e) : 8; }();   // 16 (non-portable) measured slower: 1.38 vs 1.10 ms per 16x128-level call
        const int64_t ne = (int64_t)HW * C;
        const int cl = (ne >= 262144 && HW % 16 == 0 && max_cl >= 16) ? 16 : ((ne >= 131072 && HW % 8 == 0 && max_cl >= 8) ? 8 : 1);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(B * cl));
        cfg.blockDim = dim3(512);
        cfg.dynamicSmemBytes = sm;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cl;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        static bool np_set = false;
        if (!np_set) {
            SDC_CUDA(cudaFuncSetAttribute(gn_silu_bwd_fused_kernel<16>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            np_set = true;
        }
        if (cl == 16) SDC_CUDA(cudaLaunchKernelEx(&cfg, gn_silu_bwd_fused_kernel<16>, dy, x, stats, gamma, beta, scale_shift, t_index, ss_stride, sums, dx, HW, C));
        else if (cl == 8) SDC_CUDA(cudaLaunchKernelEx(&cfg, gn_silu_bwd_fused_kernel<8>, dy, x, stats, gamma, beta, scale_shift, t_index, ss_stride, sums, dx, HW, C));
        else SDC_CUDA(cudaLaunchKernelEx(&cfg, gn_silu_bwd_fused_kernel<1>, dy, x, stats, gamma, beta, scale_shift, t_index, ss_stride, sums, dx, HW, C));
        SDC_LAUNCHED();
        return SDC_OK;
    }
    int ppc = HW;
    while (ppc > 32 && (int64_t)B * (HW / ppc) < 2 * 148 && ppc % 2 == 0) ppc /= 2;
    dim3 grid((unsigned)B, (unsigned)((HW + ppc - 1) / ppc));
    SDC_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * 2 * sizeof(double), st));
    gn_silu_bwd_reduce_kernel<<<grid, 256, sm, st>>>(dy, x, stats, gamma, beta, scale_shift, t_index, ss_stride, sums, HW, C, ppc);
    SDC_LAUNCHED();
    gn_silu_bwd_apply_kernel<<<grid, 256, sm, st>>>(dy, x, stats, gamma, beta, scale_shift, t_index, ss_stride, sums, dx, HW, C, ppc);
    SDC_LAUNCHED();
    return SDC_OK;
}

template <typename TX>
static void launch_ln_bwd(const float* dy, const void* x, const float* g, const float* add, float* dx, int64_t M, int C,
                          int operand_out, cudaStream_t st) {
    const TX* xp = (const TX*)x;
    const unsigned grid = (unsigned)((M + 7) / 8);
    if (C <= 128) channel_layernorm_bwd_kernel<1, TX><<<grid, 256, 0, st>>>(dy, xp, g, add, dx, M, C, operand_out);
    else if (C <= 256) channel_layernorm_bwd_kernel<2, TX><<<grid, 256, 0, st>>>(dy, xp, g, add, dx, M, C, operand_out);
    else if (C <= 512) channel_layernorm_bwd_kernel<4, TX><<<grid, 256, 0, st>>>(dy, xp, g, add, dx, M, C, operand_out);
    else channel_layernorm_bwd_kernel<8, TX><<<grid, 256, 0, st>>>(dy, xp, g, add, dx, M, C, operand_out);
}

extern "C" int sdc_channel_layernorm_bwd(const float* dy, const void* x, int x_half, const float* g, const float* add, float* dx,
                                         int64_t M, int C, int operand_out, void* stream) {
    SDC_REQUIRE(dy && x && g && dx && M > 0, "channel_layernorm_bwd: bad arguments");
    SDC_REQUIRE(C % 4 == 0 && C <= 1024, "channel_layernorm_bwd: C=%d unsupported (multiple of 4, <= 1024)", C);
    if (x_half) launch_ln_bwd<__half>(dy, x, g, add, dx, M, C, operand_out, as_stream(stream));
    else launch_ln_bwd<float>(dy, x, g, add, dx, M, C, operand_out, as_stream(stream));
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int64_t sdc_linear_attention_bwd_workspace(int B) { return (int64_t)B * LB_HEADS * LB_D * LB_D * sizeof(float); }

extern "C" int sdc_linear_attention_bwd(const float* qkv, const float* dout, const void* fwd_workspace, void* workspace, float* dqkv,
                                        int B, int n, void* stream) {
    SDC_REQUIRE(qkv && dout && fwd_workspace && workspace && dqkv && B > 0 && n > 0, "linear_attention_bwd: bad arguments");
    float* dctx = reinterpret_cast<float*>(workspace);
    linattn_bwd_reduce_kernel<<<(unsigned)(B * LB_HEADS), 256, 0, as_stream(stream)>>>(qkv, dout, dctx, n);
    SDC_LAUNCHED();
    dim3 grid((unsigned)(B * LB_HEADS), (unsigned)((n + 127) / 128));
    linattn_bwd_apply_kernel<<<grid, 128, 0, as_stream(stream)>>>(qkv, dout, reinterpret_cast<const float*>(fwd_workspace), dctx, dqkv, n);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_attention_bwd(const float* qkv, const float* dout, float* dqkv, int B, int n, void* stream) {
    SDC_REQUIRE(qkv && dout && dqkv && B > 0, "attention_bwd: bad arguments");
    SDC_REQUIRE(n > 0 && n <= 32, "attention_bwd: n=%d tokens unsupported", n);
    attention_bwd_kernel<<<(unsigned)(B * LB_HEADS), 32, 0, as_stream(stream)>>>(qkv, dout, dqkv, n);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_pixel_shuffle_bwd(const float* t, const float* add, float* dx, int B, int H, int W, int C, int operand_out,
                                     void* stream) {
    SDC_REQUIRE(t && dx && B > 0 && C % 4 == 0, "pixel_shuffle_bwd: bad arguments");
    const int64_t total4 = (int64_t)B * 4 * H * W * (C / 4);
    pixel_shuffle_bwd_kernel<<<grid_for(total4, 256 * 4), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(t), reinterpret_cast<const float4*>(add), reinterpret_cast<float4*>(dx), total4, H, W, C / 4,
        operand_out);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_upsample2x_bwd(const float* dy, float* dx, int B, int H, int W, int C, int operand_out, void* stream) {
    SDC_REQUIRE(dy && dx && B > 0 && C % 4 == 0, "upsample2x_bwd: bad arguments");
    const int64_t total4 = (int64_t)B * H * W * (C / 4);
    upsample2x_bwd_kernel<<<grid_for(total4, 256 * 2), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(dy),
                                                                                  reinterpret_cast<float4*>(dx), total4, H, W, C / 4, operand_out);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_add_inplace(float* a, const float* b, int64_t n, int operand_out, void* stream) {
    SDC_REQUIRE(a && b && n > 0 && n % 4 == 0, "add_inplace: bad arguments");
    add_inplace_kernel<<<grid_for(n / 4, 256 * 4), 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(a),
                                                                                reinterpret_cast<const float4*>(b), n / 4, operand_out);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_head_conv1_bwd(const float* g, const float* w, float* dx, int B, int HW, int Cin, int Cout, int operand_out,
                                  void* stream) {
    SDC_REQUIRE(g && w && dx && B > 0 && Cin % 4 == 0 && Cout >= 1, "head_conv1_bwd: bad arguments");
    const int64_t M = (int64_t)B * HW;
    head_conv1_bwd_kernel<<<grid_for(M * (Cin / 4), 256), 256, 0, as_stream(stream)>>>(g, w, dx, M, HW, Cin, Cout, operand_out);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_stem_col2im(const float* t, float* dx, int B, int Cin, int H, int W, int ld, void* stream) {
    SDC_REQUIRE(t && dx && B > 0 && ld >= Cin * 49, "stem_col2im: bad arguments");
    const int64_t total = (int64_t)B * Cin * H * W;
    stem_col2im_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(t, dx, B, Cin, H, W, ld);
    SDC_LAUNCHED();
    return SDC_OK;
}
