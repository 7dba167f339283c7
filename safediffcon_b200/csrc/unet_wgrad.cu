// Parameter-gradient kernels of the denoiser: what autograd produces for the U-Net
// parameters when the reference back-propagates the inference-time fine-tuning loss through the LAST DDIM step
// (/root/reference/1D/model/diffusion.py:524-551 under enable_grad, /root/reference/1D/inference/inference_ft.py:189-226),
// and what the post-training diffusion loss needs (/root/reference/1D/posttrain/post_train.py:206-260).
//
// The activation gradients come from the backward-data pass (unet_bwd.cu + the tcgen05 dgrad convolutions); this file
// reduces them against the saved forward activations:
//   conv weight   dW[co, ci, ky, kx] = sum_p dY[p, co] A[p + (ky, kx), ci]      implicit GEMM, reduction over pixels
//   conv bias     db[co]             = sum_p dY[p, co]
//   GroupNorm     per (sample, channel) sums of dz and dz*xhat -> d gamma, d beta and the FiLM (scale, shift) gradients
//   LayerNorm     dg[c]              = sum_p dY[p, c] xhat[p, c]
//   head 1x1      dW[o, c], db[o]
// The weight gradient runs once per chain (one of 200 denoiser evaluations) on the fine-tuning batch (50 samples in the
// shipped configs), so it is written for generality over the U-Net's conv flavours rather than for peak: mma.sync
// m16n8k8 TF32 (operands rounded to nearest, fp32 accumulate -- the precision of the dgrad convolutions), pixel tiles
// staged through shared memory, split over the pixel axis with fp32 atomics into the OIHW gradient tensor.
#include "common.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <math.h>

namespace sdc {

__device__ __forceinline__ float wg_sigmoid(float v) { return __fdividef(1.0f, 1.0f + __expf(-v)); }   // MUFU.EX2 + MUFU.RCP, ~1e-6 (see unet_bwd.cu)
__device__ __forceinline__ float wg_dsilu(float z) {
    const float s = wg_sigmoid(z);
    return s * (1.0f + z * (1.0f - s));
}
__device__ __forceinline__ uint32_t tf32_bits(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void wg_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------- conv weight gradient
struct WgradArgs {
    const void* a0; const void* a1; const float* dy; float* dw;
    int c0, c1, kind, B, H, W, Cout;      // H, W: OUTPUT resolution (kind 2: the input is 2H x 2W)
    int chunks_per_split, n_chunks;
    int64_t M;
};
constexpr int WG_PX = 32, WG_T = 64, WG_LD = WG_T + 8;   // pixel chunk, tile edge (co and ci), padded smem stride

// One CTA: 64 output channels x 64 input channels of ONE tap, a contiguous range of 32-pixel chunks.  4 warps, warp tile
// 32 x 32 = 2 (m16) x 4 (n8) mma tiles.  A operand = dY^T (rows co, K = pixels), B operand = shifted activations.
template <typename TA>
__global__ void __launch_bounds__(128) conv_wgrad_kernel(WgradArgs p) {
    __shared__ __align__(16) uint32_t dyS[WG_PX][WG_LD];
    __shared__ __align__(16) uint32_t aS[WG_PX][WG_LD];
    const int Cin = p.c0 + p.c1;
    const int taps = p.kind == 1 ? 9 : (p.kind == 2 ? 4 : 1);
    const int tap = blockIdx.x % taps, ci0 = (blockIdx.x / taps) * WG_T, co0 = blockIdx.y * WG_T;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wco = (warp >> 1) * 32, wci = (warp & 1) * 32;
    int dh = 0, dw_ = 0, Hin = p.H, Win = p.W, mul = 1;
    if (p.kind == 1) { dh = tap / 3 - 1; dw_ = tap % 3 - 1; }
    if (p.kind == 2) { dh = tap >> 1; dw_ = tap & 1; Hin = 2 * p.H; Win = 2 * p.W; mul = 2; }
    constexpr int APV = 16 / (int)sizeof(TA);           // activation elements per 16-byte piece (8 halves / 4 floats)
    constexpr int APIECES = WG_T / APV;                 // pieces per pixel row of the activation tile
    constexpr int ALOADS = WG_PX * APIECES / 128;       // pieces per thread
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    const int chunk_lo = blockIdx.z * p.chunks_per_split;
    const int chunk_hi = min(chunk_lo + p.chunks_per_split, p.n_chunks);
    float4 dreg[4];
    uint4 areg[ALOADS];
    auto fetch = [&](int chunk) {
        const int64_t m0 = (int64_t)chunk * WG_PX;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = tid + 128 * j, px = q >> 4, co = co0 + (q & 15) * 4;
            const int64_t m = m0 + px;
            dreg[j] = (m < p.M && co < p.Cout) ? *reinterpret_cast<const float4*>(p.dy + m * p.Cout + co) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < ALOADS; ++j) {
            const int q = tid + 128 * j, px = q / APIECES, ci = ci0 + (q % APIECES) * APV;
            const int64_t m = m0 + px;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (m < p.M && ci < Cin) {
                const int w = (int)(m % p.W), h = (int)((m / p.W) % p.H);
                const int64_t b = m / ((int64_t)p.W * p.H);
                const int hs = h * mul + dh, ws = w * mul + dw_;
                if (hs >= 0 && hs < Hin && ws >= 0 && ws < Win) {
                    const int64_t src = (b * Hin + hs) * Win + ws;
                    const TA* base = ci < p.c0 ? reinterpret_cast<const TA*>(p.a0) + src * p.c0 + ci
                                               : reinterpret_cast<const TA*>(p.a1) + src * p.c1 + (ci - p.c0);
                    v = *reinterpret_cast<const uint4*>(base);
                }
            }
            areg[j] = v;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int q = tid + 128 * j, px = q >> 4, c = (q & 15) * 4;
            *reinterpret_cast<uint4*>(&dyS[px][c]) = make_uint4(tf32_bits(dreg[j].x), tf32_bits(dreg[j].y), tf32_bits(dreg[j].z), tf32_bits(dreg[j].w));
        }
#pragma unroll
        for (int j = 0; j < ALOADS; ++j) {
            const int q = tid + 128 * j, px = q / APIECES, c = (q % APIECES) * APV;
            if constexpr (sizeof(TA) == 2) {
                const __half2* hp = reinterpret_cast<const __half2*>(&areg[j]);
                const float2 f0 = __half22float2(hp[0]), f1 = __half22float2(hp[1]), f2 = __half22float2(hp[2]), f3 = __half22float2(hp[3]);
                // fp16 -> fp32 is exact and already has <= 10 mantissa bits: valid TF32 bit patterns
                *reinterpret_cast<uint4*>(&aS[px][c]) = make_uint4(__float_as_uint(f0.x), __float_as_uint(f0.y), __float_as_uint(f1.x), __float_as_uint(f1.y));
                *reinterpret_cast<uint4*>(&aS[px][c + 4]) = make_uint4(__float_as_uint(f2.x), __float_as_uint(f2.y), __float_as_uint(f3.x), __float_as_uint(f3.y));
            } else {
                *reinterpret_cast<uint4*>(&aS[px][c]) = make_uint4(tf32_bits(__uint_as_float(areg[j].x)), tf32_bits(__uint_as_float(areg[j].y)),
                                                                   tf32_bits(__uint_as_float(areg[j].z)), tf32_bits(__uint_as_float(areg[j].w)));
            }
        }
    };
    if (chunk_lo < chunk_hi) fetch(chunk_lo);
    for (int chunk = chunk_lo; chunk < chunk_hi; ++chunk) {
        __syncthreads();          // previous chunk's fragments are consumed
        stage();
        __syncthreads();
        if (chunk + 1 < chunk_hi) fetch(chunk + 1);   // global loads in flight while the tensor cores work
#pragma unroll
        for (int k0 = 0; k0 < WG_PX; k0 += 8) {
            uint32_t af[2][4], bf[4][2];
#pragma unroll
            for (int mi = 0; mi < 2; ++mi) {
                const int r = wco + mi * 16 + g;
                af[mi][0] = dyS[k0 + t][r]; af[mi][1] = dyS[k0 + t][r + 8];
                af[mi][2] = dyS[k0 + t + 4][r]; af[mi][3] = dyS[k0 + t + 4][r + 8];
            }
#pragma unroll
            for (int ni = 0; ni < 4; ++ni) {
                const int c = wci + ni * 8 + g;
                bf[ni][0] = aS[k0 + t][c]; bf[ni][1] = aS[k0 + t + 4][c];
            }
#pragma unroll
            for (int mi = 0; mi < 2; ++mi)
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) wg_mma(acc[mi][ni], af[mi], bf[ni][0], bf[ni][1]);
        }
    }
    // dW index: kind 1 -> ((co*Cin + ci)*9 + tap); kind 0 -> co*Cin + ci; kind 2 -> co*4Cin + ci*4 + tap
    const int64_t s_co = (int64_t)Cin * taps, s_ci = taps;
#pragma unroll
    for (int mi = 0; mi < 2; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int co = co0 + wco + mi * 16 + g + (e >> 1) * 8;
                const int ci = ci0 + wci + ni * 8 + 2 * t + (e & 1);
                if (co < p.Cout && ci < Cin) atomicAdd(p.dw + co * s_co + ci * s_ci + tap, acc[mi][ni][e]);
            }
}

// ---------------------------------------------------------------------------------------------- column sums (conv bias)
// out[c] += sum_rows x[row, c]; thread groups of C/4 lanes walk disjoint rows, partial sums meet in shared memory
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t M, int C,
                                                     int rows_per_cta) {
    extern __shared__ float part[];   // [C]
    const int c4n = C / 4;
    for (int c = threadIdx.x; c < C; c += blockDim.x) part[c] = 0.f;
    __syncthreads();
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(r0 + (int64_t)rows_per_cta, M);
    if (c4n <= 256) {
        const int groups = 256 / c4n, grp = threadIdx.x / c4n, q = threadIdx.x % c4n;
        if (grp < groups) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int64_t r = r0 + grp; r < r1; r += groups) {
                const float4 v = *reinterpret_cast<const float4*>(x + r * C + 4 * q);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            atomicAdd(&part[4 * q], s.x); atomicAdd(&part[4 * q + 1], s.y); atomicAdd(&part[4 * q + 2], s.z); atomicAdd(&part[4 * q + 3], s.w);
        }
    } else {
        for (int q = threadIdx.x; q < c4n; q += blockDim.x) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int64_t r = r0; r < r1; ++r) {
                const float4 v = *reinterpret_cast<const float4*>(x + r * C + 4 * q);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            part[4 * q] = s.x; part[4 * q + 1] = s.y; part[4 * q + 2] = s.z; part[4 * q + 3] = s.w;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(out + c, part[c]);
}

// ---------------------------------------------------------------------------------------------- GroupNorm(1)+FiLM+SiLU parameters
// forward: z = (xhat gamma + beta)(1 + sc) + sh, y = silu(z).  With dz = dy silu'(z):
//   P0[b, c] = sum_p dz,  P1[b, c] = sum_p dz xhat      (this kernel; the caller combines them:
//   d gamma_c = sum_b (1+sc_bc) P1, d beta_c = sum_b (1+sc_bc) P0, d sc_bc = gamma_c P1 + beta_c P0, d sh_bc = P0)
__global__ void __launch_bounds__(256) gn_param_grad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const double* __restrict__ stats, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, const float* __restrict__ scale_shift,
                                                            const int32_t* __restrict__ t_index, int64_t ss_stride,
                                                            float* __restrict__ P, int HW, int C, int pix_per_cta) {
    extern __shared__ float sm[];     // coef A[C], B[C], part0[C], part1[C]
    float* coef = sm;
    float* part = sm + 2 * C;
    const int b = blockIdx.x;
    const double cnt = (double)HW * (double)C;
    const double mean_d = stats[2 * b] / cnt;
    double var_d = stats[2 * b + 1] / cnt - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + 1e-5));
    const float* ss = scale_shift ? scale_shift + (int64_t)(t_index ? t_index[b] : 0) * ss_stride : nullptr;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float gm = gamma[c] * rstd;
        float a = gm, bb = beta[c] - mean * gm;
        if (ss) {
            const float sc = ss[c] + 1.0f;
            a *= sc;
            bb = bb * sc + ss[C + c];
        }
        coef[c] = a;
        coef[C + c] = bb;
        part[c] = 0.f;
        part[C + c] = 0.f;
    }
    __syncthreads();
    const int c4n = C / 4;
    const int64_t row0 = (int64_t)b * HW + (int64_t)blockIdx.y * pix_per_cta;
    const int rows = min(pix_per_cta, HW - (int)blockIdx.y * pix_per_cta);
    auto accumulate = [&](int q, int r_first, int r_step) {
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = r_first; r < rows; r += r_step) {
            const float4 xv = *reinterpret_cast<const float4*>(x + (row0 + r) * C + 4 * q);
            const float4 dv = *reinterpret_cast<const float4*>(dy + (row0 + r) * C + 4 * q);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float z = fmaf(xs[j], coef[4 * q + j], coef[C + 4 * q + j]);
                const float dz = ds[j] * wg_dsilu(z);
                s0[j] += dz;
                s1[j] += dz * (xs[j] - mean) * rstd;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) { atomicAdd(&part[4 * q + j], s0[j]); atomicAdd(&part[C + 4 * q + j], s1[j]); }
    };
    if (c4n <= 256) {
        const int groups = 256 / c4n, grp = threadIdx.x / c4n;
        if (grp < groups) accumulate(threadIdx.x % c4n, grp, groups);
    } else {
        for (int q = threadIdx.x; q < c4n; q += 256) accumulate(q, 0, 1);
    }
    __syncthreads();
    float* Pb = P + (int64_t)b * 2 * C;
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) atomicAdd(Pb + c, part[c]);
}

// ---------------------------------------------------------------------------------------------- channel LayerNorm gain
// dg[c] += sum_rows dy[row, c] xhat[row, c]; one warp per row (grid-stride), per-lane partials over the rows it meets
template <int NV, typename TX>
__global__ void __launch_bounds__(256) ln_gain_grad_kernel(const float* __restrict__ dy, const TX* __restrict__ x,
                                                           float* __restrict__ dg, int64_t M, int C) {
    extern __shared__ float part[];   // [C]
    for (int c = threadIdx.x; c < C; c += blockDim.x) part[c] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int n4 = C / 4;
    float4 acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < M; row += wstride) {
        float4 v[NV];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            v[j] = i < n4 ? load4(x + row * C + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
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
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < n4) {
                const float4 d = *reinterpret_cast<const float4*>(dy + row * C + 4 * i);
                acc[j].x += d.x * v[j].x * rstd; acc[j].y += d.y * v[j].y * rstd;
                acc[j].z += d.z * v[j].z * rstd; acc[j].w += d.w * v[j].w * rstd;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = lane + 32 * j;
        if (i < n4) {
            atomicAdd(&part[4 * i], acc[j].x); atomicAdd(&part[4 * i + 1], acc[j].y);
            atomicAdd(&part[4 * i + 2], acc[j].z); atomicAdd(&part[4 * i + 3], acc[j].w);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dg + c, part[c]);
}

// ---------------------------------------------------------------------------------------------- head 1x1 conv parameters
// dW[o, c] += sum_{b,p} g[b, o, p] x[(b, p), c];  db[o] += sum g[b, o, p]     (g NCHW fp32, x NHWC operand, Cout <= 4)
template <typename TX>
__global__ void __launch_bounds__(256) head_wgrad_kernel(const float* __restrict__ g, const TX* __restrict__ x,
                                                         float* __restrict__ dw, float* __restrict__ db, int HW, int Cin, int Cout,
                                                         int64_t M, int rows_per_cta) {
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(r0 + (int64_t)rows_per_cta, M);
    const int groups = max(1, 256 / Cin), grp = threadIdx.x / Cin;
    float sb[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = threadIdx.x % Cin; c < Cin && grp < groups; c += 256) {
        float s[4] = {0.f, 0.f, 0.f, 0.f};
        for (int64_t r = r0 + grp; r < r1; r += groups) {
            const int64_t b = r / HW, pp = r % HW;
            const float xv = (float)x[r * Cin + c];
            for (int o = 0; o < Cout; ++o) {
                const float gv = g[(b * Cout + o) * HW + pp];
                s[o] += gv * xv;
                if (c == 0) sb[o] += gv;
            }
        }
        for (int o = 0; o < Cout; ++o) atomicAdd(dw + o * Cin + c, s[o]);
        if (Cin <= 256) break;
    }
    if (db && threadIdx.x % Cin == 0 && grp < groups)
        for (int o = 0; o < Cout; ++o) atomicAdd(db + o, sb[o]);
}

}  // namespace sdc

using namespace sdc;

extern "C" int sdc_conv_wgrad(int kind, int a_half, const void* a0, int c0, const void* a1, int c1, const float* dy, float* dw,
                              int B, int H, int W, int Cout, void* stream) {
    SDC_REQUIRE(kind >= 0 && kind <= 2 && a0 && dy && dw && B > 0 && H > 0 && W > 0 && Cout > 0, "conv_wgrad: bad arguments");
    SDC_REQUIRE(c0 > 0 && c1 >= 0 && (c1 == 0 || a1), "conv_wgrad: bad input segments");
    const int apv = a_half ? 8 : 4;
    SDC_REQUIRE(c0 % apv == 0 && c1 % apv == 0 && Cout % 4 == 0, "conv_wgrad: channel counts must be multiples of %d (Cout of 4)", apv);
    WgradArgs p{};
    p.a0 = a0; p.a1 = a1; p.dy = dy; p.dw = dw; p.c0 = c0; p.c1 = c1; p.kind = kind; p.B = B; p.H = H; p.W = W; p.Cout = Cout;
    p.M = (int64_t)B * H * W;
    p.n_chunks = (int)((p.M + WG_PX - 1) / WG_PX);
    const int taps = kind == 1 ? 9 : (kind == 2 ? 4 : 1);
    const int Cin = c0 + c1;
    const int64_t tiles = (int64_t)taps * ((Cin + WG_T - 1) / WG_T) * ((Cout + WG_T - 1) / WG_T);
    // split the pixel axis until ~4 CTAs per SM exist, keeping at least 8 chunks per CTA
    int splits = (int)((4 * 148 + tiles - 1) / tiles);
    splits = splits < 1 ? 1 : splits;
    const int max_splits = (p.n_chunks + 7) / 8;
    if (splits > max_splits) splits = max_splits;
    p.chunks_per_split = (p.n_chunks + splits - 1) / splits;
    splits = (p.n_chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    dim3 grid((unsigned)(taps * ((Cin + WG_T - 1) / WG_T)), (unsigned)((Cout + WG_T - 1) / WG_T), (unsigned)splits);
    if (a_half) conv_wgrad_kernel<__half><<<grid, 128, 0, as_stream(stream)>>>(p);
    else conv_wgrad_kernel<float><<<grid, 128, 0, as_stream(stream)>>>(p);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_colsum(const float* x, float* out, int64_t M, int C, void* stream) {
    SDC_REQUIRE(x && out && M > 0 && C > 0 && C % 4 == 0 && C <= 8192, "colsum: bad arguments (C=%d)", C);
    int rows = (int)((M + 2 * 148 - 1) / (2 * 148));
    if (rows < 64) rows = 64;
    colsum_kernel<<<(unsigned)((M + rows - 1) / rows), 256, C * sizeof(float), as_stream(stream)>>>(x, out, M, C, rows);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_gn_param_grad(const float* dy, const float* x, const double* stats, const float* gamma, const float* beta,
                                 const float* scale_shift, const int32_t* t_index, int64_t ss_stride, float* P, int B, int HW, int C,
                                 void* stream) {
    SDC_REQUIRE(dy && x && stats && gamma && beta && P && B > 0 && HW > 0, "gn_param_grad: bad arguments");
    SDC_REQUIRE(C % 4 == 0 && C <= 4096, "gn_param_grad: C=%d unsupported", C);
    int ppc = HW;
    while (ppc > 32 && (int64_t)B * (HW / ppc) < 2 * 148 && ppc % 2 == 0) ppc /= 2;
    dim3 grid((unsigned)B, (unsigned)((HW + ppc - 1) / ppc));
    gn_param_grad_kernel<<<grid, 256, 4 * C * sizeof(float), as_stream(stream)>>>(dy, x, stats, gamma, beta, scale_shift, t_index,
                                                                                ss_stride, P, HW, C, ppc);
    SDC_LAUNCHED();
    return SDC_OK;
}

template <typename TX>
static void launch_ln_gain(const float* dy, const void* x, float* dg, int64_t M, int C, cudaStream_t st) {
    const TX* xp = reinterpret_cast<const TX*>(x);
    int64_t want = (M + 7) / 8;
    const unsigned blocks = (unsigned)(want > 4 * 148 ? 4 * 148 : (want < 1 ? 1 : want));
    const size_t sm = C * sizeof(float);
    const int n4 = C / 4;
    if (n4 <= 32) ln_gain_grad_kernel<1, TX><<<blocks, 256, sm, st>>>(dy, xp, dg, M, C);
    else if (n4 <= 64) ln_gain_grad_kernel<2, TX><<<blocks, 256, sm, st>>>(dy, xp, dg, M, C);
    else if (n4 <= 128) ln_gain_grad_kernel<4, TX><<<blocks, 256, sm, st>>>(dy, xp, dg, M, C);
    else ln_gain_grad_kernel<8, TX><<<blocks, 256, sm, st>>>(dy, xp, dg, M, C);
}

extern "C" int sdc_channel_layernorm_gain_grad(const float* dy, const void* x, int x_half, float* dg, int64_t M, int C, void* stream) {
    SDC_REQUIRE(dy && x && dg && M > 0 && C % 4 == 0 && C <= 1024, "channel_layernorm_gain_grad: bad arguments (C=%d)", C);
    if (x_half) launch_ln_gain<__half>(dy, x, dg, M, C, as_stream(stream));
    else launch_ln_gain<float>(dy, x, dg, M, C, as_stream(stream));
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_head_conv1_wgrad(const float* g, const void* x, int x_half, float* dw, float* db, int B, int HW, int Cin,
                                    int Cout, void* stream) {
    SDC_REQUIRE(g && x && dw && B > 0 && HW > 0 && Cin > 0 && Cout > 0 && Cout <= 4, "head_conv1_wgrad: bad arguments (Cout <= 4)");
    const int64_t M = (int64_t)B * HW;
    int rows = (int)((M + 2 * 148 - 1) / (2 * 148));
    if (rows < 64) rows = 64;
    const unsigned blocks = (unsigned)((M + rows - 1) / rows);
    if (x_half) head_wgrad_kernel<__half><<<blocks, 256, 0, as_stream(stream)>>>(g, reinterpret_cast<const __half*>(x), dw, db, HW, Cin, Cout, M, rows);
    else head_wgrad_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(g, reinterpret_cast<const float*>(x), dw, db, HW, Cin, Cout, M, rows);
    SDC_LAUNCHED();
    return SDC_OK;
}
