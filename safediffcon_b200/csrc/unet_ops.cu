// Bandwidth-bound building blocks of the denoiser plus weight packing,
// the 3-channel stem conv and the 3-channel head conv.  Reference: /root/reference/1D/model/unet.py.
// Activations are NHWC fp32 pixel rows [B*H*W, C]; every kernel here is a single pass over its tensors with
// 16-byte coalesced accesses (the convolutions in conv_gemm.cu are the only compute-bound part).
#include "common.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <math.h>
#include <stdlib.h>

namespace sdc {

// ---------------------------------------------------------------------------------------------- weight packing
// kind 3 (nearest-upsample x2 + 3x3): Wp[(phase*Cout + co), tap*Cin + ci] with phase = 2a + b (output pixel parity), tap = 2r + s
// (position in the 2x2 low-resolution window): the sum of the 3x3 taps (ky, kx) whose upsampled pixel (2i+a+ky-1, 2j+b+kx-1)
// falls on low-resolution pixel (i+a-1+r, j+b-1+s): a = 0: r = 0 <- ky 0, r = 1 <- ky 1,2;  a = 1: r = 0 <- ky 0,1, r = 1 <- ky 2.
template <typename T>
__global__ void pack_upconv_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, int Cout, int Cin) {
    const int64_t total = (int64_t)16 * Cout * Cin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int K = 4 * Cin;
        const int row = (int)(i / K), k = (int)(i % K);
        const int phase = row / Cout, co = row % Cout, tap = k / Cin, ci = k % Cin;
        const int a = phase >> 1, b = phase & 1, r = tap >> 1, s2 = tap & 1;
        const int ky0 = a == 0 ? (r == 0 ? 0 : 1) : (r == 0 ? 0 : 2), ky1 = a == 0 ? (r == 0 ? 0 : 2) : (r == 0 ? 1 : 2);
        const int kx0 = b == 0 ? (s2 == 0 ? 0 : 1) : (s2 == 0 ? 0 : 2), kx1 = b == 0 ? (s2 == 0 ? 0 : 2) : (s2 == 0 ? 1 : 2);
        const float* wk = w + ((int64_t)co * Cin + ci) * 9;
        float v = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) v += wk[ky * 3 + kx];
        wp[i] = to_operand(v, T());
    }
}

template <typename T>
__global__ void pack_conv_weight_kernel(int kind, const float* __restrict__ w, T* __restrict__ wp, int Cout, int Cin) {
    const int taps = kind == 1 ? 9 : 1;
    const int64_t total = (int64_t)Cout * Cin * taps;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int K = Cin * taps;
        const int o = (int)(i / K), k = (int)(i % K);
        float v;
        if (kind == 1) {
            const int tap = k / Cin, ci = k % Cin;
            v = w[((int64_t)o * Cin + ci) * 9 + tap];
        } else if (kind == 2) {
            const int C = Cin / 4, pp = k / C, c = k % C;
            v = w[(int64_t)o * Cin + c * 4 + pp];
        } else {
            v = w[i];
        }
        wp[i] = to_operand(v, T());
    }
}

// ---------------------------------------------------------------------------------------------- stem 7x7 conv
// One CTA = STEM_ROWS image rows of one sample; thread = 8 pixels x 8 output channels in registers.
constexpr int STEM_ROWS = 4;
template <typename T>
__global__ void stem_conv7_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  T* __restrict__ out, int Cin, int H, int W, int Cout) {
    extern __shared__ float sm[];
    const int K = Cin * 49;
    float* ws = sm;                      // [K][Cout]
    float* xs = sm + (size_t)K * Cout;   // [Cin][7][W + 6]  (one output row at a time)
    const int b = blockIdx.x, h_base = blockIdx.y * STEM_ROWS;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int i = tid; i < K * Cout; i += nthr) {
        const int co = i / K, k = i % K;            // coalesced read of OIHW
        ws[k * Cout + co] = w[i];
    }
    const int cog = Cout / 8;
    const int co0 = (tid % cog) * 8, px0 = (tid / cog) * 8;
    const int WP = W + 6;
    for (int hr = 0; hr < STEM_ROWS; ++hr) {
        const int h = h_base + hr;
        if (h >= H) break;
        __syncthreads();
        for (int i = tid; i < Cin * 7 * WP; i += nthr) {
            const int ci = i / (7 * WP), r = (i / WP) % 7, c = i % WP;
            const int hh = h + r - 3, ww = c - 3;
            xs[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((int64_t)b * Cin + ci) * H + hh) * W + ww] : 0.f;
        }
        __syncthreads();
        float acc[8][8];
#pragma unroll
        for (int p = 0; p < 8; ++p)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[p][c] = 0.f;
        for (int cr = 0; cr < Cin * 7; ++cr) {
            float xv[14];
#pragma unroll
            for (int j = 0; j < 14; ++j) xv[j] = xs[cr * WP + px0 + j];
#pragma unroll
            for (int kx = 0; kx < 7; ++kx) {
                const float4 w0 = *reinterpret_cast<const float4*>(&ws[(cr * 7 + kx) * Cout + co0]);
                const float4 w1 = *reinterpret_cast<const float4*>(&ws[(cr * 7 + kx) * Cout + co0 + 4]);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int p = 0; p < 8; ++p)
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[p][c] = fmaf(xv[p + kx], wv[c], acc[p][c]);
            }
        }
        float bv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) bv[c] = bias ? bias[co0 + c] : 0.f;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            T* o = out + (((int64_t)b * H + h) * W + px0 + p) * Cout + co0;
            store_operand4(o, make_float4(acc[p][0] + bv[0], acc[p][1] + bv[1], acc[p][2] + bv[2], acc[p][3] + bv[3]));
            store_operand4(o + 4, make_float4(acc[p][4] + bv[4], acc[p][5] + bv[5], acc[p][6] + bv[6], acc[p][7] + bv[7]));
        }
    }
}

// ---------------------------------------------------------------------------------------------- stem im2col (tensor-core stem)
// A[m, :] for the 7x7 pad-3 stem as a GEMM operand: columns [0, K) hold the HIGH part of the patch (K = Cin*49, index
// ci*49 + ky*7 + kx), columns [KH, KH+K) the LOW part (x - high, rounded to the operand precision), the rest zeros.
// With the stem weights repeated in both column ranges the tcgen05 GEMM sees the fp32 input to ~2^-22 although its
// operands are fp16 / tf32.  One CTA = one image row; the 7 input rows it needs are staged in shared memory.
// Thread -> one group of 8 consecutive patch columns for pixels w = wl, wl + blockDim / groups, ...; it writes the group's high part
// and its low part (two 16-byte stores in FP16 mode) from the same 8 shared loads.  The patch offsets of its columns live in
// registers (the first version re-derived (pixel, column) by integer division per 4 elements: issue bound, 0.29 of the copy peak).
template <typename T>
__global__ void __launch_bounds__(256) stem_im2col_kernel(const float* __restrict__ x, T* __restrict__ a, int Cin, int H, int W,
                                                          int kp) {
    extern __shared__ float xs[];   // [Cin][7][W + 6]
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int WP = W + 6, K = Cin * 49, KH = kp / 2;
    for (int i = threadIdx.x; i < Cin * 7 * WP; i += blockDim.x) {
        const int ci = i / (7 * WP), r = (i / WP) % 7, c = i % WP;
        const int hh = h + r - 3, ww = c - 3;
        xs[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? x[(((int64_t)b * Cin + ci) * H + hh) * W + ww] : 0.f;
    }
    __syncthreads();
    const int groups = KH / 8;                       // groups of 8 patch columns; a thread writes the group's high AND low part
    const int wstep = blockDim.x / groups;           // pixels in flight per pass (threads beyond groups * wstep idle)
    const int grp = threadIdx.x % groups, wl = threadIdx.x / groups;
    if (wl >= wstep) return;
    int off[8];       // offset of column (ci, ky, kx) inside the window, or -1 for padding columns
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = grp * 8 + j;
        const int ci = k / 49, t = k - ci * 49, ky = t / 7, kx = t - ky * 7;
        off[j] = k < K ? (ci * 7 + ky) * WP + kx : -1;
    }
    T* arow = a + ((int64_t)b * H + h) * W * kp + grp * 8;
    for (int w = wl; w < W; w += wstep) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = off[j] >= 0 ? xs[off[j] + w] : 0.f;
        if constexpr (sizeof(T) == 2) {
            uint4 hi, lo;
            uint32_t* hw = reinterpret_cast<uint32_t*>(&hi);
            uint32_t* lw = reinterpret_cast<uint32_t*>(&lo);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half2 h2 = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
                const float2 f2 = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(v[2 * j] - f2.x, v[2 * j + 1] - f2.y);
                hw[j] = *reinterpret_cast<const uint32_t*>(&h2);
                lw[j] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            *reinterpret_cast<uint4*>(arow + (int64_t)w * kp) = hi;
            *reinterpret_cast<uint4*>(arow + (int64_t)w * kp + KH) = lo;
        } else {
            float hi[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) hi[j] = (float)to_operand(v[j], T());
            store_operand4(arow + (int64_t)w * kp, make_float4(hi[0], hi[1], hi[2], hi[3]));
            store_operand4(arow + (int64_t)w * kp + 4, make_float4(hi[4], hi[5], hi[6], hi[7]));
            store_operand4(arow + (int64_t)w * kp + KH, make_float4(v[0] - hi[0], v[1] - hi[1], v[2] - hi[2], v[3] - hi[3]));
            store_operand4(arow + (int64_t)w * kp + KH + 4, make_float4(v[4] - hi[4], v[5] - hi[5], v[6] - hi[6], v[7] - hi[7]));
        }
    }
}

// ---------------------------------------------------------------------------------------------- GroupNorm(1)+FiLM+SiLU
__device__ __forceinline__ float silu(float v) { return v / (1.0f + expf(-v)); }
// fast variant for the bandwidth-bound apply kernel: MUFU.EX2 + MUFU.RCP (relative error ~1e-6, far below the 2^-11 of the
// operand rounding that follows)
__device__ __forceinline__ float silu_fast(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// One CTA = `pix_per_cta` pixels of one sample, 256 threads.  When 256 is a multiple of C/4 (every channel count of the
// U-Net) a thread always meets the same 4 channels, so its scale/offset pairs live in registers and the loop is
// 4 independent 16-byte loads -> 16 fma/ex2/rcp -> 4 stores per iteration.
template <typename TR, typename TY>
__global__ void __launch_bounds__(256) gn_silu_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ scale_shift, const int32_t* __restrict__ t_index,
                                                      int64_t ss_stride, const TR* __restrict__ residual, TY* __restrict__ y,
                                                      int HW, int C, int pix_per_cta) {
    extern __shared__ float coef[];  // A[C], B[C]
    const int b = blockIdx.x;
    const double cnt = (double)HW * (double)C;
    const double mean_d = stats[2 * b] / cnt;
    double var_d = stats[2 * b + 1] / cnt - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + 1e-5));
    const float* ss = scale_shift ? scale_shift + (int64_t)(t_index ? t_index[b] : 0) * ss_stride : nullptr;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float g = gamma[c] * rstd;
        float a = g, bb = beta[c] - mean * g;
        if (ss) {
            const float sc = ss[c] + 1.0f;
            a *= sc;
            bb = bb * sc + ss[C + c];
        }
        coef[c] = a;
        coef[C + c] = bb;
    }
    __syncthreads();
    const int c4n = C / 4;
    const int64_t row0 = (int64_t)b * HW + (int64_t)blockIdx.y * pix_per_cta;
    const int rows = min(pix_per_cta, HW - (int)blockIdx.y * pix_per_cta);
    const float4* x4 = reinterpret_cast<const float4*>(x + row0 * C);
    const TR* r4 = residual ? residual + row0 * C : nullptr;
    TY* y4 = y + row0 * C;
    const int total = rows * c4n;
    if (256 % c4n == 0) {
        const int c = (threadIdx.x % c4n) * 4;
        const float a0 = coef[c], a1 = coef[c + 1], a2 = coef[c + 2], a3 = coef[c + 3];
        const float b0 = coef[C + c], b1 = coef[C + c + 1], b2 = coef[C + c + 2], b3 = coef[C + c + 3];
        int i = threadIdx.x;
        for (; i + 3 * 256 < total; i += 4 * 256) {
            float4 v[4], r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldcs(x4 + i + u * 256);   // streamed once: evict first
            if (r4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) r[u] = load4(r4 + 4 * (int64_t)(i + u * 256));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u].x = silu_fast(fmaf(v[u].x, a0, b0)); v[u].y = silu_fast(fmaf(v[u].y, a1, b1));
                v[u].z = silu_fast(fmaf(v[u].z, a2, b2)); v[u].w = silu_fast(fmaf(v[u].w, a3, b3));
                if (r4) { v[u].x += r[u].x; v[u].y += r[u].y; v[u].z += r[u].z; v[u].w += r[u].w; }
                store_operand4(y4 + 4 * (int64_t)(i + u * 256), v[u]);
            }
        }
        for (; i < total; i += 256) {
            float4 v = x4[i];
            v.x = silu_fast(fmaf(v.x, a0, b0)); v.y = silu_fast(fmaf(v.y, a1, b1));
            v.z = silu_fast(fmaf(v.z, a2, b2)); v.w = silu_fast(fmaf(v.w, a3, b3));
            if (r4) { const float4 r = load4(r4 + 4 * (int64_t)i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
            store_operand4(y4 + 4 * (int64_t)i, v);
        }
        return;
    }
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = (i % c4n) * 4;
        float4 v = x4[i];
        v.x = silu_fast(fmaf(v.x, coef[c], coef[C + c]));
        v.y = silu_fast(fmaf(v.y, coef[c + 1], coef[C + c + 1]));
        v.z = silu_fast(fmaf(v.z, coef[c + 2], coef[C + c + 2]));
        v.w = silu_fast(fmaf(v.w, coef[c + 3], coef[C + c + 3]));
        if (r4) { const float4 r = load4(r4 + 4 * (int64_t)i); v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
        store_operand4(y4 + 4 * (int64_t)i, v);
    }
}

// FP16-input variant (compact intermediates: the conv epilogue took the statistics from its fp32 accumulators and stored the
// activations as fp16): a thread owns 8 channels (256 % (C/8) == 0), 4 independent 16-byte loads in flight, 16-byte stores.
// Bytes per element: 2 in + 2 out (+2 residual) instead of 4 + 2 (+2|4).
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) { const float2 t = __half22float2(h[k]); f[2 * k] = t.x; f[2 * k + 1] = t.y; }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(f[2 * k], f[2 * k + 1]);
    return u;
}
// channel mean / rstd of one pixel row held as 8 fp16 values per lane by c8n (16 or 32) adjacent lanes; vector index i = row * c8n + lane.
// One pass (sum and sum of squares reduced side by side: two independent shuffle chains of log2(c8n) steps instead of two dependent
// passes) on the ROUNDED values the projection will read; fp32 is ample for <= 256 values of O(1).  The whole warp takes part.
__device__ __forceinline__ void row_stats8(const uint4& o, int c8n, int C, float2* __restrict__ rs, int i) {
    float v[8];
    unpack8(o, v);
    float sm = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) { sm += v[k]; q = fmaf(v[k], v[k], q); }
    for (int m = 1; m < c8n; m <<= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, m);
        q += __shfl_xor_sync(0xffffffffu, q, m);
    }
    const float inv = 1.0f / (float)C;
    const float mean = sm * inv;
    if (i % c8n == 0) rs[i / c8n] = make_float2(mean, rsqrtf(fmaxf(fmaf(-mean, mean, q * inv), 0.f) + 1e-5f));
}

// ROWSTATS: additionally emit, per pixel row of the (fp16-rounded) OUTPUT, the channel mean and 1 / sqrt(var + 1e-5) of the channel
// LayerNorm that follows in PreNorm(LinearAttention) (unet.py:53-63,65-76): the LayerNorm itself is then folded into the qkv
// projection (sdc_conv1x1_qkv_ln) and its separate pass over HBM disappears.  A row is owned by C/8 = 16 or 32 adjacent lanes.
template <bool ROWSTATS>
__global__ void __launch_bounds__(256, ROWSTATS ? 3 : 4) gn_silu_h8_kernel(const __half* __restrict__ x, const double* __restrict__ stats,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ scale_shift, const int32_t* __restrict__ t_index,
                                                         int64_t ss_stride, const __half* __restrict__ residual, __half* __restrict__ y,
                                                         int HW, int C, int pix_per_cta, float2* __restrict__ rowstats) {
    const int b = blockIdx.x;
    const double cnt = (double)HW * (double)C;
    const double mean_d = stats[2 * b] / cnt;
    double var_d = stats[2 * b + 1] / cnt - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + 1e-5));
    const float* ss = scale_shift ? scale_shift + (int64_t)(t_index ? t_index[b] : 0) * ss_stride : nullptr;
    const int c8n = C / 8;
    const int c = (threadIdx.x % c8n) * 8;
    float A[8], Bc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float g = gamma[c + k] * rstd;
        float a = g, bb = beta[c + k] - mean * g;
        if (ss) {
            const float sc = ss[c + k] + 1.0f;
            a *= sc;
            bb = bb * sc + ss[C + c + k];
        }
        A[k] = a;
        Bc[k] = bb;
    }
    const int64_t row0 = (int64_t)b * HW + (int64_t)blockIdx.y * pix_per_cta;
    const int rows = min(pix_per_cta, HW - (int)blockIdx.y * pix_per_cta);
    const uint4* x8 = reinterpret_cast<const uint4*>(x + row0 * C);
    const uint4* r8 = residual ? reinterpret_cast<const uint4*>(residual + row0 * C) : nullptr;
    uint4* y8 = reinterpret_cast<uint4*>(y + row0 * C);
    const int total = rows * c8n;
    int i = threadIdx.x;
    for (; i + 3 * 256 < total; i += 4 * 256) {
        uint4 v[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldcs(x8 + i + u * 256);   // streamed once: evict first
        if (r8) {
#pragma unroll
            for (int u = 0; u < 4; ++u) r[u] = r8[i + u * 256];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float f[8], rf[8];
            unpack8(v[u], f);
            if (r8) unpack8(r[u], rf);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                f[k] = silu_fast(fmaf(f[k], A[k], Bc[k]));
                if (r8) f[k] += rf[k];
            }
            const uint4 o = pack8(f);
            y8[i + u * 256] = o;
            if constexpr (ROWSTATS) row_stats8(o, c8n, C, rowstats + row0, i + u * 256);
        }
    }
    for (; i < total; i += 256) {
        float f[8], rf[8];
        unpack8(x8[i], f);
        if (r8) unpack8(r8[i], rf);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            f[k] = silu_fast(fmaf(f[k], A[k], Bc[k]));
            if (r8) f[k] += rf[k];
        }
        const uint4 o = pack8(f);
        y8[i] = o;
        if constexpr (ROWSTATS) row_stats8(o, c8n, C, rowstats + row0, i);
    }
}

// Last GroupNorm of the network fused with the 1x1 head convolution (final_res_block.block2's norm + SiLU + residual, then
// final_conv, unet.py:178-180,378,426): eps[b, o, p] = sum_c w[o, c] * (silu(GN(x)[p, c]) + res[p, c]) + bias[o].  The activation
// never exists in memory: it stays fp32 in registers, so the head sees it unrounded (the last rounding site in front of the
// output carried 13% of the eps error variance) and one write + one read of a [B*HW, C] tensor disappear.
// x: fp32 conv output (C = 128); residual: fp16 or fp32.  16 lanes per pixel (8 channels each), a warp covers 2 pixels; the
// (<= 4) dot products are reduced with a transposing butterfly: 5 shuffles instead of 16.
template <typename TR>
__global__ void __launch_bounds__(256, 3) gn_silu_head_kernel(const float* __restrict__ x, const double* __restrict__ stats,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           const TR* __restrict__ residual, const float* __restrict__ hw,
                                                           const float* __restrict__ hb, float* __restrict__ out, int HW, int Cout,
                                                           int pix_per_cta) {
    constexpr int C = 128;
    const int b = blockIdx.x;
    const double cnt = (double)HW * (double)C;
    const double mean_d = stats[2 * b] / cnt;
    double var_d = stats[2 * b + 1] / cnt - mean_d * mean_d;
    if (var_d < 0.0) var_d = 0.0;
    const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + 1e-5));
    const int lane = threadIdx.x & 31, sub = lane & 15;
    const int c = sub * 8;
    float A[8], Bc[8], w[4][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float g = gamma[c + k] * rstd;
        A[k] = g;
        Bc[k] = beta[c + k] - mean * g;
#pragma unroll
        for (int o = 0; o < 4; ++o) w[o][k] = o < Cout ? hw[o * C + c + k] : 0.f;
    }
    const int o_mine = ((sub & 1) << 1) | ((sub >> 1) & 1);   // output channel this lane ends up holding (lanes 0-3 of a half warp)
    const float bias = (o_mine < Cout && hb) ? hb[o_mine] : 0.f;
    const int p0 = blockIdx.y * pix_per_cta, p1 = min(HW, p0 + pix_per_cta);
    const int pl = threadIdx.x >> 4;   // 16 pixels per pass
    for (int pp = p0 + pl; pp < p1; pp += 32) {   // two pixels per lane in flight; pix_per_cta % 32 == 0: uniform trip count (shuffles below)
        float d[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int px = pp + 16 * u;
            const bool ok = px < p1;
            const int64_t off = ((int64_t)b * HW + (ok ? px : p0)) * C + c;
            const float4 x0 = __ldcs(reinterpret_cast<const float4*>(x + off)), x1 = __ldcs(reinterpret_cast<const float4*>(x + off + 4));
            const float4 r0 = load4(residual + off), r1 = load4(residual + off + 4);
            const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            const float rv[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            d[u][0] = d[u][1] = d[u][2] = d[u][3] = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float y = silu_fast(fmaf(xv[k], A[k], Bc[k])) + rv[k];
#pragma unroll
                for (int o = 0; o < 4; ++o) d[u][o] = fmaf(y, w[o][k], d[u][o]);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            // step 1 (xor 1): even lanes keep (d0, d1), odd lanes keep (d2, d3); step 2 (xor 2): bit 1 clear keeps the first
            const bool odd = sub & 1, hi = sub & 2;
            float k0 = odd ? d[u][2] : d[u][0], k1 = odd ? d[u][3] : d[u][1];
            k0 += __shfl_xor_sync(0xffffffffu, odd ? d[u][0] : d[u][2], 1);
            k1 += __shfl_xor_sync(0xffffffffu, odd ? d[u][1] : d[u][3], 1);
            float kk = hi ? k1 : k0;
            kk += __shfl_xor_sync(0xffffffffu, hi ? k0 : k1, 2);
            kk += __shfl_xor_sync(0xffffffffu, kk, 4);
            kk += __shfl_xor_sync(0xffffffffu, kk, 8);
            const int px = pp + 16 * u;
            if (sub < 4 && o_mine < Cout && px < p1) out[((int64_t)b * Cout + o_mine) * HW + px] = kk + bias;
        }
    }
}

// ---------------------------------------------------------------------------------------------- channel LayerNorm
// one warp per RPW pixel rows (RPW = 4 for C <= 256 so that 4-8 independent 16-byte loads are in flight per lane);
// C <= 1024 kept in registers, two-pass variance like torch.var(unbiased=False)
template <int RPW, int NV, typename TX, typename TY>
__global__ void __launch_bounds__(256) channel_layernorm_kernel(const TX* __restrict__ x, const float* __restrict__ g,
                                                                const TY* __restrict__ residual, TY* __restrict__ y,
                                                                int64_t M, int C, int operand_out) {
    const int lane = threadIdx.x & 31;
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (row0 >= M) return;
    const int n4 = C / 4;  // float4 per row; lane handles i = lane, lane+32, ...
    float4 v[RPW][NV];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const TX* x4 = x + (row0 + r) * C;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            v[r][j] = (i < n4 && row0 + r < M) ? load4(x4 + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        if (row0 + r >= M) break;
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) s += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            if (lane + 32 * j < n4) {
                const float a = v[r][j].x - mean, b = v[r][j].y - mean, c = v[r][j].z - mean, d = v[r][j].w - mean;
                q += (a * a + b * b) + (c * c + d * d);
            }
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + 1e-5f);
        const TY* r4 = residual ? residual + (row0 + r) * C : nullptr;
        TY* y4 = y + (row0 + r) * C;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int i = lane + 32 * j;
            if (i < n4) {
                const float4 gv = g4[i];
                float4 o = make_float4((v[r][j].x - mean) * rstd * gv.x, (v[r][j].y - mean) * rstd * gv.y,
                                       (v[r][j].z - mean) * rstd * gv.z, (v[r][j].w - mean) * rstd * gv.w);
                if (r4) { const float4 rr = load4(r4 + 4 * i); o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w; }
                if constexpr (sizeof(TY) == 2) store_operand4(y4 + 4 * i, o);
                else { if (operand_out) store_operand4(y4 + 4 * i, o); else store4(reinterpret_cast<float*>(y4) + 4 * i, o); }
            }
        }
    }
}

// FP16 in / FP16 out variant (every LayerNorm of the compact FP16 inference path).  ncu on the kernel above at C = 128: 0.59 of
// the copy peak with 81% issue utilisation -- a 32-lane row costs 10 shuffles for 4 elements per lane.  Here a row is owned by
// LPR lanes (8 for C = 128) holding VPL 16-byte vectors (8 channels) each: 2 log2(LPR) shuffles per 8 VPL elements, 32 / LPR rows
// per warp pass, and RG independent row groups in flight per lane.  Lane l of a row reads vectors l, l + LPR, ... (coalesced).
template <int LPR, int VPL, int RG>
__global__ void __launch_bounds__(256) channel_layernorm_h_kernel(const __half* __restrict__ x, const float* __restrict__ g,
                                                                  const __half* __restrict__ residual, __half* __restrict__ y,
                                                                  int64_t M) {
    constexpr int C = LPR * VPL * 8, RPP = 32 / LPR;   // rows per warp pass
    const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
    const int64_t warp_id = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t row0 = warp_id * (RPP * RG) + rsel;
    float gv[VPL][8];
#pragma unroll
    for (int j = 0; j < VPL; ++j) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(g + (j * LPR + sub) * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(g + (j * LPR + sub) * 8 + 4));
        gv[j][0] = a.x; gv[j][1] = a.y; gv[j][2] = a.z; gv[j][3] = a.w; gv[j][4] = b.x; gv[j][5] = b.y; gv[j][6] = b.z; gv[j][7] = b.w;
    }
    uint4 xv[RG][VPL], rv[RG][VPL];
#pragma unroll
    for (int r = 0; r < RG; ++r) {
        const int64_t row = row0 + (int64_t)r * RPP;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            xv[r][j] = make_uint4(0u, 0u, 0u, 0u);
            rv[r][j] = make_uint4(0u, 0u, 0u, 0u);
            if (row < M) {
                xv[r][j] = __ldcs(reinterpret_cast<const uint4*>(x + row * C) + j * LPR + sub);
                if (residual) rv[r][j] = __ldg(reinterpret_cast<const uint4*>(residual + row * C) + j * LPR + sub);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RG; ++r) {
        const int64_t row = row0 + (int64_t)r * RPP;
        float f[VPL][8];
        float sm = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j) {
            unpack8(xv[r][j], f[j]);
#pragma unroll
            for (int k = 0; k < 8; ++k) sm += f[j][k];
        }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) sm += __shfl_xor_sync(0xffffffffu, sm, o);
        const float mean = sm * (1.0f / (float)C);
        float q = 0.f;
#pragma unroll
        for (int j = 0; j < VPL; ++j)
#pragma unroll
            for (int k = 0; k < 8; ++k) { f[j][k] -= mean; q = fmaf(f[j][k], f[j][k], q); }
#pragma unroll
        for (int o = 1; o < LPR; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rstd = rsqrtf(q * (1.0f / (float)C) + 1e-5f);
        if (row < M) {
#pragma unroll
            for (int j = 0; j < VPL; ++j) {
                float rf[8];
                if (residual) unpack8(rv[r][j], rf);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    f[j][k] = f[j][k] * rstd * gv[j][k];
                    if (residual) f[j][k] += rf[k];
                }
                reinterpret_cast<uint4*>(y + row * C)[j * LPR + sub] = pack8(f[j]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------- linear attention
constexpr int LA_HEADS = 4, LA_D = 32, LA_QKV = 3 * LA_HEADS * LA_D, LA_HID = LA_HEADS * LA_D;
// workspace floats per (b, head): ctx[32][32] | kmax[32] | ksum[32]  (the softmax_n(k) statistics are kept for the backward pass)
constexpr int LA_CTX = LA_D * LA_D + 2 * LA_D;

// ctx[b,h,d,e] = sum_n softmax_n(k)[d,n] * v[e,n]          grid = B*heads, 256 threads
// 4 pixel groups x 64 threads; each thread owns a 4x4 block of the 32x32 context in registers (16 FMA per two
// 16-byte shared loads).  SINGLE pass over k and v (each read once from HBM): the column maximum of k is tracked online --
// per 128-pixel chunk the running maximum m_d is raised to the chunk's and the accumulators of row d are rescaled by
// exp(m_old - m_new) (softmax is invariant under the shift; the two-pass version read k twice: 3.1 -> 2.1 GB per 16x128 call).
constexpr int LA_CHUNK = 128;
// k / v rows: kbase / vbase + pixel * ld (+ head * 32); ld = 384 for a packed qkv tensor, 256 for the kv tensor of the fused path
// TK = float (packed fp32 qkv / kv tensors) or __half (kv written in the operand precision by the fused qkv projection: k only
// enters through exp(k - max) and v through a 2048-term sum, rounding them to 10 bits moves eps by < 1e-6 relative)
template <typename TK>
__global__ void __launch_bounds__(256) linattn_context_kernel(const TK* __restrict__ kbase, const TK* __restrict__ vbase, int ld,
                                                              float* __restrict__ ctx, int n) {
    __shared__ __align__(16) float ek[LA_CHUNK][LA_D];
    __shared__ __align__(16) float vs[LA_CHUNK][LA_D];
    __shared__ float red[8][LA_D];
    __shared__ float kmax[LA_D];
    __shared__ float rescale[LA_D];
    const int b = blockIdx.x / LA_HEADS, h = blockIdx.x % LA_HEADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const TK* kp = kbase + (int64_t)b * n * ld + h * LA_D;
    const TK* vp = vbase + (int64_t)b * n * ld + h * LA_D;
    if (tid < LA_D) kmax[tid] = -INFINITY;
    const int ng = tid >> 6, t64 = tid & 63;
    const int d0 = (t64 >> 3) * 4, e0 = (t64 & 7) * 4;
    float acc[4][4], ksum[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ksum[i] = 0.f; for (int j = 0; j < 4; ++j) acc[i][j] = 0.f; }
    const int lc4 = (tid & 7) * 4, lr = tid >> 3;   // loader role: 4 channels x pixel rows lr, lr+32, lr+64, lr+96 of the chunk
    for (int n0 = 0; n0 < n; n0 += LA_CHUNK) {
        const int cnt = min(LA_CHUNK, n - n0);
        float4 kraw[LA_CHUNK / 32];
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        __syncthreads();   // the previous chunk's tiles are consumed
#pragma unroll
        for (int rr = 0; rr < LA_CHUNK / 32; ++rr) {
            const int r = rr * 32 + lr;
            float4 kv = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), vv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < cnt) {
                kv = load4(kp + (int64_t)(n0 + r) * ld + lc4);
                vv = load4(vp + (int64_t)(n0 + r) * ld + lc4);
            }
            kraw[rr] = kv;
            m.x = fmaxf(m.x, kv.x); m.y = fmaxf(m.y, kv.y); m.z = fmaxf(m.z, kv.z); m.w = fmaxf(m.w, kv.w);
            *reinterpret_cast<float4*>(&vs[r][lc4]) = vv;
        }
        // chunk column maximum: lanes with equal (lane & 7) hold the same 4 channels
#pragma unroll
        for (int o = 8; o < 32; o <<= 1) {
            m.x = fmaxf(m.x, __shfl_xor_sync(0xffffffffu, m.x, o)); m.y = fmaxf(m.y, __shfl_xor_sync(0xffffffffu, m.y, o));
            m.z = fmaxf(m.z, __shfl_xor_sync(0xffffffffu, m.z, o)); m.w = fmaxf(m.w, __shfl_xor_sync(0xffffffffu, m.w, o));
        }
        if (lane < 8) { red[warp][lc4] = m.x; red[warp][lc4 + 1] = m.y; red[warp][lc4 + 2] = m.z; red[warp][lc4 + 3] = m.w; }
        __syncthreads();
        if (tid < LA_D) {
            float t = red[0][tid];
#pragma unroll
            for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w][tid]);
            const float mo = kmax[tid], mn = fmaxf(mo, t);
            rescale[tid] = (mo == mn) ? 1.0f : expf(mo - mn);   // first chunk: exp(-inf) = 0 on zero accumulators
            kmax[tid] = mn;
        }
        __syncthreads();
        {
            const float4 km = *reinterpret_cast<const float4*>(&kmax[lc4]);
#pragma unroll
            for (int rr = 0; rr < LA_CHUNK / 32; ++rr) {
                const int r = rr * 32 + lr;
                const float4 kv = kraw[rr];   // rows >= cnt hold -inf -> exp = 0: they add nothing
                *reinterpret_cast<float4*>(&ek[r][lc4]) = make_float4(expf(kv.x - km.x), expf(kv.y - km.y), expf(kv.z - km.z), expf(kv.w - km.w));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float sc = rescale[d0 + i];
                ksum[i] *= sc;
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] *= sc;
            }
        }
        __syncthreads();
#pragma unroll 8
        for (int r = ng * (LA_CHUNK / 4); r < (ng + 1) * (LA_CHUNK / 4); ++r) {
            const float4 a = *reinterpret_cast<const float4*>(&ek[r][d0]);
            const float4 vv = *reinterpret_cast<const float4*>(&vs[r][e0]);
            const float av[4] = {a.x, a.y, a.z, a.w}, vv4[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ksum[i] += av[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], vv4[j], acc[i][j]);
            }
        }
    }
    // reduce the 4 pixel groups through shared memory (reusing the staging tiles)
    __syncthreads();
    float* racc = &ek[0][0];   // [4][64][16] floats = 16 KB
    float* rsum = &vs[0][0];   // [4][64][4]
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rsum[(ng * 64 + t64) * 4 + i] = ksum[i];
#pragma unroll
        for (int j = 0; j < 4; ++j) racc[(ng * 64 + t64) * 16 + i * 4 + j] = acc[i][j];
    }
    __syncthreads();
    if (ng == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float ks = 0.f, o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int gq = 0; gq < 4; ++gq) {
                ks += rsum[(gq * 64 + t64) * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] += racc[(gq * 64 + t64) * 16 + i * 4 + j];
            }
            const float inv = 1.0f / ks;
            float* cb = ctx + (int64_t)blockIdx.x * LA_CTX;
            *reinterpret_cast<float4*>(cb + (d0 + i) * LA_D + e0) = make_float4(o[0] * inv, o[1] * inv, o[2] * inv, o[3] * inv);
            if (e0 == 0) { cb[LA_D * LA_D + d0 + i] = kmax[d0 + i]; cb[LA_D * LA_D + LA_D + d0 + i] = ks; }
        }
    }
}

// ---- FP16 k | v (fused inference path): the same context on mma.sync.m16n8k16 (fp16 operands, fp32 accumulate) ----
// The scalar kernel above is bound by shared-memory operand reads (two LDS.128 per 16 FMA: 0.7 ms per 16x128-level call at
// B = 1024 for 1 GB of fp16 k | v).  Here ctx[d, e] = sum_n ek[n, d] v[n, e] is a [32 x n] x [n x 32] product with the pixel axis as
// K: both operands sit in shared memory pixel-major ([n][32] halves, 80-byte rows -> conflict-free ldmatrix), and
// ldmatrix.trans delivers the A fragment (ek^T) and the B fragment (v) directly.  grid = B * heads, 128 threads; per 128-pixel
// chunk a thread converts 4 rows x 8 channels (always the same channels: their running sums stay in registers), the four
// warps each multiply a 32-pixel quarter, the 32 x 32 partial contexts are reduced through shared memory at the end.
// ek = exp(k - max_d + 8 ln 2): the factor 256 keeps small weights out of the fp16 subnormal range and cancels in ctx = acc / ksum.
constexpr int LC_LD = 40;
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_f16_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
    const __half2 r = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
    return *reinterpret_cast<const uint32_t*>(&r);
}
// Round 2: the four warps of a CTA are INDEPENDENT streams.  Warp w takes the 32-pixel blocks w, w + 4, ... of the (sample, head)
// and keeps its own running column maximum, channel sums and 32 x 32 accumulator; the lanes that loaded a block are the lanes whose
// MMAs consume it (rows of the block = rows of the warp's shared-memory tile), so the loop needs no block barrier at all -- the
// previous version met four __syncthreads per 128 pixels (maximum exchange, rescale factors, tiles written, tiles consumed) and was
// latency bound at 0.57 of the copy peak inside the step.  k rows of the next block are requested into registers and the next v
// block goes global -> shared (cp.async, double buffered) before the MMAs of the current one; accumulators are only rescaled when a
// column maximum actually rose (warp vote).  The four partial contexts are merged at the end like split-K softmax partials.
__global__ void __launch_bounds__(128) linattn_context_mma_kernel(const __half* __restrict__ kbase, const __half* __restrict__ vbase,
                                                                  int ld, float* __restrict__ ctx, int n) {
    constexpr int TILE = 32 * LC_LD;                            // halves per 32-pixel tile (80-byte rows: conflict-free ldmatrix)
    __shared__ __align__(16) __half tiles[4 * 3 * TILE];        // per warp: ek | v (two buffers); reused for the merge (16 KB of floats)
    __shared__ float wres[4][LA_D];                             // per warp: rescale factors of the current block
    __shared__ float wmax[4][LA_D], wsum[4][LA_D];
    const int b = blockIdx.x / LA_HEADS, h = blockIdx.x % LA_HEADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const __half* kp = kbase + (int64_t)b * n * ld + h * LA_D;
    const __half* vp = vbase + (int64_t)b * n * ld + h * LA_D;
    const int cg = lane & 3, pr = lane >> 2;   // loader role: channels 8 cg .. 8 cg + 7 of block rows pr, pr + 8, pr + 16, pr + 24
    __half* ek = tiles + warp * 3 * TILE;
    const uint32_t ek_s = (uint32_t)__cvta_generic_to_shared(ek), vs_s = ek_s + 2u * TILE;
    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][jj][k] = 0.f;
    float ksum[8], kmx[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) { ksum[jj] = 0.f; kmx[jj] = -INFINITY; }
    constexpr float LOG2E = 1.4426950408889634f;
    const int blocks = (n + 31) / 32;
    uint4 kr[4];
    auto load_k = [&](int blk) {
        const int n0 = blk * 32;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int r = n0 + pr + 8 * rr;
            kr[rr] = make_uint4(0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u, 0xFC00FC00u);   // -inf: exp = 0, rows past the end add nothing
            if (r < n) kr[rr] = __ldcs(reinterpret_cast<const uint4*>(kp + (int64_t)r * ld + cg * 8));
        }
    };
    auto load_v = [&](int blk, int buf) {
        const int n0 = blk * 32;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            const int rl = pr + 8 * rr, r = n0 + rl;
            const __half* src = vp + (int64_t)min(r, n - 1) * ld + cg * 8;   // clamped address, 0 source bytes past the end
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(vs_s + 2u * (uint32_t)(buf * TILE + rl * LC_LD + cg * 8)), "l"(src),
                         "r"(r < n ? 16 : 0) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (warp < blocks) { load_k(warp); load_v(warp, 0); }
    int it = 0;
    for (int blk = warp; blk < blocks; blk += 4, ++it) {
        // column maxima of this block: over the thread's four rows, then over the eight row groups of the warp
        uint32_t m2[4] = {kr[0].x, kr[0].y, kr[0].z, kr[0].w};
#pragma unroll
        for (int rr = 1; rr < 4; ++rr) {
            m2[0] = hmax2_u32(m2[0], kr[rr].x); m2[1] = hmax2_u32(m2[1], kr[rr].y);
            m2[2] = hmax2_u32(m2[2], kr[rr].z); m2[3] = hmax2_u32(m2[3], kr[rr].w);
        }
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) m2[jj] = hmax2_u32(m2[jj], __shfl_xor_sync(0xffffffffu, m2[jj], o));
        }
        float mn[8];
        bool rose = false;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&m2[jj]));
            mn[2 * jj] = fmaxf(kmx[2 * jj], f.x);
            mn[2 * jj + 1] = fmaxf(kmx[2 * jj + 1], f.y);
            rose |= mn[2 * jj] != kmx[2 * jj] || mn[2 * jj + 1] != kmx[2 * jj + 1];
        }
        if (__any_sync(0xffffffffu, rose)) {   // warp-uniform: every lane with the same cg holds the same maxima
            float sc[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                sc[jj] = (kmx[jj] == mn[jj]) ? 1.0f : __expf(kmx[jj] - mn[jj]);   // first block: exp(-inf) = 0 on zero accumulators
                ksum[jj] *= sc[jj];
                kmx[jj] = mn[jj];
            }
            if (pr == 0) {
                *reinterpret_cast<float4*>(&wres[warp][cg * 8]) = make_float4(sc[0], sc[1], sc[2], sc[3]);
                *reinterpret_cast<float4*>(&wres[warp][cg * 8 + 4]) = make_float4(sc[4], sc[5], sc[6], sc[7]);
            }
            __syncwarp();
            const int g = lane >> 2;
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float s0 = wres[warp][mt * 16 + g], s1 = wres[warp][mt * 16 + g + 8];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) { acc[mt][nt][0] *= s0; acc[mt][nt][1] *= s0; acc[mt][nt][2] *= s1; acc[mt][nt][3] *= s1; }
            }
            __syncwarp();
        }
        {
            float off[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) off[jj] = fmaf(-kmx[jj], LOG2E, 8.0f);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                float f[8];
                unpack8(kr[rr], f);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) f[jj] = ex2_approx(fmaf(f[jj], LOG2E, off[jj]));
                const uint4 e8 = pack8(f);
                unpack8(e8, f);   // the sums use the values the tensor cores see
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) ksum[jj] += f[jj];
                *reinterpret_cast<uint4*>(ek + (pr + 8 * rr) * LC_LD + cg * 8) = e8;
            }
        }
        const int buf = it & 1;
        if (blk + 4 < blocks) {   // the warp's next block: k rows into the registers just consumed, v into the other buffer
            load_k(blk + 4);
            load_v(blk + 4, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int k0 = ks * 16;
            uint32_t a[2][4], bb[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                ldsm_x4_trans(a[mt], ek_s + 2u * (uint32_t)((k0 + (lane & 7) + ((lane >> 4) << 3)) * LC_LD + mt * 16 + (((lane >> 3) & 1) << 3)));
#pragma unroll
            for (int np = 0; np < 2; ++np)
                ldsm_x4_trans(bb[np], vs_s + 2u * (uint32_t)(buf * TILE + (k0 + (lane & 7) + (((lane >> 3) & 1) << 3)) * LC_LD + np * 16 + ((lane >> 4) << 3)));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_f16_16x8x16(acc[mt][nt], a[mt], bb[nt >> 1][(nt & 1) * 2], bb[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncwarp();   // the tile is rewritten by other lanes in the next iteration
    }
    // merge the four warps' partials: maxima, channel sums, contexts
    __syncthreads();
    float* racc = reinterpret_cast<float*>(tiles);   // [4][32][32] floats = 16 KB (the tile memory holds 30 KB)
    {
        const int g = lane >> 2, c2 = 2 * (lane & 3);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float* o = racc + warp * 1024 + (mt * 16 + g) * LA_D + nt * 8 + c2;
                *reinterpret_cast<float2*>(o) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
                *reinterpret_cast<float2*>(o + 8 * LA_D) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
            }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            float t = ksum[jj];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            ksum[jj] = t;
        }
        if (pr == 0) {
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) { wsum[warp][cg * 8 + jj] = ksum[jj]; wmax[warp][cg * 8 + jj] = kmx[jj]; }
        }
    }
    __syncthreads();
    if (tid < LA_D) {   // per channel d: global maximum and the factor that brings each warp's partial to it (0 for a warp without blocks)
        const float M = fmaxf(fmaxf(wmax[0][tid], wmax[1][tid]), fmaxf(wmax[2][tid], wmax[3][tid]));
        float ks = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const float f = wmax[w][tid] == -INFINITY ? 0.f : __expf(wmax[w][tid] - M);
            wres[w][tid] = f;
            ks = fmaf(wsum[w][tid], f, ks);
        }
        wmax[0][tid] = M;
        wsum[0][tid] = ks;
    }
    __syncthreads();
    float* cb = ctx + (int64_t)blockIdx.x * LA_CTX;
    for (int i = tid; i < LA_D * LA_D; i += 128) {
        const int d = i >> 5;
        const float t = fmaf(racc[i], wres[0][d], racc[1024 + i] * wres[1][d]) + fmaf(racc[2048 + i], wres[2][d], racc[3072 + i] * wres[3][d]);
        cb[i] = t / wsum[0][d];
    }
    if (tid < LA_D) {
        cb[LA_D * LA_D + tid] = wmax[0][tid];
        cb[LA_D * LA_D + LA_D + tid] = wsum[0][tid] * (1.0f / 256.0f);
    }
}

// out[n, h*32+e] = 32^-0.5 * sum_d ctx[d,e] * softmax_d(q[n,:])[d]
// One CTA = P consecutive pixels of one sample x all 4 heads, 256 threads = 8 warps, warp w -> head w & 3, pixel tiles of 16.
// The q rows (512 bytes per pixel) are staged through shared memory with coalesced 16-byte loads; the per-head
// [16 px x 32] x [32 x 32] products run on mma.sync.m16n8k8 TF32 (operands rounded to nearest first, fp32 accumulate) --
// a K = N = 32 contraction is far too small for a tcgen05 tile, and the scalar-FMA version of this kernel was bound by
// shared-memory operand reads (3.3 ms per step at B = 1024).  In the A-fragment layout a pixel's 32 logits sit in the
// 4 lanes of a quad, so softmax_d costs two shuffles.  Results leave through the same staging tile (256 / 512 bytes
// per pixel, coalesced).  Strides: q tile 132 floats, context 40 floats per d-row (both conflict free for the fragments).
constexpr int LA_QLD = LA_HID + 4;
constexpr int LA_CLD = 40;
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int P, typename T>
__global__ void __launch_bounds__(256) linattn_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ ctx,
                                                            T* __restrict__ out, int n) {
    extern __shared__ __align__(16) float sm_apply[];
    float* cs = sm_apply;                                  // [4][32][LA_CLD] contexts of this sample (TF32-rounded)
    float* qs = sm_apply + LA_HEADS * LA_D * LA_CLD;       // [P][LA_QLD] q rows, later the output rows
    const int tiles = n / P;
    const int b = blockIdx.x / tiles, p0 = (blockIdx.x % tiles) * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < LA_HEADS * LA_D * LA_D; i += 256) {
        const int hh = i >> 10, d = (i >> 5) & 31, e = i & 31;
        cs[(hh * LA_D + d) * LA_CLD + e] = to_tf32(ctx[((int64_t)b * LA_HEADS + hh) * LA_CTX + (i & 1023)]);
    }
    const float* qbase = qkv + ((int64_t)b * n + p0) * LA_QKV;
    for (int idx = tid; idx < P * 32; idx += 256) {        // P pixels x 32 float4
        const int px = idx >> 5, c4 = (idx & 31) * 4;
        *reinterpret_cast<float4*>(qs + px * LA_QLD + c4) = __ldcs(reinterpret_cast<const float4*>(qbase + (int64_t)px * LA_QKV + c4));
    }
    __syncthreads();
    const int h = warp & 3, g = lane >> 2, t = lane & 3;
    // B fragments of this head's context: b0 = ctx[8k + t][8j + g], b1 = ctx[8k + t + 4][8j + g]
    uint32_t bf[4][4][2];
    const float* ch = cs + h * LA_D * LA_CLD;
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bf[k][j][0] = __float_as_uint(ch[(8 * k + t) * LA_CLD + 8 * j + g]);
            bf[k][j][1] = __float_as_uint(ch[(8 * k + t + 4) * LA_CLD + 8 * j + g]);
        }
    float acc[P / 32][4][4];   // this warp's pixel tiles (tile = (warp >> 2) + 2 * i) x 4 n-tiles x C fragment
#pragma unroll
    for (int i = 0; i < P / 32; ++i) {
        const int r0 = ((warp >> 2) + 2 * i) * 16;
        // A fragment source: rows r0 + g and r0 + g + 8, logits d = 8k + t and 8k + t + 4
        float qa[2][8];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                qa[rr][2 * k] = qs[(r0 + g + 8 * rr) * LA_QLD + h * LA_D + 8 * k + t];
                qa[rr][2 * k + 1] = qs[(r0 + g + 8 * rr) * LA_QLD + h * LA_D + 8 * k + t + 4];
            }
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            float mx = qa[rr][0];
#pragma unroll
            for (int k = 1; k < 8; ++k) mx = fmaxf(mx, qa[rr][k]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float den = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) { qa[rr][k] = expf(qa[rr][k] - mx); den += qa[rr][k]; }
            den += __shfl_xor_sync(0xffffffffu, den, 1);
            den += __shfl_xor_sync(0xffffffffu, den, 2);
            const float sc = 0.17677669529663687f / den;  // 32^-0.5 / sum
#pragma unroll
            for (int k = 0; k < 8; ++k) qa[rr][k] = to_tf32(qa[rr][k] * sc);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t a[4] = {__float_as_uint(qa[0][2 * k]), __float_as_uint(qa[1][2 * k]), __float_as_uint(qa[0][2 * k + 1]),
                                   __float_as_uint(qa[1][2 * k + 1])};
#pragma unroll
            for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[i][j], a, bf[k][j][0], bf[k][j][1]);
        }
    }
    __syncthreads();   // all q reads done: the tile becomes the output staging
    T* os = reinterpret_cast<T*>(qs);
    constexpr int OLD = LA_QLD * (int)(sizeof(float) / sizeof(T));   // same 528-byte row pitch in elements of T
#pragma unroll
    for (int i = 0; i < P / 32; ++i) {
        const int r0 = ((warp >> 2) + 2 * i) * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // C fragment: (row g, cols 2t, 2t+1), (row g + 8, cols 2t, 2t+1) of n-tile j
            T* o0 = os + (r0 + g) * OLD + h * LA_D + 8 * j + 2 * t;
            T* o1 = os + (r0 + g + 8) * OLD + h * LA_D + 8 * j + 2 * t;
            o0[0] = to_operand(acc[i][j][0], T()); o0[1] = to_operand(acc[i][j][1], T());
            o1[0] = to_operand(acc[i][j][2], T()); o1[1] = to_operand(acc[i][j][3], T());
        }
    }
    __syncthreads();
    // coalesced copy-out: P rows of 128 * sizeof(T) bytes
    constexpr int V = 16 / (int)sizeof(T);          // elements per 16-byte vector
    constexpr int VPR = LA_HID / V;                 // vectors per row
    T* obase = out + ((int64_t)b * n + p0) * LA_HID;
    for (int idx = tid; idx < P * VPR; idx += 256) {
        const int r = idx / VPR, c = (idx % VPR) * V;
        *reinterpret_cast<uint4*>(obase + (int64_t)r * LA_HID + c) = *reinterpret_cast<const uint4*>(os + r * OLD + c);
    }
}

// Folded output projection of LinearAttention: out_proj[n, co] = sum_{h,e} W[co, 32h + e] * sum_d ctx_h[d, e] qs_h[n, d]
//                                                            = sum_{h,d} qs_h[n, d] * Wf_b[co, 32h + d],
// Wf_b[co, 32h + d] = sum_e W[co, 32h + e] ctx_{b,h}[d, e]: a per-sample [Cout, 128] weight.  The context apply then IS the 1x1
// output projection (one tcgen05 GEMM with per-sample weights) and the [B*n, 128] attention tensor is never materialised.
// grid = B, 256 threads: the sample's four contexts are loaded once, then 16 output channels per iteration; thread -> column
// hd = tid & 127 (its 32 context values live in registers) and 8 of the slab's 16 output channels.  Shared memory is kept at
// 25 KB so that all B = 1024 CTAs are resident in one wave (with 32-channel slabs, 33 KB, 6 CTAs per SM fitted and a second,
// nearly empty wave doubled the time of this latency-bound kernel).
template <typename T>
__global__ void __launch_bounds__(256) linattn_fold_kernel(const float* __restrict__ ws, const float* __restrict__ w_out,
                                                           T* __restrict__ wf, int Cout) {
    __shared__ float cs[LA_HEADS * LA_D][LA_D + 1];   // ctx[h*32 + d][e], padded: lanes walk d
    __shared__ __align__(16) float wsm[16][LA_HID];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < LA_HEADS * LA_D * LA_D; i += 256)
        cs[i >> 5][i & 31] = ws[((int64_t)b * LA_HEADS + (i >> 10)) * LA_CTX + (i & 1023)];
    __syncthreads();
    const int hd = threadIdx.x & 127, h = hd >> 5, half_id = threadIdx.x >> 7;
    float c[LA_D];
#pragma unroll
    for (int e = 0; e < LA_D; ++e) c[e] = cs[hd][e];
    for (int co0 = 0; co0 < Cout; co0 += 16) {
        __syncthreads();
        for (int i = threadIdx.x; i < 16 * LA_HID / 4; i += 256)
            reinterpret_cast<float4*>(&wsm[0][0])[i] = __ldg(reinterpret_cast<const float4*>(w_out + (int64_t)co0 * LA_HID) + i);
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < 8; ++r) {
            const int co = half_id * 8 + r;
            float acc = 0.f;
#pragma unroll
            for (int e = 0; e < LA_D; e += 4) {
                const float4 wv = *reinterpret_cast<const float4*>(&wsm[co][h * LA_D + e]);   // same address across a warp: broadcast
                acc = fmaf(wv.x, c[e], acc); acc = fmaf(wv.y, c[e + 1], acc); acc = fmaf(wv.z, c[e + 2], acc); acc = fmaf(wv.w, c[e + 3], acc);
            }
            wf[((int64_t)b * Cout + co0 + co) * LA_HID + hd] = to_operand(acc, T());
        }
    }
}

// FP16-mode fold on tensor cores: per (sample, head) D[C x 32] = W[:, 32h : 32h + 32] (C x 32) . ctx_{b,h}^T (32 x 32) with
// mma.sync.m16n8k16, both operands split into fp16 high + low parts (hi*hi + hi*lo + lo*hi: ~2^-21 relative, the accuracy of the
// fp32 CUDA-core version above, whose 16-channel slabs with two block barriers each made it latency bound: 0.56 ms per step for
// 6 GFMA).  CTA = (sample, 128 output channels), 8 warps x 16 rows; the sample's four contexts sit in shared memory.
__device__ __forceinline__ void split_half2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__global__ void __launch_bounds__(256) linattn_fold_mma_kernel(const float* __restrict__ ws, const float* __restrict__ w_out,
                                                               __half* __restrict__ wf, int Cout) {
    __shared__ float cs[LA_HEADS][LA_D][LA_D + 8];   // ctx[h][d][e]; rows padded to 40 floats: 8-byte aligned, conflict-free b-fragment reads
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < LA_HEADS * LA_D * LA_D; i += 256)
        cs[i >> 10][(i >> 5) & 31][i & 31] = ws[((int64_t)b * LA_HEADS + (i >> 10)) * LA_CTX + (i & 1023)];
    __syncthreads();
    const int row0 = blockIdx.y * 128 + warp * 16;
    if (row0 >= Cout) return;
    const int g = lane >> 2, t4 = lane & 3;
    const float* wr0 = w_out + (int64_t)(row0 + g) * LA_HID;
    const float* wr1 = wr0 + 8 * LA_HID;
    __half* o0 = wf + ((int64_t)b * Cout + row0 + g) * LA_HID;
    __half* o1 = o0 + 8 * LA_HID;
#pragma unroll 1
    for (int h = 0; h < LA_HEADS; ++h) {
        float d[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            // A fragment (W rows g, g + 8; k = e): a0 = (g, 2t..2t+1), a1 = (g+8, 2t..), a2 = (g, 2t+8..), a3 = (g+8, 2t+8..)
            const int e0 = h * LA_D + ks * 16 + 2 * t4;
            const float2 w00 = __ldg(reinterpret_cast<const float2*>(wr0 + e0)), w10 = __ldg(reinterpret_cast<const float2*>(wr1 + e0));
            const float2 w01 = __ldg(reinterpret_cast<const float2*>(wr0 + e0 + 8)), w11 = __ldg(reinterpret_cast<const float2*>(wr1 + e0 + 8));
            uint32_t ah[4], al[4];
            split_half2(w00.x, w00.y, ah[0], al[0]);
            split_half2(w10.x, w10.y, ah[1], al[1]);
            split_half2(w01.x, w01.y, ah[2], al[2]);
            split_half2(w11.x, w11.y, ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                // B fragment (k = e, n = d): b0 = (k 2t..2t+1, n g), b1 = (k 2t+8.., n g); B[k][n] = ctx[d = n][e = k]
                const float* cr = &cs[h][nt * 8 + g][ks * 16 + 2 * t4];
                const float2 c0 = *reinterpret_cast<const float2*>(cr), c1 = *reinterpret_cast<const float2*>(cr + 8);
                uint32_t bh0, bl0, bh1, bl1;
                split_half2(c0.x, c0.y, bh0, bl0);
                split_half2(c1.x, c1.y, bh1, bl1);
                mma16816(d[nt], al, bh0, bh1);
                mma16816(d[nt], ah, bl0, bl1);
                mma16816(d[nt], ah, bh0, bh1);
            }
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int col = h * LA_D + nt * 8 + 2 * t4;
            *reinterpret_cast<__half2*>(o0 + col) = __floats2half2_rn(d[nt][0], d[nt][1]);
            *reinterpret_cast<__half2*>(o1 + col) = __floats2half2_rn(d[nt][2], d[nt][3]);
        }
    }
}

// full softmax attention for n <= 32 tokens: one warp per (b, head), lane = query token
template <typename T>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, T* __restrict__ out, int n) {
    __shared__ float ks[LA_HEADS][32][LA_D + 1];
    __shared__ float vs[LA_HEADS][32][LA_D + 1];
    const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* base = qkv + (int64_t)b * n * LA_QKV;
    for (int j = 0; j < n; ++j) {
        ks[h][j][lane] = base[(int64_t)j * LA_QKV + LA_HID + h * LA_D + lane];
        vs[h][j][lane] = base[(int64_t)j * LA_QKV + 2 * LA_HID + h * LA_D + lane];
    }
    __syncwarp();
    if (lane < n) {
        const float scale = 0.17677669529663687f;
        float q[LA_D];
#pragma unroll
        for (int d = 0; d < LA_D; ++d) q[d] = base[(int64_t)lane * LA_QKV + h * LA_D + d] * scale;
        float sim[32];
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float s = -INFINITY;
            if (j < n) {
                s = 0.f;
#pragma unroll
                for (int d = 0; d < LA_D; ++d) s = fmaf(q[d], ks[h][j][d], s);
            }
            sim[j] = s;
            mx = fmaxf(mx, s);
        }
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { sim[j] = (j < n) ? expf(sim[j] - mx) : 0.f; den += sim[j]; }
        const float inv = 1.0f / den;
        T* o = out + ((int64_t)b * n + lane) * LA_HID + h * LA_D;
#pragma unroll
        for (int d = 0; d < LA_D; ++d) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) a = fmaf(sim[j], vs[h][j][d], a);
            o[d] = to_operand(a * inv, T());
        }
    }
}

// ---------------------------------------------------------------------------------------------- misc
__global__ void upsample2x_kernel(const float4* __restrict__ x, float4* __restrict__ y, int64_t total4, int H, int W, int c4) {
    // total4 = B*2H*2W*c4 output float4s
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4);
        int64_t pix = i / c4;
        const int wo = (int)(pix % (2 * W)); pix /= (2 * W);
        const int ho = (int)(pix % (2 * H));
        const int64_t b = pix / (2 * H);
        y[i] = x[((b * H + (ho >> 1)) * W + (wo >> 1)) * c4 + c];
    }
}

// 8 lanes per pixel (a warp covers 4 consecutive pixels), Cout <= 4 dot products of length Cin; NHWC -> NCHW
template <typename T>
__global__ void __launch_bounds__(256) head_conv1_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, float* __restrict__ out, int64_t M,
                                                         int HW, int Cin, int Cout) {
    extern __shared__ float wsm[];   // [Cout][Cin]
    for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) wsm[i] = w[i];
    __syncthreads();
    const int sub = threadIdx.x & 7;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; row < M; row += ((int64_t)gridDim.x * blockDim.x) >> 3) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = sub * 4; c < Cin; c += 32) {
            const float4 xv = load4(x + row * Cin + c);
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                if (o < Cout) {
                    const float4 wv = *reinterpret_cast<const float4*>(wsm + o * Cin + c);
                    acc[o] += (xv.x * wv.x + xv.y * wv.y) + (xv.z * wv.z + xv.w * wv.w);
                }
            }
        }
        const int64_t b = row / HW, p = row % HW;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            if (o < Cout) {
                float s = acc[o];
                s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
                if (sub == 0) out[(b * Cout + o) * HW + p] = s + (bias ? bias[o] : 0.f);
            }
        }
    }
}

// Fast path of the head for the dim-128 model in FP16 mode (Cin = 128, Cout = 3): 4 lanes per pixel, a lane owns the four
// 8-channel vectors j, j + 4, j + 8, j + 12 of the pixel row (a warp's load instruction covers 8 pixels x 64 contiguous bytes) and
// keeps its 3 x 32 weights in registers; 6 shuffles per 32 elements (the generic kernel above: 9 per 16 and one shared load per
// 4 FMAs -- 0.36 of the copy peak).  Two pixels per lane in flight.
__global__ void __launch_bounds__(128) head_conv1_h128_kernel(const __half* __restrict__ x, const float* __restrict__ w,
                                                              const float* __restrict__ bias, float* __restrict__ out, int64_t M,
                                                              int HW) {
    const int lane = threadIdx.x & 31, j = lane & 3;
    float wr[3][32];
#pragma unroll
    for (int o = 0; o < 3; ++o)
#pragma unroll
        for (int v = 0; v < 4; ++v)
#pragma unroll
            for (int k = 0; k < 8; ++k) wr[o][v * 8 + k] = __ldg(w + o * 128 + (v * 4 + j) * 8 + k);
    const float b0 = bias ? bias[0] : 0.f, b1 = bias ? bias[1] : 0.f, b2 = bias ? bias[2] : 0.f;
    const int64_t gw = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t p0 = gw * 16; p0 < M; p0 += nw * 16) {   // a warp: 2 x 8 pixels per iteration
        uint4 xv[2][4];
        int64_t row[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            row[u] = p0 + u * 8 + (lane >> 2);
#pragma unroll
            for (int v = 0; v < 4; ++v)
                xv[u][v] = row[u] < M ? __ldcs(reinterpret_cast<const uint4*>(x + row[u] * 128) + v * 4 + j) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                float f[8];
                unpack8(xv[u][v], f);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    a0 = fmaf(f[k], wr[0][v * 8 + k], a0);
                    a1 = fmaf(f[k], wr[1][v * 8 + k], a1);
                    a2 = fmaf(f[k], wr[2][v * 8 + k], a2);
                }
            }
#pragma unroll
            for (int o = 1; o < 4; o <<= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                a2 += __shfl_xor_sync(0xffffffffu, a2, o);
            }
            if (row[u] < M && j < 3) {   // lane j of the quad writes output channel j
                const int64_t b = row[u] / HW, p = row[u] % HW;
                out[(b * 3 + j) * HW + p] = j == 0 ? a0 + b0 : (j == 1 ? a1 + b1 : a2 + b2);
            }
        }
    }
}

__device__ __forceinline__ float act_in(float v, int act) {
    if (act == 1) return silu(v);
    if (act == 2) return 0.5f * v * (1.0f + erff(v * 0.7071067811865476f));
    return v;
}
// warp per output element (r, n)
__global__ void __launch_bounds__(256) linear_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ b, float* __restrict__ y, int K, int N, int act) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), r = blockIdx.y;
    if (n >= N) return;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(act_in(x[(int64_t)r * K + k], act), w[(int64_t)n * K + k], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[(int64_t)r * N + n] = acc + (b ? b[n] : 0.f);
}

__global__ void sinusoidal_kernel(const float* __restrict__ t, float* __restrict__ emb, int R, int dim, float neg_step) {
    const int half = dim / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * half; i += gridDim.x * blockDim.x) {
        const int r = i / half, k = i % half;
        const float f = expf((float)k * neg_step);
        const float a = t[r] * f;
        emb[(int64_t)r * dim + k] = sinf(a);
        emb[(int64_t)r * dim + half + k] = cosf(a);
    }
}

}  // namespace sdc

using namespace sdc;

static inline unsigned blocks_for(int64_t n, int per) { int64_t b = (n + per - 1) / per; return (unsigned)(b > 148 * 16 ? 148 * 16 : (b < 1 ? 1 : b)); }

#define SDC_CHECK_PREC(name) SDC_REQUIRE(prec == SDC_PREC_TF32 || prec == SDC_PREC_F16, name ": precision %d", prec)

extern "C" int sdc_pack_conv_weight(int prec, int kind, const float* w, void* wp, int Cout, int Cin, void* stream) {
    SDC_CHECK_PREC("pack_conv_weight");
    SDC_REQUIRE(kind >= 0 && kind <= 3 && w && wp && Cout > 0 && Cin > 0, "pack_conv_weight: bad arguments");
    SDC_REQUIRE(kind != 2 || Cin % 4 == 0, "pack_conv_weight: unshuffle conv needs Cin %% 4 == 0");
    if (kind == 3) {
        const int64_t tot = (int64_t)16 * Cout * Cin;
        if (prec == SDC_PREC_F16)
            pack_upconv_weight_kernel<__half><<<blocks_for(tot, 256), 256, 0, as_stream(stream)>>>(w, (__half*)wp, Cout, Cin);
        else
            pack_upconv_weight_kernel<float><<<blocks_for(tot, 256), 256, 0, as_stream(stream)>>>(w, (float*)wp, Cout, Cin);
        SDC_LAUNCHED();
        return SDC_OK;
    }
    const int64_t total = (int64_t)Cout * Cin * (kind == 1 ? 9 : 1);
    if (prec == SDC_PREC_F16)
        pack_conv_weight_kernel<__half><<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(kind, w, (__half*)wp, Cout, Cin);
    else
        pack_conv_weight_kernel<float><<<blocks_for(total, 256), 256, 0, as_stream(stream)>>>(kind, w, (float*)wp, Cout, Cin);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_stem_conv7(int prec, const float* x, const float* w, const float* bias, void* out, int B, int Cin, int H, int W,
                              int Cout, void* stream) {
    SDC_CHECK_PREC("stem_conv7");
    SDC_REQUIRE(x && w && out && B > 0, "stem_conv7: bad arguments");
    SDC_REQUIRE(W % 8 == 0 && Cout % 8 == 0 && (W / 8) * (Cout / 8) <= 1024 && (W / 8) * (Cout / 8) >= 32,
                "stem_conv7: unsupported W=%d Cout=%d", W, Cout);
    const int threads = (W / 8) * (Cout / 8);
    const size_t smem = ((size_t)Cin * 49 * Cout + (size_t)Cin * 7 * (W + 6)) * sizeof(float);
    SDC_REQUIRE(smem <= 227 * 1024, "stem_conv7: weights do not fit shared memory (%zu bytes)", smem);
    dim3 grid((unsigned)B, (unsigned)((H + STEM_ROWS - 1) / STEM_ROWS));
    if (prec == SDC_PREC_F16) {
        SDC_CUDA(cudaFuncSetAttribute(stem_conv7_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_conv7_kernel<__half><<<grid, threads, smem, as_stream(stream)>>>(x, w, bias, (__half*)out, Cin, H, W, Cout);
    } else {
        SDC_CUDA(cudaFuncSetAttribute(stem_conv7_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        stem_conv7_kernel<float><<<grid, threads, smem, as_stream(stream)>>>(x, w, bias, (float*)out, Cin, H, W, Cout);
    }
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_stem_im2col(int prec, const float* x, void* a, int B, int Cin, int H, int W, int kp, void* stream) {
    SDC_CHECK_PREC("stem_im2col");
    SDC_REQUIRE(x && a && B > 0 && Cin > 0 && H > 0 && W > 0, "stem_im2col: bad arguments");
    SDC_REQUIRE(kp % 16 == 0 && kp / 2 >= Cin * 49 && kp / 8 <= 256, "stem_im2col: kp=%d must be a multiple of 16, >= 2*Cin*49, <= 2048", kp);
    const size_t sm = (size_t)Cin * 7 * (W + 6) * sizeof(float);
    SDC_REQUIRE(sm <= 48 * 1024, "stem_im2col: row window does not fit shared memory");
    if (prec == SDC_PREC_F16)
        stem_im2col_kernel<__half><<<(unsigned)(B * H), 256, sm, as_stream(stream)>>>(x, (__half*)a, Cin, H, W, kp);
    else
        stem_im2col_kernel<float><<<(unsigned)(B * H), 256, sm, as_stream(stream)>>>(x, (float*)a, Cin, H, W, kp);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_gn_silu(int prec, const void* xv, int x_operand, const double* stats, const float* gamma, const float* beta,
                           const float* scale_shift, const int32_t* t_index, int64_t ss_stride, const void* residual,
                           int residual_operand, void* y, int B, int HW, int C, void* stream) {
    SDC_CHECK_PREC("gn_silu");
    const float* x = (const float*)xv;
    SDC_REQUIRE(x && stats && gamma && beta && y && B > 0 && HW > 0, "gn_silu: bad arguments");
    SDC_REQUIRE(C % 4 == 0 && C <= 4096, "gn_silu: C=%d unsupported", C);
    SDC_REQUIRE(!x_operand || (prec == SDC_PREC_F16 && C % 8 == 0 && 256 % (C / 8) == 0 && (!residual || residual_operand)),
                "gn_silu: an fp16 input needs FP16 mode, C/8 dividing 256 and an fp16 residual");
    int ppc = HW;  // pixels per CTA: aim for >= 2 waves of CTAs without shrinking below 32 pixels
    while (ppc > 32 && (int64_t)B * (HW / ppc) < 2 * 148 && ppc % 2 == 0) ppc /= 2;
    // large batches: one CTA per sample is 2.3 waves at B = 1024 (3 CTAs per SM); split while a CTA keeps >= 64K elements so that the
    // partial last wave weighs less (SDC_GN_SPLIT=0 disables, for A/B timing)
    static const bool split_waves = []() { const char* e = getenv("SDC_GN_SPLIT"); return !(e && e[0] == '0'); }();
    while (split_waves && (int64_t)B * (HW / ppc) < 8 * 3 * 148 && ppc % 64 == 0 && (int64_t)(ppc / 2) * C >= 65536) ppc /= 2;
    dim3 grid((unsigned)B, (unsigned)((HW + ppc - 1) / ppc));
    const size_t sm = 2 * C * sizeof(float);
    cudaStream_t st = as_stream(stream);
    if (x_operand) {
        gn_silu_h8_kernel<false><<<grid, 256, 0, st>>>((const __half*)xv, stats, gamma, beta, scale_shift, t_index, ss_stride,
                                                       (const __half*)residual, (__half*)y, HW, C, ppc, nullptr);
    } else if (prec == SDC_PREC_F16) {
        if (residual_operand)
            gn_silu_kernel<__half, __half><<<grid, 256, sm, st>>>(x, stats, gamma, beta, scale_shift, t_index, ss_stride,
                                                                   (const __half*)residual, (__half*)y, HW, C, ppc);
        else
            gn_silu_kernel<float, __half><<<grid, 256, sm, st>>>(x, stats, gamma, beta, scale_shift, t_index, ss_stride,
                                                                  (const float*)residual, (__half*)y, HW, C, ppc);
    } else {
        gn_silu_kernel<float, float><<<grid, 256, sm, st>>>(x, stats, gamma, beta, scale_shift, t_index, ss_stride,
                                                             (const float*)residual, (float*)y, HW, C, ppc);
    }
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_gn_silu_rowstats(const void* x, const double* stats, const float* gamma, const float* beta, const float* scale_shift,
                                    const int32_t* t_index, int64_t ss_stride, const void* residual, void* y, float* rowstats, int B,
                                    int HW, int C, void* stream) {
    SDC_REQUIRE(x && stats && gamma && beta && y && rowstats && B > 0 && HW > 0, "gn_silu_rowstats: bad arguments");
    SDC_REQUIRE((C == 128 || C == 256) && HW % 32 == 0, "gn_silu_rowstats: C=%d HW=%d unsupported (C = 128 or 256: a pixel row inside one warp; HW %% 32 == 0)", C, HW);
    int ppc = HW;
    while (ppc > 32 && (int64_t)B * (HW / ppc) < 2 * 148 && ppc % 2 == 0) ppc /= 2;
    while ((int64_t)B * (HW / ppc) < 8 * 3 * 148 && ppc % 64 == 0 && (int64_t)(ppc / 2) * C >= 65536) ppc /= 2;
    dim3 grid((unsigned)B, (unsigned)((HW + ppc - 1) / ppc));
    gn_silu_h8_kernel<true><<<grid, 256, 0, as_stream(stream)>>>((const __half*)x, stats, gamma, beta, scale_shift, t_index, ss_stride,
                                                                  (const __half*)residual, (__half*)y, HW, C, ppc, (float2*)rowstats);
    SDC_LAUNCHED();
    return SDC_OK;
}

namespace sdc {
// one warp per output row: Wg[co, c] = fp16(W[co, c] * g[c]); wsum[co] = sum_c float(Wg[co, c]) (the ROUNDED weights, so that
// r * (acc - mu * wsum) equals the projection of the normalised row exactly in terms of what the tensor core multiplies)
__global__ void pack_qkv_ln_kernel(const float* __restrict__ w, const float* __restrict__ g, __half* __restrict__ wp, float* __restrict__ wsum,
                                   int Cout, int Cin) {
    const int co = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (co >= Cout) return;
    float s = 0.f;
    for (int c = lane; c < Cin; c += 32) {
        const __half h = __float2half_rn(w[(int64_t)co * Cin + c] * g[c]);
        wp[(int64_t)co * Cin + c] = h;
        s += __half2float(h);
    }
    s = warp_sum(s);
    if (lane == 0) wsum[co] = s;
}
}  // namespace sdc

extern "C" int sdc_pack_qkv_ln(const float* w, const float* g, void* w_packed, float* wsum, int Cout, int Cin, void* stream) {
    SDC_REQUIRE(w && g && w_packed && wsum && Cout > 0 && Cin > 0, "pack_qkv_ln: bad arguments");
    pack_qkv_ln_kernel<<<(Cout + 7) / 8, 256, 0, as_stream(stream)>>>(w, g, (__half*)w_packed, wsum, Cout, Cin);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_gn_silu_head(const float* x, const double* stats, const float* gamma, const float* beta, const void* residual,
                                int residual_operand, const float* head_w, const float* head_b, float* out, int B, int HW, int C,
                                int Cout, void* stream) {
    SDC_REQUIRE(x && stats && gamma && beta && residual && head_w && out && B > 0 && HW > 0, "gn_silu_head: bad arguments");
    SDC_REQUIRE(C == 128 && Cout >= 1 && Cout <= 4 && HW % 32 == 0, "gn_silu_head: C=%d Cout=%d HW=%d unsupported (C = 128, Cout <= 4, HW %% 32 == 0)", C, Cout, HW);
    // pixels per CTA: at least ~8 waves of CTAs (3 resident per SM) so that the last, partial wave costs little -- one CTA per sample
    // at B = 1024 is 2.3 waves, 77% efficient, and this kernel is not purely bandwidth bound (0.65 of the copy peak, ncu)
    int ppc = HW;
    while (ppc > 32 && (int64_t)B * (HW / ppc) < 8 * 3 * 148 && ppc % 64 == 0) ppc /= 2;
    dim3 grid((unsigned)B, (unsigned)((HW + ppc - 1) / ppc));
    cudaStream_t st = as_stream(stream);
    if (residual_operand)
        gn_silu_head_kernel<__half><<<grid, 256, 0, st>>>(x, stats, gamma, beta, (const __half*)residual, head_w, head_b, out, HW, Cout, ppc);
    else
        gn_silu_head_kernel<float><<<grid, 256, 0, st>>>(x, stats, gamma, beta, (const float*)residual, head_w, head_b, out, HW, Cout, ppc);
    SDC_LAUNCHED();
    return SDC_OK;
}

template <typename TX, typename TY>
static void launch_layernorm(const void* x, const float* g, const void* residual, void* y, int64_t M, int C, int operand_out,
                             cudaStream_t st) {
    const TX* xp = (const TX*)x;
    const TY* rp = (const TY*)residual;
    TY* yp = (TY*)y;
    if (C <= 128)
        channel_layernorm_kernel<4, 1, TX, TY><<<(unsigned)((M + 31) / 32), 256, 0, st>>>(xp, g, rp, yp, M, C, operand_out);
    else if (C <= 256)
        channel_layernorm_kernel<4, 2, TX, TY><<<(unsigned)((M + 31) / 32), 256, 0, st>>>(xp, g, rp, yp, M, C, operand_out);
    else if (C <= 512)
        channel_layernorm_kernel<2, 4, TX, TY><<<(unsigned)((M + 15) / 16), 256, 0, st>>>(xp, g, rp, yp, M, C, operand_out);
    else
        channel_layernorm_kernel<1, 8, TX, TY><<<(unsigned)((M + 7) / 8), 256, 0, st>>>(xp, g, rp, yp, M, C, operand_out);
}

extern "C" int sdc_channel_layernorm(int prec, const void* x, int x_operand, const float* g, const void* residual, void* y,
                                     int64_t M, int C, int operand_out, void* stream) {
    SDC_CHECK_PREC("channel_layernorm");
    SDC_REQUIRE(x && g && y && M > 0, "channel_layernorm: bad arguments");
    SDC_REQUIRE(C % 4 == 0 && C <= 1024, "channel_layernorm: C=%d unsupported (multiple of 4, <= 1024)", C);
    cudaStream_t st = as_stream(stream);
    if (prec == SDC_PREC_F16) {
        const __half* xh = (const __half*)x;
        const __half* rh = (const __half*)residual;
        __half* yh = (__half*)y;
        // rows per CTA = 8 warps x (32 / LPR) x RG
        if (x_operand && C == 128) channel_layernorm_h_kernel<8, 2, 2><<<(unsigned)((M + 63) / 64), 256, 0, st>>>(xh, g, rh, yh, M);
        else if (x_operand && C == 256) channel_layernorm_h_kernel<16, 2, 2><<<(unsigned)((M + 31) / 32), 256, 0, st>>>(xh, g, rh, yh, M);
        else if (x_operand && C == 512) channel_layernorm_h_kernel<32, 2, 2><<<(unsigned)((M + 15) / 16), 256, 0, st>>>(xh, g, rh, yh, M);
        else if (x_operand && C == 1024) channel_layernorm_h_kernel<32, 4, 1><<<(unsigned)((M + 7) / 8), 256, 0, st>>>(xh, g, rh, yh, M);
        else if (x_operand) launch_layernorm<__half, __half>(x, g, residual, y, M, C, 1, st);
        else launch_layernorm<float, __half>(x, g, residual, y, M, C, 1, st);
    } else {
        launch_layernorm<float, float>(x, g, residual, y, M, C, operand_out, st);
    }
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int64_t sdc_linear_attention_workspace(int B) { return (int64_t)B * LA_HEADS * LA_CTX * sizeof(float); }

extern "C" int sdc_linear_attention(int prec, const float* qkv, void* out, void* workspace, int B, int n, void* stream) {
    SDC_CHECK_PREC("linear_attention");
    SDC_REQUIRE(qkv && out && workspace && B > 0 && n > 0, "linear_attention: bad arguments");
    float* ctx = reinterpret_cast<float*>(workspace);
    linattn_context_kernel<float><<<(unsigned)(B * LA_HEADS), 256, 0, as_stream(stream)>>>(qkv + LA_HID, qkv + 2 * LA_HID, LA_QKV, ctx, n);
    SDC_LAUNCHED();
    SDC_REQUIRE(n % 32 == 0, "linear_attention: n=%d must be a multiple of 32", n);
    const size_t sm128 = (LA_HEADS * LA_D * LA_CLD + 128 * LA_QLD) * sizeof(float), sm32 = (LA_HEADS * LA_D * LA_CLD + 32 * LA_QLD) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(linattn_apply_kernel<128, __half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm128));
        SDC_CUDA(cudaFuncSetAttribute(linattn_apply_kernel<128, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm128));
        attr_set = true;
    }
    cudaStream_t st = as_stream(stream);
    if (n % 128 == 0) {
        const unsigned grid = (unsigned)(B * (n / 128));
        if (prec == SDC_PREC_F16) linattn_apply_kernel<128, __half><<<grid, 256, sm128, st>>>(qkv, ctx, (__half*)out, n);
        else linattn_apply_kernel<128, float><<<grid, 256, sm128, st>>>(qkv, ctx, (float*)out, n);
    } else {
        const unsigned grid = (unsigned)(B * (n / 32));
        if (prec == SDC_PREC_F16) linattn_apply_kernel<32, __half><<<grid, 256, sm32, st>>>(qkv, ctx, (__half*)out, n);
        else linattn_apply_kernel<32, float><<<grid, 256, sm32, st>>>(qkv, ctx, (float*)out, n);
    }
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_linear_attention_context(const void* k, const void* v, int ld, int kv_operand, void* workspace, int B, int n, void* stream) {
    SDC_REQUIRE(k && v && workspace && B > 0 && n > 0 && ld % 4 == 0, "linear_attention_context: bad arguments");
    if (kv_operand) {
        SDC_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 && (reinterpret_cast<uintptr_t>(v) & 15) == 0,
                    "linear_attention_context: fp16 k | v rows must be 16-byte aligned");
        linattn_context_mma_kernel<<<(unsigned)(B * LA_HEADS), 128, 0, as_stream(stream)>>>((const __half*)k, (const __half*)v, ld,
                                                                                          reinterpret_cast<float*>(workspace), n);
    }
    else
        linattn_context_kernel<float><<<(unsigned)(B * LA_HEADS), 256, 0, as_stream(stream)>>>((const float*)k, (const float*)v, ld,
                                                                                             reinterpret_cast<float*>(workspace), n);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_linear_attention_fold(int prec, const void* workspace, const float* w_out, void* w_folded, int B, int Cout,
                                         void* stream) {
    SDC_CHECK_PREC("linear_attention_fold");
    SDC_REQUIRE(workspace && w_out && w_folded && B > 0 && Cout > 0 && Cout % 16 == 0, "linear_attention_fold: Cout %% 16 != 0 or null pointer");
    const unsigned grid = (unsigned)B;
    static const bool use_mma = []() { const char* e = getenv("SDC_FOLD_MMA"); return !(e && e[0] == '0'); }();
    if (prec == SDC_PREC_F16 && use_mma)
        linattn_fold_mma_kernel<<<dim3(grid, (unsigned)((Cout + 127) / 128)), 256, 0, as_stream(stream)>>>(
            reinterpret_cast<const float*>(workspace), w_out, (__half*)w_folded, Cout);
    else if (prec == SDC_PREC_F16)
        linattn_fold_kernel<__half><<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(workspace), w_out, (__half*)w_folded, Cout);
    else
        linattn_fold_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float*>(workspace), w_out, (float*)w_folded, Cout);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_attention(int prec, const float* qkv, void* out, int B, int n, void* stream) {
    SDC_CHECK_PREC("attention");
    SDC_REQUIRE(qkv && out && B > 0, "attention: bad arguments");
    SDC_REQUIRE(n > 0 && n <= 32, "attention: n=%d tokens unsupported (bottleneck of the 16x128 grid has 32)", n);
    if (prec == SDC_PREC_F16) attention_kernel<__half><<<(unsigned)B, 128, 0, as_stream(stream)>>>(qkv, (__half*)out, n);
    else attention_kernel<float><<<(unsigned)B, 128, 0, as_stream(stream)>>>(qkv, (float*)out, n);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_upsample2x(int prec, const void* x, void* y, int B, int H, int W, int C, void* stream) {
    SDC_CHECK_PREC("upsample2x");
    const int eb = prec == SDC_PREC_F16 ? 2 : 4;
    SDC_REQUIRE(x && y && B > 0 && (C * eb) % 16 == 0, "upsample2x: bad arguments");
    const int v16 = C * eb / 16;   // 16-byte vectors per pixel
    const int64_t total4 = (int64_t)B * 4 * H * W * v16;
    upsample2x_kernel<<<blocks_for(total4, 256 * 4), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(x),
                                                                                 reinterpret_cast<float4*>(y), total4, H, W, v16);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_head_conv1(int prec, const void* x, const float* w, const float* bias, float* out, int B, int HW, int Cin,
                              int Cout, void* stream) {
    SDC_CHECK_PREC("head_conv1");
    SDC_REQUIRE(x && w && out && B > 0 && Cin % 4 == 0 && Cout >= 1 && Cout <= 4, "head_conv1: needs Cin %% 4 == 0, Cout <= 4");
    const int64_t M = (int64_t)B * HW;
    SDC_REQUIRE(M % 4 == 0, "head_conv1: B*HW must be a multiple of 4");   // a warp's 4 pixels are all valid or all absent
    const unsigned grid = blocks_for(M * 8, 256);
    const size_t sm = (size_t)Cout * Cin * sizeof(float);
    if (prec == SDC_PREC_F16 && Cin == 128 && Cout == 3)
        head_conv1_h128_kernel<<<blocks_for(M, 64), 128, 0, as_stream(stream)>>>((const __half*)x, w, bias, out, M, HW);
    else if (prec == SDC_PREC_F16)
        head_conv1_kernel<__half><<<grid, 256, sm, as_stream(stream)>>>((const __half*)x, w, bias, out, M, HW, Cin, Cout);
    else
        head_conv1_kernel<float><<<grid, 256, sm, as_stream(stream)>>>((const float*)x, w, bias, out, M, HW, Cin, Cout);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_linear_rows(const float* x, const float* w, const float* b, float* y, int R, int K, int N, int act_in,
                               void* stream) {
    SDC_REQUIRE(x && w && y && R > 0 && K > 0 && N > 0 && R < 65536, "linear_rows: bad arguments");
    dim3 grid((unsigned)((N + 7) / 8), (unsigned)R);
    linear_rows_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, w, b, y, K, N, act_in);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_sinusoidal_embedding(const float* t, float* emb, int R, int dim, float theta, void* stream) {
    SDC_REQUIRE(t && emb && R > 0 && dim % 2 == 0 && dim >= 4, "sinusoidal_embedding: bad arguments");
    const float neg_step = (float)(-(log((double)theta) / (double)(dim / 2 - 1)));
    sinusoidal_kernel<<<blocks_for((int64_t)R * dim / 2, 256), 256, 0, as_stream(stream)>>>(t, emb, R, dim, neg_step);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_zero_f64(double* p, int64_t n, void* stream) {
    SDC_REQUIRE(p != nullptr && n >= 0, "zero_f64: bad arguments");
    SDC_CUDA(cudaMemsetAsync(p, 0, (size_t)n * sizeof(double), as_stream(stream)));
    return SDC_OK;
}
