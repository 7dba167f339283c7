// Convolution weight gradient on tcgen05 (what autograd accumulates into Conv2d.weight when the
// reference back-propagates the fine-tuning / post-training loss, /root/reference/1D/inference/inference_ft.py:189-226):
//
//     dW[co, ci, tap] += sum_p dY[p, co] * A[p + tap, ci]            GEMM per tap: M = co, N = ci, K = PIXELS
//
// Both operands live in memory as pixel rows ([p, channels], channels contiguous), i.e. they are MN-major for this GEMM.  tcgen05
// reads MN-major operands directly (instruction-descriptor bits 15/16); for 32-bit (TF32) MN-major operands the ONLY shared-memory
// layout is "128-byte swizzle with 32-byte atoms": rows of 128 bytes (32 fp32 along M/N), K groups of 4 rows, the four 32-byte
// chunks of a row permuted by (row mod 4) -- descriptor layout type 1, LBO = stride between 32-element MN blocks, SBO = stride
// between 4-row K groups; TMA writes exactly that with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.  So NO transpose pass exists: TMA
// drops [64 pixels x 32 channels] boxes of dY and of the tap-shifted activation window (out-of-image pixels are zero-filled by
// TMA = the convolution padding) straight into the operand layout.
//
//   work item = (128 output channels, 128 input channels, tap, slice of the pixel axis); one CTA per item
//   warp 0   TMA producer: per 64-pixel stage 4 boxes of dY + 4 boxes of A (64 KB), 3 stages
//   warp 1   TMEM allocation + MMA issue: 8 x tcgen05.mma.kind::tf32 (M = 128, N = 128, K = 8 pixels) per stage
//   warps 2-5 epilogue: TMEM -> registers -> red.global.add.f32 into the OIHW gradient (the pixel slices of a tile add up there)
// Precision: TF32 operands (dY arrives TF32-rounded from the backward-data pass; fp16 activations convert exactly), FP32
// accumulation -- the same arithmetic as the mma.sync kernel of unet_wgrad.cu, which remains the fallback for shapes this
// kernel does not take (pixel-unshuffle convolutions, channel counts that are not multiples of 128).
#include "tc_ptx.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <stdlib.h>

namespace sdc {

constexpr int WT_KT = 64;                 // pixels per stage
constexpr int WT_BOX = WT_KT * 128;       // one [64 px x 32 fp32] box: 8 KB
constexpr int WT_STAGE = 8 * WT_BOX;      // 4 boxes of dY (M = 128) + 4 boxes of A (N = 128)
constexpr int WT_STAGES = 3;
constexpr int WT_THREADS = 192;

struct WgradTcParams {
    int Cout, Cin, c0, taps, kind;
    int H, W, hw;
    int wbox, bh, bb;           // activation box: wbox pixels x bh rows x bb images = 64 pixels
    int tiles_k;                // 64-pixel tiles of the whole pixel axis
    int tiles_per_split;
    int n_ci, n_co;             // 128-channel tiles
    float* dw;
};

// MN-major 32-bit operand, SWIZZLE_128B with 32-byte atoms: 128-byte rows (32 fp32 along M/N), 4 K-rows per 512-byte group.
//   bits [0,14) start >> 4 | [16,30) LBO >> 4 (stride between 32-element MN blocks) | [32,46) SBO >> 4 (stride between 4-row K groups)
//   | [46,48) version 1 | [61,64) layout 1 (SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16);
    const uint32_t hi = (sbo_bytes >> 4) | (1u << 14) | (1u << 29);
    return ((uint64_t)hi << 32) | lo;
}

__global__ void __launch_bounds__(WT_THREADS, 1)
conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_a0,
                     const __grid_constant__ CUtensorMap map_a1, const WgradTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WT_STAGES * WT_STAGE);
    uint64_t* empty_bar = full_bar + WT_STAGES;
    uint64_t* acc_full = empty_bar + WT_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_full + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // item -> (co tile, ci tile, tap, pixel slice)
    int item = blockIdx.x;
    const int ci_t = item % p.n_ci; item /= p.n_ci;
    const int tap = item % p.taps; item /= p.taps;
    const int co_t = item % p.n_co;
    const int split = item / p.n_co;
    const int kt_lo = split * p.tiles_per_split;
    const int kt_hi = min(p.tiles_k, kt_lo + p.tiles_per_split);
    const int n_kt = kt_hi - kt_lo;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_dy);
        tma_prefetch_desc(&map_a0);
        for (int s = 0; s < WT_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int ci0 = ci_t * 128;
            const bool second = ci0 >= p.c0;
            const CUtensorMap* ma = second ? &map_a1 : &map_a0;
            const int cseg = second ? ci0 - p.c0 : ci0;
            int dy = 0, dx = 0;
            if (p.kind == 1) { dy = tap / 3 - 1; dx = tap % 3 - 1; }
            for (int i = 0; i < n_kt; ++i) {
                const int s = i % WT_STAGES;
                mbar_wait(&empty_bar[s], (((uint32_t)(i / WT_STAGES)) & 1u) ^ 1u);
                uint8_t* st = smem + s * WT_STAGE;
                mbar_expect_tx(&full_bar[s], (uint32_t)WT_STAGE);
                const int p0 = (kt_lo + i) * WT_KT;
                const int b0 = p0 / p.hw, rem = p0 - b0 * p.hw;
                const int h0 = rem / p.W, w0 = rem - h0 * p.W;
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_2d(st + j * WT_BOX, &map_dy, &full_bar[s], co_t * 128 + 32 * j, p0);
#pragma unroll
                for (int j = 0; j < 4; ++j) tma_load_4d(st + (4 + j) * WT_BOX, ma, &full_bar[s], cseg + 32 * j, w0 + dx, h0 + dy, b0);
            }
        }
    } else if (warp == 1) {
        // instruction descriptor: D fp32 (bit 4), A/B tf32 (2 << 7, 2 << 10), A and B MN-major (bits 15, 16), N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int i = 0; i < n_kt; ++i) {
            const int s = i % WT_STAGES;
            mbar_wait(&full_bar[s], ((uint32_t)(i / WT_STAGES)) & 1u);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + s * WT_STAGE), sb = sa + 4 * WT_BOX;
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < WT_KT / 8; ++k) {   // 8 pixels (two 4-row K groups, 1024 bytes) per MMA
                    umma_tf32(tmem_base, make_mn_sw128_desc(sa + 1024u * k, WT_BOX, 512u), make_mn_sw128_desc(sb + 1024u * k, WT_BOX, 512u),
                              idesc, (i | k) != 0);
                }
                umma_commit(&empty_bar[s]);
                if (i == n_kt - 1) umma_commit(acc_full);
            }
            __syncwarp();
        }
    } else if (n_kt > 0) {
        const int q = warp & 3;
        mbar_wait(acc_full, 0u);
        tc_fence_after();
        const int co = co_t * 128 + q * 32 + lane;
        float* drow = p.dw + ((size_t)co * p.Cin + (size_t)ci_t * 128) * p.taps + tap;
        for (int c = 0; c < 128; c += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(drow + (size_t)(c + j) * p.taps, __uint_as_float(r[j]));
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

__global__ void half_to_float_kernel(const __half* __restrict__ x, float* __restrict__ y, int64_t n8) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
        const uint4 u = reinterpret_cast<const uint4*>(x)[i];
        const __half2* h = reinterpret_cast<const __half2*>(&u);
        const float2 a = __half22float2(h[0]), b = __half22float2(h[1]), c = __half22float2(h[2]), d = __half22float2(h[3]);
        reinterpret_cast<float4*>(y)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
        reinterpret_cast<float4*>(y)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
    }
}

}  // namespace sdc

using namespace sdc;

extern "C" int64_t sdc_conv_wgrad_tc_scratch(int a_half, int c0, int c1, int B, int H, int W) {
    return a_half ? (int64_t)B * H * W * (c0 + c1) * 4 : 0;
}

// Returns SDC_OK when handled here, -1 when the shape is not eligible (the caller then uses sdc_conv_wgrad).
extern "C" int sdc_conv_wgrad_tc(int kind, int a_half, const void* a0, int c0, const void* a1, int c1, const float* dy, float* dw, int B,
                                 int H, int W, int Cout, void* scratch, int64_t scratch_bytes, void* stream) {
    SDC_REQUIRE(a0 && dy && dw && B > 0 && H > 0 && W > 0 && Cout > 0 && c0 > 0 && c1 >= 0 && (c1 == 0 || a1), "conv_wgrad_tc: bad arguments");
    if (kind != 0 && kind != 1) return -1;
    if (Cout % 128 != 0 || c0 % 128 != 0 || c1 % 128 != 0) return -1;
    if (!(W == 16 || W == 32 || W == 64 || W == 128)) return -1;
    const int hw = H * W;
    int wbox = W < WT_KT ? W : WT_KT, bh = WT_KT / wbox;
    if (bh > H) bh = H;
    const int bb = WT_KT / (wbox * bh);
    if (wbox * bh * bb != WT_KT || (bb > 1 && bh != H) || (W <= WT_KT && H % bh != 0)) return -1;
    const int64_t M = (int64_t)B * hw;
    cudaStream_t st = as_stream(stream);
    const float* f0 = (const float*)a0;
    const float* f1 = (const float*)a1;
    if (a_half) {
        const int64_t need = M * (c0 + c1) * 4;
        SDC_REQUIRE(scratch && scratch_bytes >= need, "conv_wgrad_tc: scratch of %lld bytes, need %lld (sdc_conv_wgrad_tc_scratch)",
                    (long long)scratch_bytes, (long long)need);
        float* s0 = (float*)scratch;
        float* s1 = s0 + M * c0;
        half_to_float_kernel<<<592, 256, 0, st>>>((const __half*)a0, s0, M * c0 / 8);
        SDC_LAUNCHED();
        if (c1) { half_to_float_kernel<<<592, 256, 0, st>>>((const __half*)a1, s1, M * c1 / 8); SDC_LAUNCHED(); }
        f0 = s0;
        f1 = c1 ? s1 : nullptr;
    }
    WgradTcParams p{};
    p.Cout = Cout; p.Cin = c0 + c1; p.c0 = c0; p.kind = kind; p.taps = kind == 1 ? 9 : 1; p.H = H; p.W = W; p.hw = hw;
    p.wbox = wbox; p.bh = bh; p.bb = bb; p.dw = dw;
    p.tiles_k = (int)((M + WT_KT - 1) / WT_KT);
    p.n_ci = p.Cin / 128; p.n_co = Cout / 128;
    const int base = p.n_ci * p.n_co * p.taps;
    int splits = (2 * 148 + base - 1) / base;
    const int max_splits = (p.tiles_k + 3) / 4;          // at least 4 stages of work per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.tiles_per_split = (p.tiles_k + splits - 1) / splits;
    splits = (p.tiles_k + p.tiles_per_split - 1) / p.tiles_per_split;

    CUtensorMap mdy, ma0, ma1;
    {
        cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)M};
        cuuint64_t str[1] = {(cuuint64_t)Cout * 4};
        cuuint32_t box[2] = {32, (cuuint32_t)WT_KT};
        int rc = encode_tmap(&mdy, dy, 2, dims, str, box, false, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
        if (rc) return rc;
    }
    auto enc_act = [&](CUtensorMap* m, const float* a, int C) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)hw * C * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)wbox, (cuuint32_t)bh, (cuuint32_t)bb};
        return encode_tmap(m, a, 4, dims, str, box, false, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
    };
    int rc = enc_act(&ma0, f0, c0);
    if (rc) return rc;
    if (c1) { rc = enc_act(&ma1, f1, c1); if (rc) return rc; } else ma1 = ma0;
    const int smem_bytes = WT_STAGES * WT_STAGE + (2 * WT_STAGES + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    conv_wgrad_tc_kernel<<<(unsigned)(base * splits), WT_THREADS, smem_bytes, st>>>(mdy, ma0, ma1, p);
    SDC_LAUNCHED();
    return SDC_OK;
}
