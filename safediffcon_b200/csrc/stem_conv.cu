// Stem: the 7x7 pad-3 convolution of the NCHW model input (/root/reference/1D/model/unet.py:326,392) as ONE tcgen05 kernel.
//
// Round 1 ran it as im2col (1.3 GB of fp16 patches written to HBM: 51x the 25 MB input) + a 1x1 GEMM reading them back: 0.6 ms per
// step at B = 1024.  The first in-kernel version built the full [128 px x 320] patch tile of an image row in shared memory (every
// input value written 49 x 2 times): 0.33 ms, bound by the builder warps.  This version uses the structure of the patch matrix: for a
// fixed horizontal tap kx the patch columns (ci, ky) of pixel w are the patch columns of pixel w + 1 for tap kx - 1, so ONE
// operand tile per image row serves all seven horizontal taps as ROW-SHIFTED views (the halo trick of conv_row.cu):
//   T[w' = 0..133][k' = ci * 7 + ky] = x[ci, h + ky - 3, w' - 3]        (zero outside the image = the padding)
//   out[w, co] = sum_kx sum_k' T[w + kx][k'] * W[co, ci, ky, kx]         -> 7 accumulating MMAs groups with A = T rows kx .. kx + 127
// A row of T is 128 bytes: fp16 HIGH parts of the 21 values at halfs [0, 21), LOW parts (x - high) at [24, 45), zeros elsewhere (the
// weights repeat in both ranges, so the product sees x to ~2^-22 although the operands are fp16); K = 48 = three 16-wide MMA steps.
//   builder warps (2 groups x 5, alternating image rows) read the 21 inputs of a T row straight from global memory (L2: the 25 MB
//                      input is read 7 times), split them and write the row in the K-major SWIZZLE_128B layout tcgen05 expects;
//                      one fence.proxy.async + one mbarrier arrive per warp and tile; 3-stage ring of 17 KB tiles;
//   warp 1             issues 7 taps x 3 tcgen05.mma.kind::f16 (M = 128 pixels, N = 128) against the seven [128 x 64] weight tiles that
//                      stay RESIDENT in shared memory (gathered once per CTA from the packed [Cout, 320] stem matrix);
//                      accumulators double-buffered in TMEM;
//   warps 2-9          epilogue (tc_ptx.cuh: epilogue_chunk, two warps per TMEM lane quarter): + bias, fp16, TMA store of the NHWC rows.
// Per image row: 134 x 8 shared-memory stores instead of 128 x 40, 21 MMAs instead of 20.  HBM traffic: 25 MB in, 537 MB out.
#include "tc_ptx.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <stdlib.h>

namespace sdc {

constexpr int ST_W = 128;                 // image width = UMMA M
constexpr int ST_COUT = 128;
constexpr int ST_TAPS = 7;
constexpr int ST_TROWS = 136;             // rows of the operand tile T (134 used), 17 groups of 8
constexpr int ST_TBYTES = ST_TROWS * 128; // 17 KB, keeps every stage 1024-byte aligned
constexpr int ST_STAGES = 3;
constexpr int ST_WBYTES = ST_COUT * 128;  // one weight tile (tap kx): 128 output channels x 64 halfs
constexpr int ST_EPI = 8;
constexpr int ST_GROUP_WARPS = 5;         // builder warps per group (160 threads >= 134 rows)
constexpr int ST_THREADS = 32 * (2 + ST_EPI + 2 * ST_GROUP_WARPS);   // warp 0 idle after setup, 1 MMA, 2-9 epilogue, 10-19 builders
constexpr int ST_LO = 24;                 // first half of the LOW parts inside a T row

struct StemParams {
    int B, H, Cin, tiles_total, tiles_per_cta, kp;
    int dbg;   // timing experiments (SDC_STEM_DBG): 1 = no epilogue work, 2 = builders only signal, 4 = no MMAs
    const float* x;
    const __half* w;      // [Cout, kp] fp16, column ci * 49 + ky * 7 + kx
    const float* bias;
};

__global__ void __launch_bounds__(ST_THREADS, 1)
stem_conv7_tc_kernel(const __grid_constant__ CUtensorMap map_out, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* w_tile = smem;                                   // [7][16 KB]
    uint8_t* a_tile = smem + ST_TAPS * ST_WBYTES;             // [3][17 KB]
    uint8_t* staging = a_tile + ST_STAGES * ST_TBYTES;        // 8 x 4 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + ST_EPI * 4096);
    uint64_t* a_full = bars;                  // [3]
    uint64_t* a_empty = a_full + ST_STAGES;   // [3]
    uint64_t* acc_full = a_empty + ST_STAGES; // [2]
    uint64_t* acc_empty = acc_full + 2;       // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_lo = blockIdx.x * p.tiles_per_cta;
    const int tile_hi = min(p.tiles_total, tile_lo + p.tiles_per_cta);
    const int K7 = p.Cin * 7;   // live columns of a T row (<= 21)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < ST_STAGES; ++s) { mbar_init(&a_full[s], ST_GROUP_WARPS); mbar_init(&a_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], ST_EPI); }
        fence_barrier_init();
    }
    // resident weights: tile kx, row co, 16-byte chunk ch (8 halfs) in the K-major SWIZZLE_128B layout
    for (int i = threadIdx.x; i < ST_TAPS * ST_COUT * 8; i += ST_THREADS) {
        const int ch = i & 7, co = (i >> 3) & (ST_COUT - 1), kx = i >> 10;
        uint32_t wv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __half2 h2;
            __half hv[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = ch * 8 + 2 * e + u;                         // half index inside the row
                const int k = col < ST_LO ? col : col - ST_LO;              // k' = ci * 7 + ky
                const bool live = k < K7 && col < ST_LO + 21 && (col < 21 || col >= ST_LO);
                hv[u] = live ? p.w[(size_t)co * p.kp + (k / 7) * 49 + (k % 7) * 7 + kx] : __float2half(0.f);
            }
            h2 = __halves2half2(hv[0], hv[1]);
            wv[e] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        const uint32_t dst = smem_u32(w_tile + kx * ST_WBYTES) + (uint32_t)((co >> 3) * 1024 + (co & 7) * 128 + ((ch ^ (co & 7)) << 4));
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
    }
    fence_proxy_async();
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 1) {
        const uint32_t idesc = Operand<true>::idesc(ST_COUT, ST_W);
        int it = 0, s = 0;
        uint32_t ph = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            mbar_wait(&a_full[s], ph);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * 128u;
            const uint32_t ta = smem_u32(a_tile + s * ST_TBYTES), wa = smem_u32(w_tile);
            if (elect_one()) {
#pragma unroll
                for (int kx = 0; kx < ((p.dbg & 4) ? 0 : ST_TAPS); ++kx) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)   // A = rows kx .. kx + 127 of T (row-shifted view), 32 bytes of K per step
                        umma_f16(tmem_d, make_sw128_desc(ta + 128u * kx + 32u * k), make_sw128_desc(wa + (uint32_t)(kx * ST_WBYTES) + 32u * k),
                                 idesc, (kx | k) != 0);
                }
                umma_commit(&a_empty[s]);
                umma_commit(&acc_full[buf]);
            }
            __syncwarp();
            if (++s == ST_STAGES) { s = 0; ph ^= 1u; }
        }
    } else if (warp >= 2 && warp < 2 + ST_EPI) {
        const int q = warp & 3, half_id = (warp - 2) >> 2;
        const uint32_t stg = smem_u32(staging + (warp - 2) * 4096);
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int m_w = tile * ST_W + q * 32;
            float s1 = 0.f, s2 = 0.f;
            for (int c = 32 * half_id; c < ((p.dbg & 1) ? 0 : ST_COUT); c += 64) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 128u + (uint32_t)c;
                epilogue_chunk<true, __half>(taddr, stg, &map_out, c, m_w, true, p.bias, nullptr, false, s1, s2, lane, -1, 0, true, false);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (lane == 0) bulk_wait<0>();
        tc_fence_before();
    } else if (warp >= 2 + ST_EPI) {
        // ---------------- builders: two groups of 160 threads, group g builds the CTA's tiles g, g + 2, ... ----------------
        const int bw = warp - (2 + ST_EPI), g = bw / ST_GROUP_WARPS;
        const int r = (bw - g * ST_GROUP_WARPS) * 32 + lane;     // row of T (pixel w' = r - 3)
        const int HW = p.H * ST_W;
        const int ww = r - 3;
        const bool col_ok = ww >= 0 && ww < ST_W;
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        const uint32_t sw = (uint32_t)(r & 7);
        // the 21 inputs of this row for tile `it` (independent loads, in flight together; zero = padding)
        float v[21];
        auto load_row = [&](int it) {
            const int tile = tile_lo + it;
            const int b = tile / p.H, h = tile - b * p.H;
            const float* xb = p.x + (size_t)b * p.Cin * HW + ww;
#pragma unroll
            for (int k = 0; k < 21; ++k) {
                const int ci = k / 7, ky = k - 7 * ci;
                const int hh = h + ky - 3;
                v[k] = (col_ok && k < K7 && hh >= 0 && hh < p.H) ? __ldg(xb + (size_t)ci * HW + hh * ST_W) : 0.f;
            }
        };
        if (tile_lo + g < tile_hi && !(p.dbg & 2)) load_row(g);
        for (int it = g; tile_lo + it < tile_hi; it += 2) {
            uint32_t hi[12], lo[12];   // halfs [0, 24) of the high and of the low range (k = 21 .. 23 zero)
#pragma unroll
            for (int e = 0; e < 12; ++e) {
                const float v0 = 2 * e < 21 ? v[2 * e] : 0.f, v1 = 2 * e + 1 < 21 ? v[2 * e + 1] : 0.f;
                const __half2 h2 = __floats2half2_rn(v0, v1);
                const float2 f2 = __half22float2(h2);
                const __half2 l2 = __floats2half2_rn(v0 - f2.x, v1 - f2.y);
                hi[e] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[e] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            if (tile_lo + it + 2 < tile_hi && !(p.dbg & 2)) load_row(it + 2);   // the group's next tile: its L2 round trip overlaps the stores / fence below
            const int s = it % ST_STAGES;
            mbar_wait(&a_empty[s], ((uint32_t)(it / ST_STAGES) & 1u) ^ 1u);
            if (r < ST_W + 6 && !(p.dbg & 2)) {
                const uint32_t base = smem_u32(a_tile + s * ST_TBYTES) + row_off;
                // chunks of 8 halfs: 0-2 = high [0, 24), 3-5 = low [24, 48); chunks 6, 7 (halfs 48 .. 63) are never read (K = 48)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (((uint32_t)ch ^ sw) << 4)), "r"(hi[4 * ch]), "r"(hi[4 * ch + 1]),
                                 "r"(hi[4 * ch + 2]), "r"(hi[4 * ch + 3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + (((uint32_t)(ch + 3) ^ sw) << 4)), "r"(lo[4 * ch]), "r"(lo[4 * ch + 1]),
                                 "r"(lo[4 * ch + 2]), "r"(lo[4 * ch + 3]) : "memory");
                }
            }
            fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[s]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
}

}  // namespace sdc

using namespace sdc;

// Returns SDC_OK when handled, -1 when the shape is not eligible (then: sdc_stem_im2col + sdc_conv_gemm).
// w_packed: the [Cout, kp = 320] fp16 matrix sdc_pack_conv_weight(kind 0) makes of the stem weight repeated at columns 0 and 160.
extern "C" int sdc_stem_conv7_tc(const float* x, const void* w_packed, const float* bias, void* out, int B, int Cin, int H, int W, int Cout,
                                 int kp, void* stream) {
    SDC_REQUIRE(x && w_packed && out && B > 0 && H > 0, "stem_conv7_tc: bad arguments");
    if (W != ST_W || Cout != ST_COUT || kp != 320 || Cin < 1 || Cin > 3 || Cin * 49 > 160) return -1;
    CUtensorMap mo;
    {
        const int rc = encode_out_tmap(&mo, out, (int64_t)B * H * W, Cout, true);
        if (rc) return rc;
    }
    StemParams p{};
    p.B = B; p.H = H; p.Cin = Cin; p.x = x; p.bias = bias; p.kp = kp;
    p.w = reinterpret_cast<const __half*>(w_packed);
    p.tiles_total = B * H;
    { const char* e = getenv("SDC_STEM_DBG"); p.dbg = e ? atoi(e) : 0; }
    int n_sm = 148, dev = 0;
    SDC_CUDA(cudaGetDevice(&dev));
    SDC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int ctas = p.tiles_total < n_sm ? p.tiles_total : n_sm;
    p.tiles_per_cta = (p.tiles_total + ctas - 1) / ctas;
    const int grid = (p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const int smem_bytes = ST_TAPS * ST_WBYTES + ST_STAGES * ST_TBYTES + ST_EPI * 4096 + 10 * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(stem_conv7_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    stem_conv7_tc_kernel<<<grid, ST_THREADS, smem_bytes, as_stream(stream)>>>(mo, p);
    SDC_LAUNCHED();
    return SDC_OK;
}
