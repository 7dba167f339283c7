// Stem: the 7x7 pad-3 convolution of the NCHW model input (/root/reference/1D/model/unet.py:326,392) as ONE tcgen05 kernel.
//
// Round 1 ran it as im2col (1.3 GB of fp16 patches written to HBM: 51x the 25 MB input) + a 1x1 GEMM reading them back: 0.6 ms per
// step at B = 1024.  Here the patch matrix never exists in memory: per image row (128 pixels = one UMMA M tile)
//   builder warps (16) stage the 7 input rows the tile needs in shared memory (cp.async, zero fill = padding; the next tile's rows
//                      are in flight while this one is built) and WRITE THE A OPERAND THEMSELVES: the [128 px x 320] patch tile in
//                      the K-major SWIZZLE_128B layout tcgen05 expects (what TMA would have produced), generic-proxy stores +
//                      ONE fence.proxy.async per tile (a fence per K block made the builders latency bound).  Columns [0, 147) hold the fp16 HIGH part of
//                      x[ci, h+ky-3, w+kx-3] (k = ci*49 + ky*7 + kx), [160, 307) the LOW part (x - high): with the weights
//                      repeated in both ranges the product sees x to ~2^-22 although the operands are fp16;
//   warp 1             issues 5 K blocks x 4 tcgen05.mma.kind::f16 (M = 128, N = 128) against the weight matrix that stays RESIDENT in
//                      shared memory (80 KB, loaded once per CTA by TMA); accumulators double-buffered in TMEM;
//   warps 2-5          epilogue (tc_ptx.cuh: epilogue_chunk): + bias, fp16, TMA store of the NHWC rows.
// The five K blocks of the A tile form a ring (full / empty barrier per block): the next tile's first block is rewritten as soon
// as this tile's MMAs have released it, so building tile i+1 overlaps the MMAs and the epilogue of tile i.  HBM traffic: 25 MB in, 537 MB out.
#include "tc_ptx.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <stdlib.h>

namespace sdc {

constexpr int ST_W = 128;                 // image width = UMMA M
constexpr int ST_COUT = 128;
constexpr int ST_KB = 5;                  // K blocks of 64 fp16 (kp = 320)
constexpr int ST_KBLK = ST_W * 128;       // 16 KB per K block (A or W)
constexpr int ST_WP = ST_W + 6;
constexpr int ST_BUILD_WARPS = 16;
constexpr int ST_THREADS = 32 * (6 + ST_BUILD_WARPS);   // warp 0 weights TMA, 1 MMA, 2-5 epilogue, 6-21 builders
constexpr int ST_STG = 4 * 4096;

struct StemParams {
    int B, H, Cin, tiles_total, tiles_per_cta, dbg;
    const float* x;
    const float* bias;
};

__device__ __forceinline__ float lds32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

__global__ void __launch_bounds__(ST_THREADS, 1)
stem_conv7_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* a_tile = smem;                                  // [5][16 KB]
    uint8_t* w_tile = smem + ST_KB * ST_KBLK;                // [5][16 KB]
    uint8_t* staging = w_tile + ST_KB * ST_KBLK;             // 4 x 4 KB
    float* xs = reinterpret_cast<float*>(staging + ST_STG);  // [2][Cin * 7 * WP] (Cin <= 4)
    const int xs_elems = p.Cin * 7 * ST_WP;
    int* koff = reinterpret_cast<int*>(xs + 2 * 4 * 7 * ST_WP);          // [320] offset of patch column k inside the window, -1 = padding
    uint64_t* bars = reinterpret_cast<uint64_t*>(koff + 320);
    uint64_t* w_full = bars;             // [1]
    uint64_t* a_full = bars + 1;         // [5]
    uint64_t* a_empty = a_full + ST_KB;  // [5]
    uint64_t* acc_full = a_empty + ST_KB;   // [2]
    uint64_t* acc_empty = acc_full + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_lo = blockIdx.x * p.tiles_per_cta;
    const int tile_hi = min(p.tiles_total, tile_lo + p.tiles_per_cta);
    const int K = p.Cin * 49;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_out);
        mbar_init(w_full, 1);
        for (int s = 0; s < ST_KB; ++s) { mbar_init(&a_full[s], ST_BUILD_WARPS); mbar_init(&a_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 4); }
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k < 320; k += ST_THREADS) {
        const int kk = k < 160 ? k : k - 160;
        const int ci = kk / 49, t = kk - ci * 49, ky = t / 7, kx = t - ky * 7;
        koff[k] = kk < K ? (ci * 7 + ky) * ST_WP + kx : -1;
    }
    if (warp == 1) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {   // the whole weight matrix, once
            mbar_expect_tx(w_full, (uint32_t)(ST_KB * ST_KBLK));
            for (int kb = 0; kb < ST_KB; ++kb) tma_load_2d(w_tile + kb * ST_KBLK, &map_w, w_full, kb * 64, 0);
        }
    } else if (warp == 1) {
        const uint32_t idesc = Operand<true>::idesc(ST_COUT, ST_W);
        mbar_wait(w_full, 0u);
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)buf * 128u;
            for (int kb = 0; kb < ST_KB; ++kb) {
                mbar_wait(&a_full[kb], (uint32_t)it & 1u);
                tc_fence_after();
                const uint64_t adesc = make_sw128_desc(smem_u32(a_tile + kb * ST_KBLK));
                const uint64_t bdesc = make_sw128_desc(smem_u32(w_tile + kb * ST_KBLK));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                    umma_commit(&a_empty[kb]);
                    if (kb == ST_KB - 1) umma_commit(&acc_full[buf]);
                }
                __syncwarp();
            }
        }
    } else if (warp < 6) {
        const int q = warp & 3;
        const uint32_t stg = smem_u32(staging + q * 4096);
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int m_w = tile * ST_W + q * 32;
            float s1 = 0.f, s2 = 0.f;
            for (int c = 0; c < ((p.dbg & 2) ? 0 : ST_COUT); c += 32) {
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 128u + (uint32_t)c;
                if (p.dbg & 1) {   // timing experiment: TMEM load + conversion only, no staging / fence / store
                    uint32_t r[32];
                    tmem_ld32(taddr, r);
                    float acc = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc += __uint_as_float(r[j]);
                    if (acc == 123.456f) s1 += acc;
                } else
                epilogue_chunk<true, __half>(taddr, stg, &map_out, c, m_w, true, p.bias, nullptr, false, s1, s2, lane, -1, 0, true, false);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);
        }
        if (lane == 0) bulk_wait<0>();
        tc_fence_before();
    } else {
        // ---------------- builders: 256 threads ----------------
        const int bt = threadIdx.x - 6 * 32;
        const int HW = p.H * ST_W;
        auto prefetch = [&](int tile, int slot) {   // the 7 x (W + 6) x Cin window of image row `tile` -> xs[slot], zero padded
            const int b = tile / p.H, h = tile - b * p.H;
            float* dst = xs + slot * (4 * 7 * ST_WP);
            for (int i = bt; i < xs_elems; i += 32 * ST_BUILD_WARPS) {
                const int ci = i / (7 * ST_WP), r = (i / ST_WP) % 7, c = i % ST_WP;
                const int hh = h + r - 3, ww = c - 3;
                const bool ok = hh >= 0 && hh < p.H && ww >= 0 && ww < ST_W;
                const float* src = ok ? p.x + ((size_t)b * p.Cin + ci) * HW + (size_t)hh * ST_W + ww : p.x;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst + i)), "l"(src), "r"(ok ? 4 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        if (tile_lo < tile_hi) prefetch(tile_lo, 0);
        // Work item = (16-byte chunk c of the 20 high-part chunks, pixel row r): the thread loads the 8 patch values ONCE and writes both
        // the high chunk c and the low chunk c + 20 (the kernel is shared-memory-bandwidth bound: window reads + operand writes +
        // the tensor core's own operand reads; one thread per (chunk, row) of the 40 chunks read every window value twice).
        // Lanes = 32 consecutive rows of one chunk: window reads are consecutive floats, the swizzled 16-byte stores conflict free.
        const int r = bt & 127, c0 = bt >> 7;   // chunks c0, c0 + 4, .., c0 + 16
        int off[5][8];                         // window offsets of the 8 patch columns of every chunk (-1 = padding column)
#pragma unroll
        for (int i = 0; i < 5; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e) off[i][e] = koff[(c0 + 4 * i) * 8 + e];
        const uint32_t row_off = (uint32_t)((r >> 3) * 1024 + (r & 7) * 128);
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 2, 512;" ::: "memory");          // every builder's copies of this tile have landed
            if (tile + 1 < tile_hi) prefetch(tile + 1, (it + 1) & 1);
            // explicit shared-window loads: `xs` descends from the re-aligned dynamic-smem base, so plain dereferences compile to
            // generic LD.E (64-bit addressing, long scoreboard) -- ld.shared through a 32-bit address instead
            const uint32_t win = smem_u32(xs + (it & 1) * (4 * 7 * ST_WP) + r);
            uint32_t waited = 0;   // K blocks whose release by the previous tile's MMAs this thread has already observed
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const int c = c0 + 4 * i;
                const int kb_h = c >> 3, kb_l = (c + 20) >> 3;
                if (!(waited >> kb_h & 1u)) { mbar_wait(&a_empty[kb_h], ((uint32_t)it & 1u) ^ 1u); waited |= 1u << kb_h; }
                if (!(waited >> kb_l & 1u)) { mbar_wait(&a_empty[kb_l], ((uint32_t)it & 1u) ^ 1u); waited |= 1u << kb_l; }
                uint32_t hv[4], lv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float v0 = off[i][2 * e] >= 0 ? lds32(win + 4u * (uint32_t)off[i][2 * e]) : 0.f;
                    const float v1 = off[i][2 * e + 1] >= 0 ? lds32(win + 4u * (uint32_t)off[i][2 * e + 1]) : 0.f;
                    const __half2 h2 = __floats2half2_rn(v0, v1);
                    const float2 f2 = __half22float2(h2);
                    const __half2 l2 = __floats2half2_rn(v0 - f2.x, v1 - f2.y);
                    hv[e] = *reinterpret_cast<const uint32_t*>(&h2);
                    lv[e] = *reinterpret_cast<const uint32_t*>(&l2);
                }
                // K-major SWIZZLE_128B: row r at (r / 8) * 1024 + (r % 8) * 128, 16-byte chunk ch at ((ch ^ (r % 8)) * 16)
                const uint32_t dh = smem_u32(a_tile + kb_h * ST_KBLK) + row_off + (uint32_t)(((c & 7) ^ (r & 7)) << 4);
                const uint32_t dl = smem_u32(a_tile + kb_l * ST_KBLK) + row_off + (uint32_t)((((c + 20) & 7) ^ (r & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dh), "r"(hv[0]), "r"(hv[1]), "r"(hv[2]), "r"(hv[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dl), "r"(lv[0]), "r"(lv[1]), "r"(lv[2]), "r"(lv[3]) : "memory");
            }
            fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int kb = 0; kb < ST_KB; ++kb) mbar_arrive(&a_full[kb]);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
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
    CUtensorMap mw, mo;
    {
        cuuint64_t wd[2] = {(cuuint64_t)kp, (cuuint64_t)Cout};
        cuuint64_t ws[1] = {(cuuint64_t)kp * 2};
        cuuint32_t wb[2] = {64, (cuuint32_t)Cout};
        int rc = encode_tmap(&mw, w_packed, 2, wd, ws, wb, true);
        if (rc) return rc;
        rc = encode_out_tmap(&mo, out, (int64_t)B * H * W, Cout, true);
        if (rc) return rc;
    }
    StemParams p{};
    p.B = B; p.H = H; p.Cin = Cin; p.x = x; p.bias = bias;
    p.tiles_total = B * H;
    { const char* e = getenv("SDC_STEM_DBG"); p.dbg = e ? atoi(e) : 0; }
    int n_sm = 148, dev = 0;
    SDC_CUDA(cudaGetDevice(&dev));
    SDC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int ctas = p.tiles_total < n_sm ? p.tiles_total : n_sm;
    p.tiles_per_cta = (p.tiles_total + ctas - 1) / ctas;
    const int grid = (p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta;
    const int smem_bytes = 2 * ST_KB * ST_KBLK + ST_STG + 2 * 4 * 7 * ST_WP * 4 + 320 * 4 + 16 * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(stem_conv7_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    stem_conv7_tc_kernel<<<grid, ST_THREADS, smem_bytes, as_stream(stream)>>>(mw, mo, p);
    SDC_LAUNCHED();
    return SDC_OK;
}
