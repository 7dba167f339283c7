// 3x3 convolution for the 64-pixel-wide level (W = 64, Cout <= 128) with ACTIVATION REUSE ACROSS THE VERTICAL TAPS.
//
// Same math as conv_gemm.cu kind 1 (reference: nn.Conv2d(C, C', 3, padding=1) inside Block, /root/reference/1D/model/unet.py:132).
// With only 128 output channels the generic implicit GEMM is bound by its L2 -> shared-memory operand stream (16 KB of
// activations + 8 KB of weights per K block and CTA: 96 B/clk against the ~45-50 B/clk an SM ingests; tensor pipe 43 % active on
// the four 128 -> 128 convolutions of the 8x64 level, profiles/r02_per_launch_metrics_B1024.csv).  The halo kernel of the
// 16x128 level (conv_row.cu) cannot be used: a 128-row UMMA operand must be ONE contiguous run of shared-memory rows, and with
// W = 64 an M tile is two image rows, which are 66 pixels apart inside a halo.  Here the horizontal shift is done by TMA and
// the vertical one by the descriptor:
//   per channel chunk (128 bytes) and horizontal tap dx a CTA loads ONE box of 6 image rows x 64 pixels starting at column
//   dx - 1 (out-of-range rows / columns zero-filled = the padding): a dense [6][64] array of 128-byte pixel rows, 48 KB;
//   the A operand of M tile j (image rows 2j, 2j + 1 of the CTA's four) and vertical tap dy is the contiguous run of 128 pixel
//   rows starting at row (2j + dy) * 64 of that box -- expressed through the UMMA descriptor start address only;
//   each weight tile (chunk, dx, dy) is used by both M tiles.
// A cluster of two CTAs (cta_group::2) computes EIGHT image rows: CTA r rows h0 + 4r .. h0 + 4r + 3, every MMA is M = 256
// (M tile j of both CTAs) and each CTA stages half of every weight tile.  Operand traffic per CTA and 8-row item:
// 6 boxes x 48 KB + 18 half tiles x 8 KB = 432 KB against 864 KB for the same pixels in the generic kernel.
//
//   warp 0   TMA producer of the weight half-tile ring (6 stages);   warp 6: activation-box ring (3 stages)
//   warp 1   MMA issuer: per box 3 vertical taps x 2 M tiles x 4 (32 bytes of K) tcgen05.mma, accumulators in TMEM
//            (2 M tiles x Cout columns, double buffered)
//   warps 2-5 epilogue (tc_ptx.cuh: epilogue_chunk): TMEM -> registers (bias, residual, GN statistics) -> swizzled smem -> TMA store
#include "tc_ptx.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <stdlib.h>

namespace sdc {

constexpr int W64 = 64;
constexpr int W64_BOX_ROWS = 6;
constexpr int W64_ABYTES = W64_BOX_ROWS * W64 * 128;   // 48 KB
constexpr int W64_ASTAGES = 3;
constexpr int W64_BSTAGES = 6;
constexpr int W64_THREADS = 224;
constexpr int W64_STG = 4 * 4096;

struct Row64Params {
    int B, H, Cout, bn;
    int c0, c1;
    int items_total, items_per_cluster, items_per_image;   // item = 8 image rows of one image
    int operand_out;
    const float* bias;
    const void* residual;
    double* stats;
};

template <bool HALF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W64_THREADS, 1)
conv_row64_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out, const Row64Params p) {
    using Op = Operand<HALF>;
    using act_t = typename ActT<HALF>::type;
    constexpr int BK = Op::kBK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = (p.bn / 2) * 128;                 // weight rows staged by this CTA (multiple of 1024: bn % 16 == 0)
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    uint8_t* aring = smem;                                // [3][48 KB]
    uint8_t* bring = smem + W64_ASTAGES * W64_ABYTES;     // [6][b_bytes]
    uint8_t* staging = bring + W64_BSTAGES * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + W64_STG);
    uint64_t* a_full = bars;                       // [3]
    uint64_t* a_empty = a_full + W64_ASTAGES;      // [3]
    uint64_t* b_full = a_empty + W64_ASTAGES;      // [6]
    uint64_t* b_empty = b_full + W64_BSTAGES;      // [6]
    uint64_t* acc_full = b_empty + W64_BSTAGES;    // [2]
    uint64_t* acc_empty = acc_full + 2;            // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctot = p.c0 + p.c1;
    const int chunks = ctot / BK;
    uint32_t acc_cols = 32;
    while ((int)acc_cols < p.bn) acc_cols <<= 1;
    const int wid = (int)(blockIdx.x >> 1);
    const int item_lo = wid * p.items_per_cluster;
    const int n_items = max(0, min(p.items_total, item_lo + p.items_per_cluster) - item_lo);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        if (p.c1) tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < W64_ASTAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < W64_BSTAGES; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, 4 * acc_cols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // weight half tiles: flat sequence (item, chunk, dx, dy); tap index of the packed weights = dy * 3 + dx
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < n_items; ++it)
                for (int ch = 0; ch < chunks; ++ch)
                    for (int dx = 0; dx < 3; ++dx)
                        for (int dy = 0; dy < 3; ++dy) {
                            mbar_wait(&b_empty[s], ph ^ 1u);
                            if (leader) mbar_expect_tx(&b_full[s], (uint32_t)(2 * b_bytes));
                            tma_load_2d_2sm(bring + s * b_bytes, &map_w, &b_full[s], (dy * 3 + dx) * ctot + ch * BK, (int)rank * (p.bn / 2));
                            if (++s == W64_BSTAGES) { s = 0; ph ^= 1u; }
                        }
        }
    } else if (warp == 6) {
        if (lane == 0) {
            // activation boxes: flat sequence (item, chunk, dx)
            int s = 0;
            uint32_t ph = 0;
            for (int it = 0; it < n_items; ++it) {
                const int item = item_lo + it;
                const int b = item / p.items_per_image, h0 = 8 * (item - b * p.items_per_image) + 4 * (int)rank;
                for (int ch = 0; ch < chunks; ++ch) {
                    const int cc = ch * BK;
                    const bool second = cc >= p.c0;
                    for (int dx = 0; dx < 3; ++dx) {
                        mbar_wait(&a_empty[s], ph ^ 1u);
                        if (leader) mbar_expect_tx(&a_full[s], (uint32_t)(2 * W64_ABYTES));
                        tma_load_4d_2sm(aring + s * W64_ABYTES, second ? &map_a1 : &map_a0, &a_full[s], second ? cc - p.c0 : cc, dx - 1, h0 - 1, b);
                        if (++s == W64_ASTAGES) { s = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {   // whole warp: uniform control flow, one elected lane issues
            const uint32_t idesc = Op::idesc(p.bn, 256);
            int sa = 0, sb = 0;
            uint32_t pha = 0, phb = 0;
            for (int it = 0; it < n_items; ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * 2u * acc_cols;
                for (int cd = 0; cd < 3 * chunks; ++cd) {   // (chunk, dx)
                    mbar_wait(&a_full[sa], pha);
                    tc_fence_after();
                    const uint32_t aa0 = smem_u32(aring + sa * W64_ABYTES);
                    for (int dy = 0; dy < 3; ++dy) {
                        mbar_wait(&b_full[sb], phb);
                        tc_fence_after();
                        const uint32_t ba = smem_u32(bring + sb * b_bytes);
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                // A operand = 128 consecutive pixel rows of the box starting at image row 2j + dy
                                const uint32_t aa = aa0 + (uint32_t)((2 * j + dy) * W64 * 128);
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    Op::template mma<true>(tmem_d + (uint32_t)j * acc_cols, make_sw128_desc(aa + 32u * k), make_sw128_desc(ba + 32u * k),
                                                           idesc, (cd | dy | k) != 0);
                            }
                            umma_commit_2sm(&b_empty[sb]);
                            if (dy == 2) {
                                umma_commit_2sm(&a_empty[sa]);
                                if (cd == 3 * chunks - 1) umma_commit_2sm(&acc_full[buf]);
                            }
                        }
                        __syncwarp();
                        if (++sb == W64_BSTAGES) { sb = 0; phb ^= 1u; }
                    }
                    if (++sa == W64_ASTAGES) { sa = 0; pha ^= 1u; }
                }
            }
        }
    } else {
        // ---- epilogue: warp w may only touch TMEM lanes [32 * (w % 4), 32 * (w % 4) + 32); lane = accumulator row ----
        const int q = warp & 3;
        const uint32_t stg = smem_u32(staging + q * 4096);
        const act_t* resid = reinterpret_cast<const act_t*>(p.residual);
        const bool out_half = HALF && p.operand_out;
        for (int it = 0; it < n_items; ++it) {
            const int item = item_lo + it;
            const int b = item / p.items_per_image, h0 = 8 * (item - b * p.items_per_image) + 4 * (int)rank;
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 2u * acc_cols;
            for (int j = 0; j < 2; ++j) {
                const int m_w = (b * p.H + h0 + 2 * j) * W64 + q * 32;   // global output row of lane 0 (two image rows = 128 pixels per M tile)
                // GroupNorm partial sums exactly as the generic kernel forms them (conv_gemm.cu: two epilogue warps per lane quarter
                // taking alternate 32-column chunks, one fp32 partial each): even chunks -> (e1, e2), odd chunks -> (o1, o2)
                float e1 = 0.f, e2 = 0.f, o1 = 0.f, o2 = 0.f;
                for (int c = 0; c < p.bn; c += 32) {
                    const uint32_t taddr = tacc + (uint32_t)j * acc_cols + (uint32_t)c;
                    const act_t* rrow = resid ? resid + (size_t)(m_w + lane) * p.Cout + c : nullptr;
                    const bool odd = (c >> 5) & 1;
                    float& s1 = odd ? o1 : e1;
                    float& s2 = odd ? o2 : e2;
                    if (out_half) epilogue_chunk<true, act_t>(taddr, stg, &map_out, c, m_w, true, p.bias, rrow, false, s1, s2, lane, -1, 0, true, p.stats != nullptr);
                    else epilogue_chunk<false, act_t>(taddr, stg, &map_out, c, m_w, true, p.bias, rrow, p.operand_out != 0, s1, s2, lane, -1, 0, true, p.stats != nullptr);
                }
                if (j == 1) {   // accumulator buffer fully read -> hand it back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
                }
                if (p.stats) {
                    e1 = warp_sum(e1); e2 = warp_sum(e2); o1 = warp_sum(o1); o2 = warp_sum(o2);
                    if (lane == 0) {
                        atomicAdd(p.stats + 2 * b, (double)e1);
                        atomicAdd(p.stats + 2 * b + 1, (double)e2);
                        if (p.bn > 32) {
                            atomicAdd(p.stats + 2 * b, (double)o1);
                            atomicAdd(p.stats + 2 * b + 1, (double)o2);
                        }
                    }
                }
            }
        }
        if (lane == 0) bulk_wait<0>();   // all output stores complete before the CTA's shared memory goes away
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, 4 * acc_cols);
    }
}

}  // namespace sdc

using namespace sdc;

// Returns SDC_OK when the problem was handled here, -1 when the shape is not eligible (caller: conv_gemm).  Called by
// sdc_conv3x3_row (conv_row.cu) for W = 64.
int conv3x3_row64_launch(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                         const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout, void* stream) {
    const bool half = prec == SDC_PREC_F16;
    const int BK = half ? 64 : 32;
    static const bool enabled = []() { const char* e = getenv("SDC_ROW64"); return !(e && e[0] == '0'); }();
    if (!enabled || W != W64 || H % 8 != 0 || Cout > 128 || Cout % 32 != 0 || c0 % BK != 0 || c1 % BK != 0 || c0 <= 0) return -1;
    int n_sm = 148, dev = 0;
    SDC_CUDA(cudaGetDevice(&dev));
    SDC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    Row64Params p{};
    p.B = B; p.H = H; p.Cout = Cout; p.bn = Cout; p.c0 = c0; p.c1 = c1; p.operand_out = operand_out;
    p.bias = bias; p.residual = residual; p.stats = stats;
    p.items_per_image = H / 8;
    p.items_total = B * p.items_per_image;
    const int workers = n_sm / 2;
    // fewer items than clusters: the generic kernel's 128-pixel tiles spread over more SMs
    if (p.items_total < workers) return -1;
    SDC_REQUIRE(a0 && w_packed && out && (c1 == 0 || a1), "conv3x3_row64: null pointer");
    p.items_per_cluster = (p.items_total + workers - 1) / workers;
    const int grid = (p.items_total + p.items_per_cluster - 1) / p.items_per_cluster;

    CUtensorMap ma0, ma1, mw, mo;
    const cuuint64_t eb = half ? 2 : 4;
    auto enc_act = [&](CUtensorMap* m, const void* a, int C) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)W64, (cuuint32_t)W64_BOX_ROWS, 1};
        return encode_tmap(m, a, 4, dims, str, box, half);
    };
    int rc = enc_act(&ma0, a0, c0);
    if (rc) return rc;
    if (c1) { rc = enc_act(&ma1, a1, c1); if (rc) return rc; } else ma1 = ma0;
    const cuuint64_t ktot = (cuuint64_t)9 * (c0 + c1);
    cuuint64_t wd[2] = {ktot, (cuuint64_t)Cout};
    cuuint64_t ws[1] = {ktot * eb};
    cuuint32_t wb[2] = {(cuuint32_t)BK, (cuuint32_t)(Cout / 2)};
    rc = encode_tmap(&mw, w_packed, 2, wd, ws, wb, half);
    if (rc) return rc;
    SDC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv3x3_row64: out must be 16-byte aligned (TMA store)");
    rc = encode_out_tmap(&mo, out, (int64_t)B * H * W, Cout, half && operand_out);
    if (rc) return rc;
    const int smem_bytes = W64_ASTAGES * W64_ABYTES + W64_BSTAGES * (Cout / 2) * 128 + W64_STG + 24 * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(conv_row64_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_row64_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    cudaStream_t st = as_stream(stream);
    if (half) conv_row64_kernel<true><<<2 * grid, W64_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
    else conv_row64_kernel<false><<<2 * grid, W64_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
    SDC_LAUNCHED();
    return SDC_OK;
}
