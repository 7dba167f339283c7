// 3x3 convolution for the full-resolution level (W = 128, Cout <= 128) with ACTIVATION-HALO REUSE.
//
// Same math as conv_gemm.cu kind 1 (reference: nn.Conv2d(C, C', 3, padding=1) inside Block / Upsample2d,
// /root/reference/1D/model/unet.py:132,33-37,370), different data movement.  The generic implicit GEMM re-loads
// the activation window once per tap (9x) and, with only 128 output channels, is bound by L2->SMEM operand
// traffic (ncu: ~50 B/clk/SM, tensor pipe 30-39% active).  Here one CTA computes TWO image rows (M = 2 x 128
// pixels) per tile and, per 128-byte channel chunk (64 fp16 / 32 tf32 channels), loads the 4 x 130-pixel halo ONCE (66.5 KB, one TMA box with zero
// fill = padding); the nine taps' A operands are row-shifted views of that halo, expressed purely through the
// UMMA shared-memory descriptor start address (the 128B swizzle phase follows the absolute address).  Each weight tile is
// used by both rows.  Operand traffic drops from 64 to ~23 KB per 128x128x32 MAC block.
//
//   warp 0   TMA producer of the weight-tile ring (4 stages; 8 half tiles with CTA pairs);  warp 6: halo ring (2 stages)
//   warp 1   MMA issuer: 9 taps x 2 rows x 4 (32 bytes of K) tcgen05.mma.kind::f16|tf32 per chunk, accumulators in TMEM
//            (2 rows x Cout columns, double buffered)
//   warps 2-5 epilogue (same as conv_gemm.cu): TMEM -> registers (bias, residual, GN statistics) -> swizzled smem -> TMA store
#include "tc_ptx.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <math.h>
#include <stdlib.h>

namespace sdc {

constexpr int RW = 128;                      // image width handled by this kernel (= UMMA M)
constexpr int HALO_W = RW + 2;
constexpr int HALO_ROWS = 4 * HALO_W;        // 520 pixel rows of 128 bytes
constexpr int HALO_BYTES = HALO_ROWS * 128;  // 66,560 = 65 * 1024 (keeps every stage 1024-byte aligned)
constexpr int ROW_BSTAGES = 4;
constexpr int ROW_THREADS = 224;   // warp 0 weight TMA, 1 MMA, 2-5 epilogue, 6 halo TMA
constexpr int RSTG_BUF = 4096;             // one staging buffer per epilogue warp (32 rows x 128 bytes, TMA-store box)
constexpr int RSTG_BYTES = 4 * RSTG_BUF;   // single buffered: the halo + weight rings leave no room for a second set

struct RowParams {
    int B, H, Cout, bn;      // bn = Cout (single N tile, multiple of 32, <= 128)
    int c0, c1;
    int pairs_total, pairs_per_cta, pairs_per_image;
    int operand_out;         // 1: store as a tensor-core operand (TF32-rounded fp32 / fp16), 0: plain fp32
    int dbg;                 // experiments: bit0 = no TMA (MMA runs on whatever is in smem), bit1 = epilogue skips global stores
    const float* bias;
    const void* residual;    // operand precision
    void* out;
    double* stats;
    // Fused GroupNorm(1, C) + FiLM + SiLU (+ residual) (FP16 pair kernel, conv_row2_gn_kernel): DEFERRED EPILOGUE.  Work items
    // (4 image rows) are dealt round-robin to the clusters, so the pairs_per_image items of a sample are computed by adjacent
    // clusters in the same round.  EIGHT epilogue warps (two per TMEM lane quadrant, one image row each): a warp first reads its
    // accumulator row from TMEM for the statistics only (pass 1) and publishes its partial (sum, sum of squares) in its own 8-byte
    // slot of the sample -- the data is its own flag (slots are pre-set to 0xFF..), so there is no atomic, no fence and one L2
    // round trip; every warp then polls the sample's 64 slots with one coalesced load, reduces them in a fixed order
    // (deterministic statistics) and reads the accumulators AGAIN to store silu(GN(acc) * (1 + scale) + shift) (+ residual)
    // as fp16 (pass 2), while the MMA warp is busy with the next item in the other TMEM buffer.  The convolution output never touches HBM un-normalised
    // and is normalised from the fp32 accumulators (one rounding site fewer than conv -> fp16 -> sdc_gn_silu).  No deadlock: the
    // grid is persistent (<= one cluster per SM pair, all co-resident) and a round's counters depend only on that round's MMAs.
    int gn_apply;
    const float* gn_gamma;
    const float* gn_beta;
    const float* gn_ss;          // [n_t, ss_stride] rows (scale | shift) or null
    const int32_t* gn_tindex;    // [B] row of gn_ss per sample, or null (row 0)
    int64_t gn_ss_stride;
    const __half* gn_residual;   // [B*H*W, Cout] fp16 added after the activation, or null
    uint2* gn_slots;             // [B][128] exchange slots (partial sum, sum of squares as two floats), every byte 0xFF on entry
    int n_clusters;              // GN mode: clusters of the launch (multiple of pairs_per_image)
    // GN mode, last block of the network: instead of storing the activation, apply the 1x1 head convolution to it in registers
    // (lane = pixel, so out[o] = sum_c w[o, c] * y[c] is a per-lane dot product) and write NCHW fp32 (unet.py:178-180,378,426)
    const float* head_w;         // [head_cout, Cout] or null
    const float* head_b;         // [head_cout] or null
    float* head_out;             // [B, head_cout, H*W]
    int head_cout;               // <= 4
};

// timing experiment (SDC_ROW_DBG & 128): globaltimer stamps of cluster 0, CTA 0 -- [0..]: MMA warp (start, end per item),
// [2048..]: epilogue warp 0 (accumulators ready, pass 1 done, partners arrived, coefficients ready, pass 2 done per item)
__device__ long long g_row_trace[4096];
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ float row_silu(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// PAIR = true: cta_group::2.  A cluster of two CTAs computes FOUR image rows (CTA r: rows h0+2r, h0+2r+1); every MMA is
// M = 256 (row j of both CTAs) and each CTA stages only half of each weight tile, so the weight ring is twice as deep
// for the same shared memory and the per-SM operand traffic drops from 46 to 30 B/clk.
template <bool HALF, bool PAIR, bool GN = false>
__device__ __forceinline__ void conv_row_body(const CUtensorMap& map_a0, const CUtensorMap& map_a1, const CUtensorMap& map_w,
                                              const CUtensorMap& map_out, const RowParams& p) {
    using Op = Operand<HALF>;
    using act_t = typename ActT<HALF>::type;
    constexpr int BK = Op::kBK;   // channels per 128-byte chunk
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_bytes = (PAIR ? p.bn / 2 : p.bn) * 128;   // weight rows staged by this CTA
    constexpr int NB = PAIR ? 2 * ROW_BSTAGES : ROW_BSTAGES;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    uint8_t* halo = smem;                                   // [2][HALO_BYTES]
    uint8_t* bring = smem + 2 * HALO_BYTES;                 // [NB][b_bytes] (b_bytes multiple of 1024)
    uint8_t* staging = bring + NB * b_bytes;   // 1024-byte aligned (b_bytes is a multiple of 1024)
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + RSTG_BYTES);
    uint64_t* halo_full = bars;            // [2]
    uint64_t* halo_empty = bars + 2;       // [2]
    uint64_t* b_full = bars + 4;           // [NB]
    uint64_t* b_empty = b_full + NB;
    uint64_t* acc_full = b_empty + NB;   // [2]
    uint64_t* acc_empty = acc_full + 2;           // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
    float* gn_coef = reinterpret_cast<float*>(tmem_slot + 4);   // GN: [8 epilogue warps][A[128] | B[128]], then gamma | beta | bias (16-byte aligned)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctot = p.c0 + p.c1;
    const int chunks = ctot / BK;
    uint32_t acc_cols = 32;
    while ((int)acc_cols < p.bn) acc_cols <<= 1;
    // work item = 2 (or 4) image rows.  Plain kernels: contiguous range per CTA / cluster; GN: round-robin (item = cluster + k * clusters)
    const int wid = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int pair_lo = GN ? wid : wid * p.pairs_per_cta;
    const int pair_step = GN ? p.n_clusters : 1;
    const int n_items = GN ? (p.pairs_total - wid + p.n_clusters - 1) / p.n_clusters
                           : max(0, min(p.pairs_total, pair_lo + p.pairs_per_cta) - pair_lo);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        if (p.c1) tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < 2; ++s) { mbar_init(&halo_full[s], 1); mbar_init(&halo_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], GN ? 32 : (PAIR ? 8 : 4)); }
        fence_barrier_init();
    }
    if (warp == 1) { if constexpr (PAIR) tmem_alloc_2sm(tmem_slot, 4 * acc_cols); else tmem_alloc(tmem_slot, 4 * acc_cols); }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0 && !(p.dbg & 1)) {
            // weight-tile producer: flat sequence of (pair, chunk, tap) over the weight ring
            const int items = n_items * chunks;
            int g = 0;
            for (int item = 0; item < items; ++item) {
                const int cc = (item % chunks) * BK;
                for (int tap = 0; tap < 9; ++tap, ++g) {
                    const int s = g % NB;
                    mbar_wait(&b_empty[s], (((uint32_t)(g / NB)) & 1u) ^ 1u);
                    if constexpr (PAIR) {
                        if (leader) mbar_expect_tx(&b_full[s], (uint32_t)(2 * b_bytes));
                        tma_load_2d_2sm(bring + s * b_bytes, &map_w, &b_full[s], tap * ctot + cc, (int)rank * (p.bn / 2));
                    } else {
                        mbar_expect_tx(&b_full[s], (uint32_t)b_bytes);
                        tma_load_2d(bring + s * b_bytes, &map_w, &b_full[s], tap * ctot + cc, 0);
                    }
                }
            }
        }
    } else if (warp == (GN ? 2 : 6)) {
        if (lane == 0 && !(p.dbg & 1)) {
            // halo producer (own warp so that waiting for a free halo slot never stalls the weight ring): the halo of
            // item i+1 is requested as soon as the MMAs of item i-1 have released its slot
            const int items = n_items * chunks;
            for (int item = 0; item < items; ++item) {
                const int pair = pair_lo + (item / chunks) * pair_step, ch = item % chunks;
                const int b = pair / p.pairs_per_image, h0 = (PAIR ? 4 : 2) * (pair - b * p.pairs_per_image) + 2 * (int)rank;
                const int cc = ch * BK;
                const bool second = cc >= p.c0;
                const int hs = item & 1;
                mbar_wait(&halo_empty[hs], (((uint32_t)(item >> 1)) & 1u) ^ 1u);
                if constexpr (PAIR) {
                    if (leader) mbar_expect_tx(&halo_full[hs], (uint32_t)(2 * HALO_BYTES));
                    tma_load_4d_2sm(halo + hs * HALO_BYTES, second ? &map_a1 : &map_a0, &halo_full[hs], second ? cc - p.c0 : cc, -1, h0 - 1, b);
                } else {
                    mbar_expect_tx(&halo_full[hs], (uint32_t)HALO_BYTES);
                    tma_load_4d(halo + hs * HALO_BYTES, second ? &map_a1 : &map_a0, &halo_full[hs], second ? cc - p.c0 : cc, -1, h0 - 1, b);
                }
            }
        }
    } else if (warp == 1) {
        if (leader) {   // whole warp: uniform control flow, one elected lane issues
            const uint32_t idesc = Op::idesc(p.bn, (PAIR ? 2 : 1) * RW);
            int g = 0, item = 0;
            for (int it = 0; it < n_items; ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
                if (GN && (p.dbg & 128) && blockIdx.x == 0 && lane == 0 && it < 512) g_row_trace[2 * it] = gtime();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * 2u * acc_cols;
                for (int ch = 0; ch < chunks; ++ch, ++item) {
                    const int hs = item & 1;
                    if (!(p.dbg & 1)) mbar_wait(&halo_full[hs], ((uint32_t)(item >> 1)) & 1u);
                    tc_fence_after();
                    const uint32_t ha = smem_u32(halo + hs * HALO_BYTES);
                    for (int tap = 0; tap < 9; ++tap, ++g) {
                        const int s = g % NB;
                        if (!(p.dbg & 1)) mbar_wait(&b_full[s], ((uint32_t)(g / NB)) & 1u);
                        tc_fence_after();
                        const uint32_t ba = smem_u32(bring + s * b_bytes);
                        const int dy = tap / 3, dx = tap - 3 * dy;
                        if (elect_one()) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                // A operand = 128 consecutive halo pixels starting at (row j+dy, col dx)
                                const uint32_t aa = ha + (uint32_t)(((j + dy) * HALO_W + dx) * 128);
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    Op::template mma<PAIR>(tmem_d + (uint32_t)j * acc_cols, make_sw128_desc(aa + 32u * k),
                                                           make_sw128_desc(ba + 32u * k), idesc, (ch | tap | k) != 0);
                                }
                            }
                            if constexpr (PAIR) umma_commit_2sm(&b_empty[s]); else umma_commit(&b_empty[s]);
                            if (tap == 8) {
                                if constexpr (PAIR) umma_commit_2sm(&halo_empty[hs]); else umma_commit(&halo_empty[hs]);
                                if (ch == chunks - 1) {
                                    if constexpr (PAIR) umma_commit_2sm(&acc_full[buf]); else umma_commit(&acc_full[buf]);
                                    if (GN && (p.dbg & 128) && blockIdx.x == 0 && it < 512) g_row_trace[2 * it + 1] = gtime();
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (GN && warp == 3) {
        // spare warp: keeps the 16 epilogue warps aligned to the TMEM lane quadrants (warp & 3)
    } else if constexpr (GN) {
        // ====== fused GroupNorm epilogue: 16 warps (4 per scheduler), warp = TMEM lane quadrant q x image row j x channel half hf ======
        const int ew = warp - 4, q = ew & 3, j = (ew >> 3) & 1, hf = (ew >> 2) & 1;
        const int C = p.Cout;
        const int c_split = ((C / 32 + 1) / 2) * 32, c_lo = hf ? c_split : 0, c_hi = hf ? C : c_split;   // this warp's channels
        // per-kernel constants in shared memory (generic loads queue behind the kernel's own TMA traffic: ~1 us per L1 miss)
        float* cgam = gn_coef + 2048;   // gamma[128] | beta[128] | bias[128], written by epilogue warp 0, read after the named barrier
        if (ew == 0) {
            for (int c = lane; c < C; c += 32) {
                cgam[c] = p.gn_gamma[c];
                cgam[128 + c] = p.gn_beta[c];
                cgam[256 + c] = p.bias ? p.bias[c] : 0.f;
            }
        }
        asm volatile("bar.sync 1, 512;" ::: "memory");   // the 16 epilogue warps only
        float* cA = gn_coef + ew * 128 - c_lo;   // this warp's A[c_lo .. c_hi) | B[c_lo .. c_hi): 64 + 64 floats
        float* cB = cA + 64;
        const uint32_t cA_s = smem_u32(cA), cB_s = smem_u32(cB), cbias_s = smem_u32(cgam + 256);
        const uint32_t need_slots = 32u * (uint32_t)p.pairs_per_image;   // <= 128
        const double inv_cnt = 1.0 / ((double)p.H * RW * (double)C);
        __half* outp = reinterpret_cast<__half*>(p.out);
        for (int it = 0; it < n_items; ++it) {
            const int pair = pair_lo + it * pair_step;
            const int b = pair / p.pairs_per_image, in_sample = pair - b * p.pairs_per_image, h0 = 4 * in_sample + 2 * (int)rank;
            const int buf = it & 1;
            const int m_w = (b * p.H + h0 + j) * RW + q * 32;   // global output row of lane 0
            // ---- prologue, overlapped with the MMAs of this item: everything that does not depend on the statistics ----
            float psc[2], psh[2];   // FiLM (scale + 1, shift) of channels c_lo + lane + 32 i
#pragma unroll
            for (int i = 0; i < 2; ++i) { psc[i] = 1.f; psh[i] = 0.f; }
            if (p.gn_ss) {
                const float* ss = p.gn_ss + (int64_t)(p.gn_tindex ? __ldg(p.gn_tindex + b) : 0) * p.gn_ss_stride;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int c = c_lo + lane + 32 * i;
                    if (c < c_hi) { psc[i] = __ldg(ss + c) + 1.0f; psh[i] = __ldg(ss + C + c); }
                }
            }
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 2u * acc_cols + (uint32_t)j * acc_cols;
            const bool tr = (p.dbg & 128) && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 400;
            if (tr) g_row_trace[2048 + 5 * it] = gtime();
            // ---- pass 1: statistics of this warp's 32 pixels x 64 channels of accumulators (+ bias) ----
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = c_lo; c < ((p.dbg & 8) ? 0 : c_hi); c += 32) {   // dbg 8 (timing experiment): no pass 1
                uint32_t r[32];
                tmem_ld32(trow + (uint32_t)c, r);
                float a1 = 0.f, a2 = 0.f;
#pragma unroll
                for (int k = 0; k < 32; k += 4) {
                    const float4 bb = lds128(cbias_s + (uint32_t)(c + k) * 4u);
                    const float x0 = __uint_as_float(r[k]) + bb.x, x1 = __uint_as_float(r[k + 1]) + bb.y;
                    const float x2 = __uint_as_float(r[k + 2]) + bb.z, x3 = __uint_as_float(r[k + 3]) + bb.w;
                    a1 += (x0 + x1) + (x2 + x3);
                    a2 = fmaf(x0, x0, a2); a2 = fmaf(x1, x1, a2); a2 = fmaf(x2, x2, a2); a2 = fmaf(x3, x3, a2);
                }
                s1 += a1;
                s2 += a2;
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (tr) g_row_trace[2048 + 5 * it + 1] = gtime();
            uint2* slots = p.gn_slots + (size_t)b * 128;
            if (lane == 0) {
                uint32_t w1 = __float_as_uint(s1);
                if (w1 == 0xFFFFFFFFu) w1 = 0x7FFFFFFFu;   // keep the "empty" pattern unique (still a NaN)
                // ONE 8-byte scalar store (single-copy atomic): a reader sees the slot empty or complete
                const unsigned long long wv = (unsigned long long)w1 | ((unsigned long long)__float_as_uint(s2) << 32);
                asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slots + in_sample * 32 + (int)rank * 16 + ew), "l"(wv) : "memory");
            }
            // this lane's residual row (<= 2 chunks x 32 fp16 channels): in flight while the partners arrive
            uint4 rq[2][4];
            if (p.gn_residual) {
                const uint4* rrow = reinterpret_cast<const uint4*>(p.gn_residual + (size_t)(m_w + lane) * C + c_lo);
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
                    if (c_lo + cc * 32 < c_hi) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) rq[cc][k] = __ldg(rrow + cc * 4 + k);
                    }
            }
            // wait for every epilogue warp of the sample (items x 2 CTAs x 16 warps): lane l polls slots 4l .. 4l + 3
            float f1 = 0.f, f2 = 0.f;
            for (;;) {
                unsigned long long w0, w1, w2, w3;
                const uint4* sp = reinterpret_cast<const uint4*>(slots) + 2 * lane;
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(sp) : "memory");
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w2), "=l"(w3) : "l"(sp + 1) : "memory");
                const bool ok = (4u * (uint32_t)lane >= need_slots) ||
                                ((uint32_t)w0 != 0xFFFFFFFFu && (uint32_t)w1 != 0xFFFFFFFFu && (uint32_t)w2 != 0xFFFFFFFFu && (uint32_t)w3 != 0xFFFFFFFFu);
                if (__all_sync(0xffffffffu, ok) || (p.dbg & 4)) {   // dbg 4 (timing experiment): do not wait for the partners
                    if (4u * (uint32_t)lane < need_slots) {
                        f1 = (__uint_as_float((uint32_t)w0) + __uint_as_float((uint32_t)w1)) + (__uint_as_float((uint32_t)w2) + __uint_as_float((uint32_t)w3));
                        f2 = (__uint_as_float((uint32_t)(w0 >> 32)) + __uint_as_float((uint32_t)(w1 >> 32))) +
                             (__uint_as_float((uint32_t)(w2 >> 32)) + __uint_as_float((uint32_t)(w3 >> 32)));
                    }
                    break;
                }
                __nanosleep(64);
            }
            if (tr) g_row_trace[2048 + 5 * it + 2] = gtime();
            // fixed-order pairwise tree over the 128 fp32 partials (deterministic; FP64 is avoided on purpose: ncu showed the
            // DADDs of a double reduction as the top math stall of this warp-starved epilogue), then mean / variance in double
            f1 = warp_sum(f1);
            f2 = warp_sum(f2);
            const double d1 = (double)f1, d2 = (double)f2;
            const double mean_d = d1 * inv_cnt;
            double var_d = d2 * inv_cnt - mean_d * mean_d;
            if (var_d < 0.0) var_d = 0.0;
            const float mean = (float)mean_d, rstd = 1.0f / sqrtf((float)var_d + 1e-5f);
            if (in_sample == 0 && rank == 0 && ew == 0 && lane == 0) { p.stats[2 * b] = d1; p.stats[2 * b + 1] = d2; }
            // per-channel affine of this sample: y = silu(A_c * acc + B_c); the conv bias is folded into B_c
            __syncwarp();   // the previous item's pass 2 has finished reading the coefficients
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int c = c_lo + lane + 32 * i;
                if (c < c_hi) {
                    const float g = cgam[c] * rstd;
                    const float a = g * psc[i];
                    float bb = (cgam[128 + c] - mean * g) * psc[i] + psh[i];
                    bb = fmaf(a, cgam[256 + c], bb);
                    cA[c] = a;
                    cB[c] = bb;
                }
            }
            __syncwarp();
            if (tr) g_row_trace[2048 + 5 * it + 3] = gtime();
            // ---- pass 2: normalise from TMEM; store fp16 rows straight from registers (or feed the head convolution) ----
            float hd[4] = {0.f, 0.f, 0.f, 0.f};
            __half* orow = outp + (size_t)(m_w + lane) * C;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int c = c_lo + cc * 32;
                if (c < c_hi) {
                    uint32_t r[32];
                    tmem_ld32(trow + (uint32_t)c, r);
                    float v[32];
#pragma unroll
                    for (int k = 0; k < 32; k += 4) {
                        const float4 a4 = lds128(cA_s + (uint32_t)(c + k) * 4u), b4 = lds128(cB_s + (uint32_t)(c + k) * 4u);
                        v[k] = fmaf(__uint_as_float(r[k]), a4.x, b4.x);
                        v[k + 1] = fmaf(__uint_as_float(r[k + 1]), a4.y, b4.y);
                        v[k + 2] = fmaf(__uint_as_float(r[k + 2]), a4.z, b4.z);
                        v[k + 3] = fmaf(__uint_as_float(r[k + 3]), a4.w, b4.w);
                        if (!(p.dbg & 16)) {   // dbg 16 (timing experiment): no SiLU
                            v[k] = row_silu(v[k]); v[k + 1] = row_silu(v[k + 1]); v[k + 2] = row_silu(v[k + 2]); v[k + 3] = row_silu(v[k + 3]);
                        }
                    }
                    if (p.gn_residual) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const __half2* h2 = reinterpret_cast<const __half2*>(&rq[cc][k]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(h2[e]);
                                v[8 * k + 2 * e] += f.x;
                                v[8 * k + 2 * e + 1] += f.y;
                            }
                        }
                    }
                    if (p.head_out) {
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            if (o < p.head_cout) {
                                const float* wrow = p.head_w + o * C + c;
#pragma unroll
                                for (int k = 0; k < 32; k += 4) {
                                    const float4 w4 = __ldg(reinterpret_cast<const float4*>(wrow + k));
                                    hd[o] = fmaf(v[k], w4.x, hd[o]); hd[o] = fmaf(v[k + 1], w4.y, hd[o]);
                                    hd[o] = fmaf(v[k + 2], w4.z, hd[o]); hd[o] = fmaf(v[k + 3], w4.w, hd[o]);
                                }
                            }
                        }
                    } else if (!(p.dbg & 32)) {   // dbg 32 (timing experiment): no output stores
                        // 64 contiguous bytes per lane (two full 32-byte sectors): no staging tile, no TMA-store round trip
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint4 o;
                            __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                            for (int e = 0; e < 4; ++e) o2[e] = __floats2half2_rn(v[8 * k + 2 * e], v[8 * k + 2 * e + 1]);
                            *reinterpret_cast<uint4*>(orow + c + 8 * k) = o;
                        }
                    }
                }
            }
            if (p.head_out) {
                const int px = (h0 + j) * RW + q * 32 + lane;
#pragma unroll
                for (int o = 0; o < 4; ++o)
                    if (o < p.head_cout)   // two warps (channel halves) per pixel: x + y onto the caller's zeros is order independent
                        atomicAdd(p.head_out + ((size_t)b * p.head_cout + o) * ((size_t)p.H * RW) + px,
                                  hd[o] + ((hf == 0 && p.head_b) ? __ldg(p.head_b + o) : 0.f));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&acc_empty[buf]);
            if (tr) g_row_trace[2048 + 5 * it + 4] = gtime();
        }
    } else {
        const int q = warp & 3;
        const uint32_t stg = smem_u32(staging + q * RSTG_BUF);
        const act_t* resid = reinterpret_cast<const act_t*>(p.residual);
        const bool out_half = HALF && p.operand_out;
        for (int it = 0; it < n_items; ++it) {
            const int pair = pair_lo + it * pair_step;
            const int b = pair / p.pairs_per_image, h0 = (PAIR ? 4 : 2) * (pair - b * p.pairs_per_image) + 2 * (int)rank;
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 2u * acc_cols;
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < 2; ++j) {
                if (h0 + j >= p.H || (p.dbg & 2)) continue;   // phantom row of an odd-height image (warp-uniform)
                const int m_w = (b * p.H + h0 + j) * RW + q * 32;   // global output row of lane 0
                for (int c = 0; c < p.bn; c += 32) {
                    if (lane == 0) bulk_wait_read<0>();   // the previous store has finished reading the staging buffer
                    __syncwarp();
                    const uint32_t taddr = tacc + (uint32_t)j * acc_cols + (uint32_t)c;
                    const act_t* rrow = resid ? resid + (size_t)(m_w + lane) * p.Cout + c : nullptr;
                    if (out_half) epilogue_chunk<true, act_t>(taddr, stg, &map_out, c, m_w, true, p.bias, rrow, false, s1, s2, lane);
                    else epilogue_chunk<false, act_t>(taddr, stg, &map_out, c, m_w, true, p.bias, rrow, p.operand_out != 0, s1, s2, lane);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&acc_empty[buf]); else mbar_arrive(&acc_empty[buf]); }
            if (p.stats) {
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                if (lane == 0 && h0 < p.H) {
                    atomicAdd(p.stats + 2 * b, (double)s1);
                    atomicAdd(p.stats + 2 * b + 1, (double)s2);
                }
            }
        }
        if (lane == 0) bulk_wait<0>();   // all output stores complete before the CTA's shared memory goes away
        tc_fence_before();
    }
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 4 * acc_cols); else tmem_dealloc(tmem_base, 4 * acc_cols);
    }
}

template <bool HALF>
__global__ void __launch_bounds__(ROW_THREADS, 1)
conv_row_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out, const RowParams p) {
    conv_row_body<HALF, false>(map_a0, map_a1, map_w, map_out, p);
}
template <bool HALF>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ROW_THREADS, 1)
conv_row2_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out, const RowParams p) {
    conv_row_body<HALF, true>(map_a0, map_a1, map_w, map_out, p);
}

constexpr int ROW_GN_THREADS = 640;   // warps 0 weights, 1 MMA, 2 halo, 3 spare, 4-19 epilogue (4 per scheduler and TMEM lane quadrant)
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ROW_GN_THREADS, 1)
conv_row2_gn_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                    const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out, const RowParams p) {
    conv_row_body<true, true, true>(map_a0, map_a1, map_w, map_out, p);
}

}  // namespace sdc

using namespace sdc;

struct RowGn {
    const float* gamma; const float* beta; const float* ss; const int32_t* tindex; int64_t ss_stride; const void* residual;
    void* slots; const float* head_w; const float* head_b; float* head_out; int head_cout;
};

// W = 64 variant (conv_row64.cu): one activation box per horizontal tap, vertical taps as row-shifted views
int conv3x3_row64_launch(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                         const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout, void* stream);

// Returns SDC_OK when the problem was handled here, -1 when the shape is not eligible (caller uses conv_gemm).
static int conv3x3_row_launch(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                              const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout,
                              const RowGn* gn, void* stream) {
    SDC_REQUIRE(prec == SDC_PREC_TF32 || prec == SDC_PREC_F16, "conv3x3_row: precision %d", prec);
    const bool half = prec == SDC_PREC_F16;
    const int BK = half ? 64 : 32;
    if (W == 64 && !gn) return conv3x3_row64_launch(prec, a0, c0, a1, c1, w_packed, bias, residual, out, stats, operand_out, B, H, W, Cout, stream);
    if (W != RW || Cout > 128 || Cout % 32 != 0 || c0 % BK != 0 || c1 % BK != 0 || c0 <= 0) return -1;
    SDC_REQUIRE(a0 && w_packed && out && B > 0 && H > 0 && (c1 == 0 || a1), "conv3x3_row: bad arguments");
    RowParams p{};
    p.B = B; p.H = H; p.Cout = Cout; p.bn = Cout; p.c0 = c0; p.c1 = c1; p.operand_out = operand_out;
    p.bias = bias; p.residual = residual; p.out = out; p.stats = stats;
    if (gn) {
        p.gn_apply = 1; p.gn_gamma = gn->gamma; p.gn_beta = gn->beta; p.gn_ss = gn->ss; p.gn_tindex = gn->tindex;
        p.gn_ss_stride = gn->ss_stride; p.gn_residual = (const __half*)gn->residual; p.gn_slots = (uint2*)gn->slots;
        p.head_w = gn->head_w; p.head_b = gn->head_b; p.head_out = gn->head_out; p.head_cout = gn->head_cout;
    }
    { const char* e = getenv("SDC_ROW_DBG"); p.dbg = e ? atoi(e) : 0; }
    int n_sm = 148, dev = 0;
    SDC_CUDA(cudaGetDevice(&dev));
    SDC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    static const bool allow_pair = []() { const char* e = getenv("SDC_NO_2CTA"); return !(e && e[0] == '1'); }();
    // (the fused-GroupNorm kernel exists as a CTA-pair kernel only: small batches simply use fewer clusters)
    const bool pair = allow_pair && Cout % 32 == 0 && (Cout / 2) % 8 == 0 && (gn || B * ((H + 3) / 4) >= n_sm / 2);
    p.pairs_per_image = pair ? (H + 3) / 4 : (H + 1) / 2;
    p.pairs_total = B * p.pairs_per_image;
    const int workers = pair ? n_sm / 2 : n_sm;
    const int ctas = p.pairs_total < workers ? p.pairs_total : workers;
    p.pairs_per_cta = (p.pairs_total + ctas - 1) / ctas;
    int grid = (p.pairs_total + p.pairs_per_cta - 1) / p.pairs_per_cta;
    if (gn) {
        // deferred epilogue: FP16 pair kernel, whole 4-row items, the items of a sample on adjacent clusters of the same round
        if (!pair || !half || !operand_out || !stats || residual || H % 4 != 0 || Cout % 32 != 0) return -1;
        SDC_REQUIRE(gn->slots, "conv3x3_row_gn: null exchange slots");
        if (H > 16) return -1;   // 32 slots per 4-row item, 128 per sample
        SDC_REQUIRE(!gn->head_out || (gn->head_w && gn->head_cout >= 1 && gn->head_cout <= 4), "conv3x3_row_gn: bad head arguments");
        if (p.pairs_per_image > workers) return -1;
        grid = p.pairs_total < workers ? p.pairs_total : workers / p.pairs_per_image * p.pairs_per_image;
        p.n_clusters = grid;
    }

    CUtensorMap ma0, ma1, mw;
    const cuuint64_t eb = half ? 2 : 4;
    auto enc_act = [&](CUtensorMap* m, const void* a, int C) {
        cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
        cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)HALO_W, 4, 1};
        return encode_tmap(m, a, 4, dims, str, box, half);
    };
    int rc = enc_act(&ma0, a0, c0);
    if (rc) return rc;
    if (c1) { rc = enc_act(&ma1, a1, c1); if (rc) return rc; } else ma1 = ma0;
    const cuuint64_t ktot = (cuuint64_t)9 * (c0 + c1);
    cuuint64_t wd[2] = {ktot, (cuuint64_t)Cout};
    cuuint64_t ws[1] = {ktot * eb};
    cuuint32_t wb[2] = {(cuuint32_t)BK, (cuuint32_t)(pair ? Cout / 2 : Cout)};
    rc = encode_tmap(&mw, w_packed, 2, wd, ws, wb, half);
    if (rc) return rc;
    SDC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv3x3_row: out must be 16-byte aligned (TMA store)");
    CUtensorMap mo;
    rc = encode_out_tmap(&mo, out, (int64_t)B * H * W, Cout, half && operand_out);
    if (rc) return rc;
    const int smem_bytes = 2 * HALO_BYTES + ROW_BSTAGES * Cout * 128 + RSTG_BYTES + 24 * 8 + 16 + (gn ? 8192 + 1536 : 0) + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(conv_row_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_row2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_row_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_row2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    cudaStream_t st = as_stream(stream);
    if (gn) {
        static bool gn_attr = false;
        if (!gn_attr) { SDC_CUDA(cudaFuncSetAttribute(conv_row2_gn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); gn_attr = true; }
        conv_row2_gn_kernel<<<2 * grid, ROW_GN_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
    } else if (pair) {
        if (half) conv_row2_kernel<true><<<2 * grid, ROW_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
        else conv_row2_kernel<false><<<2 * grid, ROW_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
    } else {
        if (half) conv_row_kernel<true><<<grid, ROW_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
        else conv_row_kernel<false><<<grid, ROW_THREADS, smem_bytes, st>>>(ma0, ma1, mw, mo, p);
    }
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_debug_row_trace(long long* host_out) {   // timing experiment only (not declared in the public header)
    SDC_CUDA(cudaMemcpyFromSymbol(host_out, g_row_trace, sizeof(long long) * 4096));
    return SDC_OK;
}

extern "C" int sdc_conv3x3_row(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                               const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout,
                               void* stream) {
    return conv3x3_row_launch(prec, a0, c0, a1, c1, w_packed, bias, residual, out, stats, operand_out, B, H, W, Cout, nullptr, stream);
}

extern "C" int sdc_conv3x3_row_gn(const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias, void* out,
                                  double* stats, void* sync_slots, const float* gamma, const float* beta, const float* scale_shift,
                                  const int32_t* t_index, int64_t ss_stride, const void* gn_residual, int B, int H, int W, int Cout,
                                  void* stream) {
    SDC_REQUIRE(gamma && beta && stats && out, "conv3x3_row_gn: null GroupNorm arguments");
    RowGn gn{gamma, beta, scale_shift, t_index, ss_stride, gn_residual, sync_slots, nullptr, nullptr, nullptr, 0};
    return conv3x3_row_launch(SDC_PREC_F16, a0, c0, a1, c1, w_packed, bias, nullptr, out, stats, 1, B, H, W, Cout, &gn, stream);
}

extern "C" int sdc_conv3x3_row_gn_head(const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                                       double* stats, void* sync_slots, const float* gamma, const float* beta,
                                       const void* gn_residual, const float* head_w, const float* head_b, float* out_nchw,
                                       int head_cout, int B, int H, int W, int Cout, void* stream) {
    SDC_REQUIRE(gamma && beta && stats && head_w && out_nchw, "conv3x3_row_gn_head: null arguments");
    RowGn gn{gamma, beta, nullptr, nullptr, 0, gn_residual, sync_slots, head_w, head_b, out_nchw, head_cout};
    // `out` is only used to build the (unused) store map: any valid 16-byte aligned device address of the activations will do
    return conv3x3_row_launch(SDC_PREC_F16, a0, c0, a1, c1, w_packed, bias, nullptr, const_cast<void*>(a0), stats, 1, B, H, W, Cout, &gn, stream);
}
