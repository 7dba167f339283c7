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
    // Fused GroupNorm(1, C) + FiLM + SiLU (+ residual) apply (FP16 pair kernel only, conv_row2_gn_kernel): a cluster owns whole
    // samples (its work items are sample aligned), so once the last rows of a sample have been stored and both CTAs' statistics
    // have landed, eight extra warps re-read the CTA's rows of the sample -- still resident in L2 -- and normalise them IN PLACE;
    // the separate read-modify-write pass over HBM (sdc_gn_silu) disappears.
    // MEASURED (B = 1024, 128 -> 128, scripts/time_row_gn.py): correct, but NOT faster yet -- 640 us fused against 413 us conv +
    // 190 us GroupNorm kernel.  The eight warps need 43 us per sample (the convolution: 29 us): their loads queue behind the
    // kernel's own TMA traffic (two 66 KB halo boxes + the weight ring in flight per SM), ~2.5 us per round trip with at most
    // 48 KB of their own requests outstanding.  Off by default (Unet2D.fuse_groupnorm); kept as the starting point for a version
    // that stages the sample through shared memory with bulk copies once the halo ring is shrunk.
    int gn_apply;
    const float* gn_gamma;
    const float* gn_beta;
    const float* gn_ss;          // [n_t, ss_stride] rows (scale | shift) or null
    const int32_t* gn_tindex;    // [B] row of gn_ss per sample, or null (row 0)
    int64_t gn_ss_stride;
    const __half* gn_residual;   // [B*H*W, Cout] fp16 added after the activation, or null
};

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
    // fused GroupNorm: monotonic count of epilogue-warp arrivals (both CTAs of the pair), 8 per finished sample
    uint32_t* gn_ready = reinterpret_cast<uint32_t*>(acc_empty + 2);
    uint32_t* tmem_slot = gn_ready + 2;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ctot = p.c0 + p.c1;
    const int chunks = ctot / BK;
    uint32_t acc_cols = 32;
    while ((int)acc_cols < p.bn) acc_cols <<= 1;
    const int pair_lo = (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x) * p.pairs_per_cta;   // work item = 2 (or 4) image rows
    const int pair_hi = min(p.pairs_total, pair_lo + p.pairs_per_cta);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        if (p.c1) tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_out);
        for (int s = 0; s < 2; ++s) { mbar_init(&halo_full[s], 1); mbar_init(&halo_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], PAIR ? 8 : 4); }
        *gn_ready = 0u;
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
            const int items = (pair_hi - pair_lo) * chunks;
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
    } else if (warp == 6) {
        if (lane == 0 && !(p.dbg & 1)) {
            // halo producer (own warp so that waiting for a free halo slot never stalls the weight ring): the halo of
            // item i+1 is requested as soon as the MMAs of item i-1 have released its slot
            const int items = (pair_hi - pair_lo) * chunks;
            for (int item = 0; item < items; ++item) {
                const int pair = pair_lo + item / chunks, ch = item % chunks;
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
            int g = 0, item = 0, it = 0;
            for (int pair = pair_lo; pair < pair_hi; ++pair, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);
                tc_fence_after();
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
                                if (ch == chunks - 1) { if constexpr (PAIR) umma_commit_2sm(&acc_full[buf]); else umma_commit(&acc_full[buf]); }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (GN && warp >= 7) {
        // ---- GroupNorm warps (8): per finished sample, normalise this CTA's rows of it in place (L2 resident) ----
        // 256 threads; thread -> one 8-channel group and the pixels pl, pl + pstep, ... of each of the CTA's 8 image rows of the
        // sample: n_vec 16-byte vectors, processed 4 at a time with the next 4 (and their residuals) already in flight.
        if constexpr (GN) {
            const int C = p.Cout, c8n = C >> 3;
            const int te = (warp - 7) * 32 + lane;          // 0..255
            const int cgp = te % c8n, pl = te / c8n, pstep = 256 / c8n;   // 8-channel group, pixel lane, pixels per pass
            const int vpr = RW / pstep;                     // vectors per image row per thread
            const int n_vec = 2 * p.pairs_per_image * vpr;  // multiple of 8 (host check)
            __half* outp = reinterpret_cast<__half*>(p.out);
            const __half* resp = p.gn_residual;
            const double cnt = (double)p.H * RW * (double)C;
            const int s_lo = pair_lo / p.pairs_per_image, s_hi = (pair_hi + p.pairs_per_image - 1) / p.pairs_per_image;
            for (int b = s_lo; b < s_hi; ++b) {
                const uint32_t need = 8u * (uint32_t)(b - s_lo + 1);
                uint32_t have;
                do {
                    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(smem_u32(gn_ready)) : "memory");
                    if (have < need) __nanosleep(100);
                } while (have < need);
                if (p.dbg & 4) continue;   // experiment: synchronisation only
                const double mean_d = __ldcg(p.stats + 2 * b) / cnt;
                double var_d = __ldcg(p.stats + 2 * b + 1) / cnt - mean_d * mean_d;
                if (var_d < 0.0) var_d = 0.0;
                const float mean = (float)mean_d, rstd = (float)(1.0 / sqrt(var_d + 1e-5));
                const float* ss = p.gn_ss ? p.gn_ss + (int64_t)(p.gn_tindex ? p.gn_tindex[b] : 0) * p.gn_ss_stride : nullptr;
                float A[8], Bc[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int c = cgp * 8 + k;
                    const float g = p.gn_gamma[c] * rstd;
                    float a = g, bb = p.gn_beta[c] - mean * g;
                    if (ss) {
                        const float sc = ss[c] + 1.0f;
                        a *= sc;
                        bb = bb * sc + ss[C + c];
                    }
                    A[k] = a;
                    Bc[k] = bb;
                }
                // vector i of this thread: CTA row r = i / vpr (item r / 2, row r % 2 of the item), pixel pl + (i % vpr) * pstep
                auto offset = [&](int i) -> int64_t {
                    const int r = i / vpr, it = i - r * vpr;
                    const int hr = 4 * (r >> 1) + 2 * (int)rank + (r & 1);
                    return ((int64_t)(b * p.H + hr) * RW + pl + it * pstep) * C + cgp * 8;
                };
                auto load4 = [&](int i0, uint4 (&v)[4], uint4 (&r)[4]) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t off = offset(i0 + u);
                        v[u] = __ldcg(reinterpret_cast<const uint4*>(outp + off));
                        r[u] = resp ? __ldcs(reinterpret_cast<const uint4*>(resp + off)) : make_uint4(0u, 0u, 0u, 0u);
                    }
                };
                auto apply4 = [&](int i0, const uint4 (&v)[4], const uint4 (&r)[4]) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const __half2* h2 = reinterpret_cast<const __half2*>(&v[u]);
                        const __half2* r2 = reinterpret_cast<const __half2*>(&r[u]);
                        uint4 o;
                        __half2* o2 = reinterpret_cast<__half2*>(&o);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(h2[e]);
                            const float2 rr = __half22float2(r2[e]);
                            const float y0 = row_silu(fmaf(f.x, A[2 * e], Bc[2 * e])) + rr.x;
                            const float y1 = row_silu(fmaf(f.y, A[2 * e + 1], Bc[2 * e + 1])) + rr.y;
                            o2[e] = __floats2half2_rn(y0, y1);
                        }
                        if (!(p.dbg & 8) || o.x == 0x12345678u) *reinterpret_cast<uint4*>(outp + offset(i0 + u)) = o;   // dbg 8: no stores
                    }
                };
                if (resp) {
                    uint4 va[4], ra[4], vb[4], rb[4];
                    load4(0, va, ra);
                    for (int i = 0; i < n_vec; i += 8) {
                        load4(i + 4, vb, rb);
                        apply4(i, va, ra);
                        if (i + 8 < n_vec) load4(i + 8, va, ra);
                        apply4(i + 4, vb, rb);
                    }
                } else {
                    // no residual: the freed registers hold a third group in flight
                    const uint4 z4[4] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
                    uint4 va[4], vb[4], vc[4], vd[4];
                    auto loadv = [&](int i0, uint4 (&v)[4]) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = __ldcg(reinterpret_cast<const uint4*>(outp + offset(i0 + u)));
                    };
                    loadv(0, va);
                    loadv(4, vb);
                    for (int i = 0; i < n_vec; i += 16) {   // n_vec is a multiple of 16 for C >= 32 ... checked on the host
                        loadv(i + 8, vc);
                        apply4(i, va, z4);
                        loadv(i + 12, vd);
                        apply4(i + 4, vb, z4);
                        if (i + 16 < n_vec) loadv(i + 16, va);
                        apply4(i + 8, vc, z4);
                        if (i + 16 < n_vec) loadv(i + 20, vb);
                        apply4(i + 12, vd, z4);
                    }
                }
            }
        }
    } else {
        const int q = warp & 3;
        const uint32_t stg = smem_u32(staging + q * RSTG_BUF);
        const act_t* resid = reinterpret_cast<const act_t*>(p.residual);
        const bool out_half = HALF && p.operand_out;
        int it = 0;
        for (int pair = pair_lo; pair < pair_hi; ++pair, ++it) {
            const int b = pair / p.pairs_per_image, h0 = (PAIR ? 4 : 2) * (pair - b * p.pairs_per_image) + 2 * (int)rank;
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            float s1 = 0.f, s2 = 0.f;
            for (int j = 0; j < 2; ++j) {
                if (h0 + j >= p.H || (p.dbg & 2)) continue;   // phantom row of an odd-height image (warp-uniform)
                const int m_w = (b * p.H + h0 + j) * RW + q * 32;   // global output row of lane 0
                for (int c = 0; c < p.bn; c += 32) {
                    if (lane == 0) bulk_wait_read<0>();   // the previous store has finished reading the staging buffer
                    __syncwarp();
                    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * 2u * acc_cols + (uint32_t)j * acc_cols + (uint32_t)c;
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
            if constexpr (GN) {
                // Publish finished samples to the GroupNorm warps of both CTAs (they need every row of their own CTA and the
                // statistics of both).  A sample is published one item LATE -- after the first item of the next sample, when its
                // stores have long completed and wait_group returns at once -- except for the cluster's last sample.
                const int in_sample = (pair - pair_lo) % p.pairs_per_image;
                const bool last_of_cluster = pair + 1 == pair_hi;
                if (lane == 0 && ((in_sample == 0 && pair > pair_lo) || last_of_cluster)) {
                    const int n_pub = (in_sample == 0 && pair > pair_lo ? 1 : 0) + (last_of_cluster && in_sample == p.pairs_per_image - 1 ? 1 : 0);
                    if (last_of_cluster) bulk_wait<0>();
                    else if (p.bn == 128) bulk_wait<8>();     // the 2 x bn/32 store groups of the current item may still be pending
                    else if (p.bn == 64) bulk_wait<4>();
                    else if (p.bn == 32) bulk_wait<2>();
                    else bulk_wait<0>();
                    asm volatile("fence.proxy.async;" ::: "memory");   // async-proxy (TMA) writes before generic-proxy reads
                    uint32_t peer;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peer) : "r"(smem_u32(gn_ready)), "r"(rank ^ 1u));
                    asm volatile("red.release.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(smem_u32(gn_ready)), "r"(n_pub) : "memory");
                    asm volatile("red.release.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(peer), "r"(n_pub) : "memory");
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

constexpr int ROW_GN_THREADS = ROW_THREADS + 256;   // + warps 7-14: GroupNorm apply
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ROW_GN_THREADS, 1)
conv_row2_gn_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                    const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out, const RowParams p) {
    conv_row_body<true, true, true>(map_a0, map_a1, map_w, map_out, p);
}

}  // namespace sdc

using namespace sdc;

struct RowGn {
    const float* gamma; const float* beta; const float* ss; const int32_t* tindex; int64_t ss_stride; const void* residual;
};

// Returns SDC_OK when the problem was handled here, -1 when the shape is not eligible (caller uses conv_gemm).
static int conv3x3_row_launch(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                              const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout,
                              const RowGn* gn, void* stream) {
    SDC_REQUIRE(prec == SDC_PREC_TF32 || prec == SDC_PREC_F16, "conv3x3_row: precision %d", prec);
    const bool half = prec == SDC_PREC_F16;
    const int BK = half ? 64 : 32;
    if (W != RW || Cout > 128 || Cout % 32 != 0 || c0 % BK != 0 || c1 % BK != 0 || c0 <= 0) return -1;
    SDC_REQUIRE(a0 && w_packed && out && B > 0 && H > 0 && (c1 == 0 || a1), "conv3x3_row: bad arguments");
    RowParams p{};
    p.B = B; p.H = H; p.Cout = Cout; p.bn = Cout; p.c0 = c0; p.c1 = c1; p.operand_out = operand_out;
    p.bias = bias; p.residual = residual; p.out = out; p.stats = stats;
    if (gn) {
        p.gn_apply = 1; p.gn_gamma = gn->gamma; p.gn_beta = gn->beta; p.gn_ss = gn->ss; p.gn_tindex = gn->tindex;
        p.gn_ss_stride = gn->ss_stride; p.gn_residual = (const __half*)gn->residual;
    }
    { const char* e = getenv("SDC_ROW_DBG"); p.dbg = e ? atoi(e) : 0; }
    int n_sm = 148, dev = 0;
    SDC_CUDA(cudaGetDevice(&dev));
    SDC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    static const bool allow_pair = []() { const char* e = getenv("SDC_NO_2CTA"); return !(e && e[0] == '1'); }();
    const bool pair = allow_pair && Cout % 32 == 0 && (Cout / 2) % 8 == 0 && B * ((H + 3) / 4) >= n_sm / 2;
    p.pairs_per_image = pair ? (H + 3) / 4 : (H + 1) / 2;
    p.pairs_total = B * p.pairs_per_image;
    const int workers = pair ? n_sm / 2 : n_sm;
    const int ctas = p.pairs_total < workers ? p.pairs_total : workers;
    p.pairs_per_cta = (p.pairs_total + ctas - 1) / ctas;
    if (gn) {
        // the fused apply needs sample-aligned work ranges on the FP16 pair kernel, fp16 output, statistics, whole 4-row items
        if (!pair || !half || !operand_out || !stats || residual || H % 4 != 0 || Cout % 32 != 0 || 256 % (Cout / 8) != 0) return -1;
        p.pairs_per_cta = (p.pairs_per_cta + p.pairs_per_image - 1) / p.pairs_per_image * p.pairs_per_image;
    }
    const int grid = (p.pairs_total + p.pairs_per_cta - 1) / p.pairs_per_cta;

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
    const int smem_bytes = 2 * HALO_BYTES + ROW_BSTAGES * Cout * 128 + RSTG_BYTES + 25 * 8 + 16 + 1024;
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

extern "C" int sdc_conv3x3_row(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                               const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout,
                               void* stream) {
    return conv3x3_row_launch(prec, a0, c0, a1, c1, w_packed, bias, residual, out, stats, operand_out, B, H, W, Cout, nullptr, stream);
}

extern "C" int sdc_conv3x3_row_gn(const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias, void* out,
                                  double* stats, const float* gamma, const float* beta, const float* scale_shift,
                                  const int32_t* t_index, int64_t ss_stride, const void* gn_residual, int B, int H, int W, int Cout,
                                  void* stream) {
    SDC_REQUIRE(gamma && beta && stats, "conv3x3_row_gn: null GroupNorm arguments");
    RowGn gn{gamma, beta, scale_shift, t_index, ss_stride, gn_residual};
    return conv3x3_row_launch(SDC_PREC_F16, a0, c0, a1, c1, w_packed, bias, nullptr, out, stats, 1, B, H, W, Cout, &gn, stream);
}
