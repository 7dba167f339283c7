// Implicit-GEMM convolution on tcgen05 tensor cores fed by TMA.
//
// Replaces every nn.Conv2d of Unet2D except the 3-channel stem (/root/reference/1D/model/unet.py:132,161,
// 189-192,232-233,33-43,345,370,378):   out[M, Cout] = im2col(A)[M, K] * Wp[Cout, K]^T + bias
//   M = B*H*W output pixels (NHWC activations: fp16, or fp32 containers holding TF32-rounded values)
//   K = taps * Cin   (3x3 pad 1: 9 taps; 1x1: 1 tap; pixel-unshuffle 2x2/stride 2: 4 taps; nearest-upsample x2 + 3x3: four
//       2x2 phase convolutions of 4 taps on the LOW-resolution input -- the upsampled tensor never exists and the
//       contraction is 2.25x shorter, see kind 3 below),
//       Cin may be the concatenation of two tensors (U-Net skip connections) -> two K segments, no torch.cat.
//
// Mapping (persistent: one CTA per SM walks a contiguous range of 128 x BN output tiles, 192 threads):
//   warp 0   TMA producer: per K block (128 bytes of channels of one tap) one 4-D/5-D box load of the shifted activation
//            window (out-of-bounds rows/cols are zero-filled by TMA = the conv padding) + one 2-D load of the
//            weight slab, both SWIZZLE_128B, completing on an mbarrier;
//   warp 1   allocates TMEM, then one thread issues tcgen05.mma.kind::f16|tf32 (M=128|256, N=BN, 32 bytes of K) x4 per K block,
//            accumulating in TMEM, and releases smem stages with tcgen05.commit;
//   warps 2-9 epilogue (two per 32-lane TMEM quarter, alternating 32-column chunks: the per-chunk chain TMEM load -> registers
//            -> shared -> TMA store is latency bound, ncu showed 31% issue utilisation and 62% of HBM on the qkv projection with
//            four warps): tcgen05.ld the accumulator (lane = output row), add
//            bias / residual in registers, accumulate per-sample GroupNorm statistics (sum, sum of squares -> fp64
//            atomics), round to the operand precision if the output feeds another convolution, write the 32x32
//            chunk into a swizzled shared staging buffer and hand it to ONE TMA store (tc_ptx.cuh: epilogue_chunk).
//            A/B on B200 against a smem-transpose + st.global epilogue: 1x1 convolutions 1.7x faster (5 TB/s),
//            3x3 convolutions equal or faster (profiles/r01_epilogue_ab.txt).
//   The accumulator is double buffered in TMEM (2 x BN columns): the epilogue of tile i overlaps the MMAs of i+1.
// Precision: FP16 or TF32 operands (10-bit mantissa, rounded to nearest when produced; template HALF), FP32 accumulation.
#include "tc_ptx.cuh"
#include "../../include/safediffcon_b200_unet.h"
#include <math.h>
#include <stdlib.h>

namespace sdc {

// ------------------------------------------------------------------------------------------ kernel
constexpr int BM = 128;        // output pixels per tile (= UMMA M)
constexpr int A_BYTES = BM * 128;   // one K block of activations: 128 rows of 128 bytes (32 TF32 or 64 FP16 channels)
// Epilogue warps: EPI = 8 (two per TMEM lane quarter, alternating 32-column chunks) for the tensor-bound convolutions; EPI = 16
// (four per quarter) for the 1x1 projections with a short K loop, whose epilogue is the whole kernel: with two warps per scheduler
// the per-chunk chain (TMEM load -> ~250 instructions -> shared -> TMA store) kept the issue slots 40 % busy and the qkv projection
// at 0.64 of the copy peak (profiles/r02_ncu_full_B1024.summary.txt).
constexpr int STG_BUF = 4096;  // one epilogue staging buffer per epilogue warp: 32 rows x 128 bytes (TMA-store box)

struct GemmParams {
    int kind;            // 0: 1x1, 1: 3x3 pad 1, 2: 2x2 stride-2 (pixel-unshuffle + 1x1), 3: nearest-upsample x2 + 3x3 pad 1
                         //    (Upsample2d, unet.py:33-37).  Kind 3: output pixel (2i+a, 2j+b) only sees the 2x2 input window
                         //    rows {i+a-1, i+a}, cols {j+b-1, j+b}: per phase (a, b) a 4-tap convolution whose weights are sums
                         //    of the 3x3 taps that land on the same input pixel (packed by sdc_pack_conv_weight kind 3 as
                         //    Wp[phase*Cout + co, tap*Cin + ci]); zero padding of the upsampled image == out-of-range input rows.
                         //    M, H, W describe the INPUT (low) resolution; every phase writes a strided quarter of the output.
    int phases;          // 4 for kind 3, else 1; tile id = (m * phases + phase) * tiles_n + n
    int M;               // valid output rows
    int Cout;
    int bn;              // N tile (multiple of 32, <= 256)
    int H, W;            // OUTPUT spatial size
    int bh, bb;          // tile = bb images x bh rows x W cols
    int c0, c1;          // channels of segment 0 / 1 (c1 = 0: single input)
    int stages;
    int operand_out;     // 1: out feeds another tensor-core op -> TF32-rounded fp32 (TF32 mode) / fp16 (F16 mode); 0: plain fp32
    int hw_per_sample;   // H*W
    int tiles_n;         // Cout / bn
    int tiles_total;     // tiles_m * tiles_n, tile id = m * tiles_n + n (consecutive ids share the A window)
    int tiles_per_cta;
    int q_cols;          // > 0: LinearAttention qkv mode -- output columns [0, q_cols) get softmax_d * 32^-0.5 per 32-column head and
                         //      go to a second tensor (map_q, operand precision); columns [q_cols, Cout) go to `out` (fp32, or fp16
                         //      when operand_out is set in FP16 mode: k and v are then read at half the bytes by the context pass)
    int chunk_major;     // 3x3: K order (chunk, dx, dy) instead of (tap, chunk); see the producer loop
    int w_sample_rows;   // > 0: per-sample weights -- sample b uses weight rows [b * w_sample_rows, (b + 1) * w_sample_rows)
    // folded channel LayerNorm of the INPUT rows (PreNorm in front of the qkv projection, unet.py:65-76): the weights were packed as
    // W * g (sdc_pack_qkv_ln); the epilogue computes r_m * (acc - mu_m * wsum[col]) from per-row (mean, rstd)
    const float2* ln_rowstats;   // [M] or null
    const float* ln_wsum;        // [Cout]
    // channel LayerNorm of the OUTPUT rows + residual (LinearAttention.to_out = conv -> LayerNorm, then Residual, unet.py:190-193,
    // 16-22): out = LN(acc + bias) * g + residual, fp16; needs the whole row in one N tile (Cout == bn)
    const float* ln_out_gain;    // [Cout] or null
    const float* bias;       // [Cout] or null
    const void* residual;    // [M, Cout] in the operand precision, or null (added after bias)
    void* out;               // [M, Cout]
    double* stats;           // [B, 2] (sum, sumsq) accumulated with atomics, or null
};

// Persistent kernel: one CTA (PAIR = false) or one CTA pair (PAIR = true) per SM (pair) walks a contiguous range of
// output tiles.  The accumulator is double buffered in TMEM so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// PAIR (cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  Each CTA loads its own 128 activation rows but
// only HALF of the weight slab (BN/2 rows); tcgen05.mma.cta_group::2 (issued by the leader CTA) reads A and B from both
// CTAs' shared memory and writes each CTA's 128 accumulator rows into its own TMEM.  Per-SM operand traffic per K block
// drops from 16 KB + BN*128 B to 16 KB + BN*64 B, which is what bounds the 1-CTA kernel.  Barriers: TMA of both CTAs
// credits the leader's `full`; the MMA commit multicasts to both CTAs' `empty` / `acc_full`; both epilogues arrive on
// the leader's `acc_empty`.
template <bool HALF, bool PAIR, int EPI_WARPS>
__device__ __forceinline__ void conv_gemm_body(const CUtensorMap& map_a0, const CUtensorMap& map_a1, const CUtensorMap& map_w,
                                               const CUtensorMap& map_out, const CUtensorMap& map_q, const GemmParams& p) {
    using Op = Operand<HALF>;
    using act_t = typename ActT<HALF>::type;
    constexpr int BK = Op::kBK;
    constexpr int STG_BYTES = EPI_WARPS * STG_BUF;
    constexpr int COL_STEP = 32 * (EPI_WARPS / 4);   // column stride of one epilogue warp: the warps of a quarter interleave 32-column chunks
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int b_rows = PAIR ? p.bn / 2 : p.bn;            // weight rows staged by this CTA
    const int stage_bytes = A_BYTES + b_rows * 128;
    uint8_t* staging = smem + p.stages * stage_bytes;   // 1024-byte aligned (stage_bytes is a multiple of 1024)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes + STG_BYTES);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* acc_full = empty_bar + p.stages;   // [2]
    uint64_t* acc_empty = acc_full + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int taps = p.kind == 1 ? 9 : (p.kind >= 2 ? 4 : 1);
    const int ctot = p.c0 + p.c1;
    const int chunks = ctot / BK;
    const int tiles_pn = p.phases * p.tiles_n;
    const int num_kb = taps * chunks;
    uint32_t acc_cols = 32;                      // TMEM columns per accumulator buffer (power of two >= bn)
    while ((int)acc_cols < p.bn) acc_cols <<= 1;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_lo = worker * p.tiles_per_cta;            // PAIR: pair tiles, id = m2 * tiles_n + n
    const int tile_hi = min(p.tiles_total, tile_lo + p.tiles_per_cta);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a0);
        if (p.c1) tma_prefetch_desc(&map_a1);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_out);
        if (p.q_cols) tma_prefetch_desc(&map_q);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
        fence_barrier_init();
    }
    if (warp == 1) { if constexpr (PAIR) tmem_alloc_2sm(tmem_slot, 2 * acc_cols); else tmem_alloc(tmem_slot, 2 * acc_cols); }
    tc_fence_before();
    // PAIR: barriers of both CTAs initialised and TMEM allocated before any remote signalling
    if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;          // smem ring position across tiles
            uint32_t ph = 0;    // its phase parity
            // tile id = (mq * phases + phase) * tiles_n + nt, walked incrementally (no per-tile integer divisions)
            int mq = tile_lo / tiles_pn, phase = (tile_lo - mq * tiles_pn) / p.tiles_n, nt = tile_lo - mq * tiles_pn - phase * p.tiles_n;
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const int mt = PAIR ? 2 * mq + (int)rank : mq;
                // tile origin in (image, row); tiles always span full rows (bw == W)
                const int pix0 = mt * BM;
                const int b0 = pix0 / p.hw_per_sample;
                const int h0 = (pix0 - b0 * p.hw_per_sample) / p.W;
                // ring position (s, ph), tap and channel chunk advance incrementally: this single thread's instruction stream is
                // the critical path of the short-K 1x1 convolutions (2 K blocks per tile), integer divisions do not belong in it
                // K order: kinds 0, 2, 3 walk the channel chunks of a tap, then the next tap.  3x3 (kind 1): chunk-major with the taps
                // of a chunk in the order (dx, dy) -- the order of conv_row64.cu, so that a sample's result does not depend on which of
                // the two kernels the batch size selects (same fp32 accumulation sequence; tests/test_full_size_gpu.py)
                for (int kb = 0, tap = 0, ck = 0, t_dx = 0, t_dy = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    uint8_t* sa = smem + s * stage_bytes;
                    uint8_t* sb = sa + A_BYTES;
                    if constexpr (PAIR) { if (leader) mbar_expect_tx(&full_bar[s], (uint32_t)(2 * stage_bytes)); }   // bytes of BOTH CTAs
                    else mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
                    const int cc = ck * BK;          // channel offset in the concatenated input
                    const bool second = cc >= p.c0;
                    const CUtensorMap* ma = second ? &map_a1 : &map_a0;
                    const int cseg = second ? cc - p.c0 : cc;
                    if (p.kind == 2) {
                        // input viewed as [B, H, 2(p1), W, 2*C]: coordinate (p2*C + c, w, p1, h, b)
                        const int cin = second ? p.c1 : p.c0;
                        if constexpr (PAIR) tma_load_5d_2sm(sa, ma, &full_bar[s], (tap & 1) * cin + cseg, 0, tap >> 1, h0, b0);
                        else tma_load_5d(sa, ma, &full_bar[s], (tap & 1) * cin + cseg, 0, tap >> 1, h0, b0);
                    } else {
                        int dy = 0, dx = 0;
                        if (p.kind == 1) { const int ty = (tap * 11) >> 5; dy = ty - 1; dx = tap - 3 * ty - 1; }   // tap / 3 for tap < 9
                        else if (p.kind == 3) { dy = (phase >> 1) - 1 + (tap >> 1); dx = (phase & 1) - 1 + (tap & 1); }
                        if constexpr (PAIR) tma_load_4d_2sm(sa, ma, &full_bar[s], cseg, dx, h0 + dy, b0);
                        else tma_load_4d(sa, ma, &full_bar[s], cseg, dx, h0 + dy, b0);
                    }
                    // per-sample weights: tiles never straddle samples (host check); kind 3: one weight block per phase
                    const int wrow = nt * p.bn + b0 * p.w_sample_rows + phase * p.Cout;
                    if constexpr (PAIR) tma_load_2d_2sm(sb, &map_w, &full_bar[s], tap * ctot + cc, wrow + (int)rank * b_rows);
                    else tma_load_2d(sb, &map_w, &full_bar[s], tap * ctot + cc, wrow);
                    if (p.chunk_major) {
                        if (++t_dy == 3) { t_dy = 0; if (++t_dx == 3) { t_dx = 0; ++ck; } }
                        tap = t_dy * 3 + t_dx;
                    } else if (++ck == chunks) { ck = 0; ++tap; }
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
                if (++nt == p.tiles_n) { nt = 0; if (++phase == p.phases) { phase = 0; ++mq; } }
            }
        }
    } else if (warp == 1) {
        if (leader) {   // whole warp walks the loops (uniform control flow); one elected lane issues the MMAs and commits
            const uint32_t idesc = Op::idesc(p.bn, PAIR ? 2 * BM : BM);
            int s = 0, it = 0;
            uint32_t ph = 0;
            for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
                const int buf = it & 1;
                mbar_wait(&acc_empty[buf], ((uint32_t)(it >> 1) & 1u) ^ 1u);   // epilogue(s) drained this buffer
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * acc_cols;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * stage_bytes);
                    const uint64_t adesc = make_sw128_desc(sa);
                    const uint64_t bdesc = make_sw128_desc(sa + A_BYTES);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            // advance 32 bytes (8 TF32 / 16 FP16) along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
                            Op::template mma<PAIR>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
                        }
                        if constexpr (PAIR) umma_commit_2sm(&empty_bar[s]); else umma_commit(&empty_bar[s]);
                        if (kb == num_kb - 1) { if constexpr (PAIR) umma_commit_2sm(&acc_full[buf]); else umma_commit(&acc_full[buf]); }
                    }
                    __syncwarp();
                    if (++s == p.stages) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ---- epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); lane = accumulator row ----
        const int q = warp & 3, half_id = (warp - 2) >> 2;   // TMEM lane quarter; which of the EPI_WARPS / 4 interleaved column-chunk sets
        const uint32_t stg = smem_u32(staging + (warp - 2) * STG_BUF);
        const act_t* resid = reinterpret_cast<const act_t*>(p.residual);
        const bool out_half = HALF && p.operand_out;
        int it = 0;
        int mq = tile_lo / tiles_pn, phase = (tile_lo - mq * tiles_pn) / p.tiles_n, nt = tile_lo - mq * tiles_pn - phase * p.tiles_n;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int up = p.kind == 3 ? phase : -1;
            const int mt = PAIR ? 2 * mq + (int)rank : mq;
            const int buf = it & 1;
            mbar_wait(&acc_full[buf], (uint32_t)(it >> 1) & 1u);
            tc_fence_after();
            const int m_w = mt * BM + q * 32;          // first row of this warp
            const int m = m_w + lane;
            const bool row_ok = m < p.M;
            float s1 = 0.f, s2 = 0.f;
            float ln_mu = 0.f, ln_r = 1.f;
            if (p.ln_rowstats && row_ok) { const float2 rs = __ldg(p.ln_rowstats + m); ln_mu = rs.x; ln_r = rs.y; }
            if (p.ln_out_gain) {
                if (m_w < p.M) {
                    // ---- output LayerNorm: pass 1 over ALL columns of this lane's row (both warps of the quarter, redundantly) ----
                    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * acc_cols;
                    float a1 = 0.f, a2 = 0.f;
                    for (int c = 0; c < p.bn; c += 32) {
                        uint32_t r[32];
                        tmem_ld32(trow + (uint32_t)c, r);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float x0 = __uint_as_float(r[j]) + bb.x, x1 = __uint_as_float(r[j + 1]) + bb.y;
                            const float x2 = __uint_as_float(r[j + 2]) + bb.z, x3 = __uint_as_float(r[j + 3]) + bb.w;
                            a1 += (x0 + x1) + (x2 + x3);
                            a2 = fmaf(x0, x0, a2); a2 = fmaf(x1, x1, a2); a2 = fmaf(x2, x2, a2); a2 = fmaf(x3, x3, a2);
                        }
                    }
                    const float inv_c = 1.0f / (float)p.bn;
                    const float mean = a1 * inv_c;
                    const float rstd = rsqrtf(fmaxf(a2 * inv_c - mean * mean, 0.f) + 1e-5f);
                    // ---- pass 2: this warp's column chunks: normalise, gain, residual, fp16 store ----
                    for (int c = 32 * half_id; c < p.bn; c += COL_STEP) {
                        uint32_t r[32];
                        tmem_ld32(trow + (uint32_t)c, r);
                        float v[32];
                        const act_t* rrow = resid + (size_t)m * p.Cout + c;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + c + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float4 gg = __ldg(reinterpret_cast<const float4*>(p.ln_out_gain + c + j));
                            const float4 rr = row_ok ? load4_nc(rrow + j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            v[j] = fmaf((__uint_as_float(r[j]) + bb.x - mean) * rstd, gg.x, rr.x);
                            v[j + 1] = fmaf((__uint_as_float(r[j + 1]) + bb.y - mean) * rstd, gg.y, rr.y);
                            v[j + 2] = fmaf((__uint_as_float(r[j + 2]) + bb.z - mean) * rstd, gg.z, rr.z);
                            v[j + 3] = fmaf((__uint_as_float(r[j + 3]) + bb.w - mean) * rstd, gg.w, rr.w);
                        }
                        if (lane == 0) bulk_wait_read<0>();
                        __syncwarp();
                        stage_store_half(v, stg, &map_out, c, m_w, lane);
                    }
                }
            } else
            for (int c = 32 * half_id; c < p.bn; c += COL_STEP) {
                const int col = nt * p.bn + c;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * acc_cols + (uint32_t)c;
                const act_t* rrow = resid ? resid + (size_t)m * p.Cout + col : nullptr;
                const float* lw = p.ln_rowstats ? p.ln_wsum + col : nullptr;
                if (m_w < p.M) {
                    if (col < p.q_cols) epilogue_chunk_qsoftmax<HALF>(taddr, stg, &map_q, col, m_w, lane, true, lw, ln_mu, ln_r);
                    else if (out_half) epilogue_chunk<true, act_t>(taddr, stg, &map_out, col - p.q_cols, m_w, row_ok, p.bias, rrow, false, s1, s2, lane, up, p.W, true, p.stats != nullptr, lw, ln_mu, ln_r);
                    else epilogue_chunk<false, act_t>(taddr, stg, &map_out, col - p.q_cols, m_w, row_ok, p.bias, rrow, p.operand_out != 0, s1, s2, lane, up, p.W, true, p.stats != nullptr, lw, ln_mu, ln_r);
                }
            }
            // accumulator buffer fully read -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (PAIR) mbar_arrive_leader(&acc_empty[buf]); else mbar_arrive(&acc_empty[buf]); }
            if (p.stats) {
                // all 32 rows of a warp belong to one sample (H*W is a multiple of 32)
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                if (lane == 0 && m_w < p.M) {
                    const int b = m_w / p.hw_per_sample;
                    atomicAdd(p.stats + 2 * b, (double)s1);
                    atomicAdd(p.stats + 2 * b + 1, (double)s2);
                }
            }
            if (++nt == p.tiles_n) { nt = 0; if (++phase == p.phases) { phase = 0; ++mq; } }
        }
        if (lane == 0) bulk_wait<0>();   // all output stores complete before the CTA's shared memory goes away
        tc_fence_before();
    }
    // PAIR: neither CTA may exit (or free TMEM) while its partner can still read its shared memory / signal its barriers
    __syncthreads();
    if constexpr (PAIR) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, 2 * acc_cols); else tmem_dealloc(tmem_base, 2 * acc_cols);
    }
}

template <bool HALF, int EPI_WARPS = 8>
__global__ void __launch_bounds__(64 + 32 * EPI_WARPS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                 const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                 const __grid_constant__ CUtensorMap map_q, const GemmParams p) {
    conv_gemm_body<HALF, false, EPI_WARPS>(map_a0, map_a1, map_w, map_out, map_q, p);
}
template <bool HALF, int EPI_WARPS = 8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EPI_WARPS, 1)
conv_gemm2_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_out,
                  const __grid_constant__ CUtensorMap map_q, const GemmParams p) {
    conv_gemm_body<HALF, true, EPI_WARPS>(map_a0, map_a1, map_w, map_out, map_q, p);
}

// ------------------------------------------------------------------------------------------ host side
// activation map: NHWC [B, Hin, Win, C]; for kind 2 the 5-D pixel-unshuffle view
static int encode_act(CUtensorMap* map, const void* a, int kind, int B, int H, int W, int C, int bh, int bb, bool half) {
    const cuuint64_t eb = half ? 2 : 4;
    const cuuint32_t bk = half ? 64 : 32;
    if (kind == 2) {
        // input spatial 2H x 2W;  dims (inner->outer): {2C, W, 2, H, B}
        cuuint64_t dims[5] = {(cuuint64_t)2 * C, (cuuint64_t)W, 2, (cuuint64_t)H, (cuuint64_t)B};
        cuuint64_t str[4] = {(cuuint64_t)2 * C * eb, (cuuint64_t)2 * W * C * eb, (cuuint64_t)4 * W * C * eb,
                             (cuuint64_t)4 * H * W * C * eb};
        cuuint32_t box[5] = {bk, (cuuint32_t)W, 1, (cuuint32_t)bh, (cuuint32_t)bb};
        return encode_tmap(map, a, 5, dims, str, box, half);
    }
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)C * eb, (cuuint64_t)W * C * eb, (cuuint64_t)H * W * C * eb};
    cuuint32_t box[4] = {bk, (cuuint32_t)W, (cuuint32_t)bh, (cuuint32_t)bb};
    return encode_tmap(map, a, 4, dims, str, box, half);
}

}  // namespace sdc

using namespace sdc;

struct GemmLn { const float* rowstats = nullptr; const float* wsum = nullptr; const float* out_gain = nullptr; };

static int conv_gemm_launch(int prec, int kind, const void* a0, int c0, const void* a1, int c1, const void* w_packed,
                            const float* bias, const void* residual, void* out, double* stats, int operand_out, int B, int H,
                            int W, int Cout, void* q_out, int q_cols, int per_sample_weights, void* stream, const GemmLn* ln = nullptr) {
    SDC_REQUIRE(prec == SDC_PREC_TF32 || prec == SDC_PREC_F16, "conv_gemm: precision %d", prec);
    const bool half = prec == SDC_PREC_F16;
    const int BK = half ? 64 : 32;
    SDC_REQUIRE(kind >= 0 && kind <= 3, "conv_gemm: kind %d", kind);
    SDC_REQUIRE(kind != 3 || (c1 == 0 && !residual && !stats && !q_cols && !per_sample_weights && (W == 16 || W % 32 == 0)),
                "conv_gemm: the fused upsample convolution takes one input, no residual / statistics, input W of 16 or a multiple of 32");
    SDC_REQUIRE(B > 0 && H > 0 && W > 0, "conv_gemm: empty problem");
    SDC_REQUIRE(a0 && w_packed && out, "conv_gemm: null pointer");
    SDC_REQUIRE(c0 > 0 && c0 % BK == 0 && c1 >= 0 && c1 % BK == 0 && (c1 == 0 || a1), "conv_gemm: channels must be multiples of %d", BK);
    SDC_REQUIRE(Cout % 32 == 0, "conv_gemm: Cout=%d must be a multiple of 32", Cout);
    SDC_REQUIRE(W <= BM && BM % W == 0, "conv_gemm: W=%d must divide %d", W, BM);
    int bh = BM / W;
    if (bh > H) bh = H;
    SDC_REQUIRE(H % bh == 0, "conv_gemm: H=%d not tileable by %d rows", H, bh);
    const int bb = BM / (W * bh);
    SDC_REQUIRE(bb * bh * W == BM && (H * W) % 32 == 0, "conv_gemm: H*W=%d cannot be tiled into 128-pixel blocks", H * W);
    SDC_REQUIRE(bb == 1 || bh == H, "conv_gemm: tile spans images only when it holds whole images");
    int bn = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : (Cout % 64 == 0 ? 64 : 32));
    SDC_REQUIRE(q_cols >= 0 && q_cols % 32 == 0 && q_cols < Cout && (q_cols == 0 || (q_out && !bias && !residual && !stats && (!operand_out || half))),
                "conv_gemm: qkv mode needs q_out, q_cols %% 32 == 0 and a kv output that is plain fp32 or (FP16 mode) an operand");
    GemmParams p{};
    p.q_cols = q_cols;
    static const bool chunk_major = []() { const char* e = getenv("SDC_KORDER"); return !(e && e[0] == '0'); }();
    p.chunk_major = kind == 1 && chunk_major;
    p.w_sample_rows = per_sample_weights ? Cout : 0;
    p.phases = kind == 3 ? 4 : 1;
    p.kind = kind; p.M = B * H * W; p.Cout = Cout; p.bn = bn; p.H = H; p.W = W; p.bh = bh; p.bb = bb;
    p.c0 = c0; p.c1 = c1; p.operand_out = operand_out; p.hw_per_sample = H * W;
    p.bias = bias; p.residual = residual; p.out = out; p.stats = stats;
    if (ln) {
        p.ln_rowstats = (const float2*)ln->rowstats; p.ln_wsum = ln->wsum; p.ln_out_gain = ln->out_gain;
        SDC_REQUIRE(!ln->rowstats || ln->wsum, "conv_gemm: folded input LayerNorm needs the weight row sums");
        SDC_REQUIRE(!ln->out_gain || (half && operand_out && residual && Cout == bn && !stats && !q_cols && kind == 0),
                    "conv_gemm: the fused output LayerNorm needs FP16 mode, an fp16 output, a residual and Cout (%d) in one N tile (128 or 256)", Cout);
    }
    int n_sm = 148;
    {
        int dev = 0;
        SDC_CUDA(cudaGetDevice(&dev));
        SDC_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    static const bool allow_pair = []() { const char* e = getenv("SDC_NO_2CTA"); return !(e && e[0] == '1'); }();
    const int tiles_m = (p.M + BM - 1) / BM;
    // CTA pairs (cta_group::2) whenever there are at least as many 256-row pair tiles as SM pairs
    // per-sample weights: every (pair) tile must lie inside one sample
    SDC_REQUIRE(!per_sample_weights || (H * W) % BM == 0, "conv_gemm: per-sample weights need H*W %% 128 == 0");
    // ... or, with fewer tiles than that, when the contraction is long (>= 32 K blocks) and the tile count even enough: such launches
    // (the 2x16 / 4x32 levels at the reference's batch sizes) are bound by each SM re-reading the weight slab from L2, which a pair halves
    static const int pair_long_k = []() { const char* e = getenv("SDC_PAIR_LONGK"); return e ? atoi(e) : 32; }();
    const int taps_ = kind == 1 ? 9 : (kind >= 2 ? 4 : 1);
    const bool long_k = pair_long_k > 0 && taps_ * ((c0 + c1) / BK) >= pair_long_k && tiles_m >= 2 && tiles_m * (Cout / bn) * p.phases <= n_sm;
    const bool pair = allow_pair && (((tiles_m + 1) / 2) * (Cout / bn) * p.phases >= n_sm / 2 || long_k) && (!per_sample_weights || (H * W) % (2 * BM) == 0);
    const int stage_bytes = A_BYTES + (pair ? bn / 2 : bn) * 128;
    // 16 epilogue warps when the epilogue is the kernel: 1x1 convolutions with at most 4 K blocks (FP16 mode; SDC_EPI16=0 restores 8)
    static const bool allow_epi16 = []() { const char* e = getenv("SDC_EPI16"); return !(e && e[0] == '0'); }();
    // (measured: +10-15 % on the qkv and per-sample projections with K = 128; the plain 1x1 convolutions already stream at 0.9+ of the
    // copy peak with 8 warps and lose 10 % to the smaller register budget)
    const bool epi16 = allow_epi16 && half && kind == 0 && (c0 + c1) / BK <= 2 && (q_cols > 0 || per_sample_weights) && !(ln && ln->out_gain);
    const int epi_warps = epi16 ? 16 : 8;
    const int stg_bytes = epi_warps * STG_BUF;
    int stages = (222 * 1024 - stg_bytes) / stage_bytes;
    if (stages > 8) stages = 8;
    if (epi16 && stages > 6) stages = 6;   // short K loops: a deeper ring holds nothing
    p.stages = stages;
    const int smem_bytes = stages * stage_bytes + stg_bytes + (2 * stages + 4) * 8 + 16 + 1024;
    p.tiles_n = Cout / bn;
    p.tiles_total = (pair ? (tiles_m + 1) / 2 : tiles_m) * p.tiles_n * p.phases;
    const int workers = pair ? n_sm / 2 : n_sm;
    const int ctas = p.tiles_total < workers ? p.tiles_total : workers;
    p.tiles_per_cta = (p.tiles_total + ctas - 1) / ctas;

    SDC_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "conv_gemm: out must be 16-byte aligned (TMA store)");
    CUtensorMap ma0, ma1, mw;
    int rc = encode_act(&ma0, a0, kind, B, H, W, c0, bh, bb, half);
    if (rc) return rc;
    if (c1) { rc = encode_act(&ma1, a1, kind, B, H, W, c1, bh, bb, half); if (rc) return rc; } else ma1 = ma0;
    const int taps = kind == 1 ? 9 : (kind >= 2 ? 4 : 1);
    const cuuint64_t ktot = (cuuint64_t)taps * (c0 + c1);
    cuuint64_t wd[2] = {ktot, (cuuint64_t)Cout * (cuuint64_t)(per_sample_weights ? B : p.phases)};
    cuuint64_t ws[1] = {ktot * (half ? 2 : 4)};
    cuuint32_t wb[2] = {(cuuint32_t)BK, (cuuint32_t)(pair ? bn / 2 : bn)};
    rc = encode_tmap(&mw, w_packed, 2, wd, ws, wb, half);
    if (rc) return rc;
    CUtensorMap mo, mq;
    if (kind == 3) {
        // high-resolution output [B, 2H, 2W, Cout] viewed as {Cout, 2 (b), W, 2 (a), B*H}; a 32-row chunk = 32 consecutive
        // low-resolution pixels = box {32 cols, 1, min(W, 32), 1, 32 / min(W, 32)}
        const bool oh = half && operand_out;
        const cuuint64_t eb = oh ? 2 : 4;
        const cuuint32_t bw = W < 32 ? W : 32;
        cuuint64_t od[5] = {(cuuint64_t)Cout, 2, (cuuint64_t)W, 2, (cuuint64_t)B * H};
        cuuint64_t os[4] = {(cuuint64_t)Cout * eb, (cuuint64_t)2 * Cout * eb, (cuuint64_t)2 * W * Cout * eb, (cuuint64_t)4 * W * Cout * eb};
        cuuint32_t ob[5] = {32, 1, bw, 1, 32 / bw};
        rc = encode_tmap(&mo, out, 5, od, os, ob, oh, oh ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
    } else {
        rc = encode_out_tmap(&mo, out, p.M, Cout - q_cols, half && operand_out);
    }
    if (rc) return rc;
    if (q_cols) { rc = encode_out_tmap(&mq, q_out, p.M, q_cols, half); if (rc) return rc; } else mq = mo;

    static bool attr_set = false;
    if (!attr_set) {
        SDC_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute(conv_gemm2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute((conv_gemm_kernel<true, 16>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SDC_CUDA(cudaFuncSetAttribute((conv_gemm2_kernel<true, 16>), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = (p.tiles_total + p.tiles_per_cta - 1) / p.tiles_per_cta;
    cudaStream_t st = as_stream(stream);
    const int threads = 64 + 32 * epi_warps;
    if (pair) {
        if (epi16) conv_gemm2_kernel<true, 16><<<2 * grid, threads, smem_bytes, st>>>(ma0, ma1, mw, mo, mq, p);
        else if (half) conv_gemm2_kernel<true><<<2 * grid, threads, smem_bytes, st>>>(ma0, ma1, mw, mo, mq, p);
        else conv_gemm2_kernel<false><<<2 * grid, threads, smem_bytes, st>>>(ma0, ma1, mw, mo, mq, p);
    } else {
        if (epi16) conv_gemm_kernel<true, 16><<<grid, threads, smem_bytes, st>>>(ma0, ma1, mw, mo, mq, p);
        else if (half) conv_gemm_kernel<true><<<grid, threads, smem_bytes, st>>>(ma0, ma1, mw, mo, mq, p);
        else conv_gemm_kernel<false><<<grid, threads, smem_bytes, st>>>(ma0, ma1, mw, mo, mq, p);
    }
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_conv_gemm(int prec, int kind, const void* a0, int c0, const void* a1, int c1, const void* w_packed,
                             const float* bias, const void* residual, void* out, double* stats, int operand_out, int B, int H,
                             int W, int Cout, void* stream) {
    return conv_gemm_launch(prec, kind, a0, c0, a1, c1, w_packed, bias, residual, out, stats, operand_out, B, H, W, Cout, nullptr, 0, 0,
                            stream);
}

extern "C" int sdc_conv1x1_qkv(int prec, const void* a, int c, const void* w_packed, void* q_out, void* kv_out, int kv_operand, int B,
                               int H, int W, int hidden, void* stream) {
    SDC_REQUIRE(hidden > 0 && hidden % 32 == 0 && q_out && kv_out, "conv1x1_qkv: bad arguments");
    SDC_REQUIRE(!kv_operand || prec == SDC_PREC_F16, "conv1x1_qkv: an operand-precision kv tensor exists only in FP16 mode");
    return conv_gemm_launch(prec, 0, a, c, nullptr, 0, w_packed, nullptr, nullptr, kv_out, nullptr, kv_operand ? 1 : 0, B, H, W, 3 * hidden,
                            q_out, hidden, 0, stream);
}

extern "C" int sdc_conv1x1_qkv_ln(const void* a, int c, const void* w_folded, const float* wsum, const float* rowstats, void* q_out, void* kv_out,
                                  int B, int H, int W, int hidden, void* stream) {
    SDC_REQUIRE(hidden > 0 && hidden % 32 == 0 && q_out && kv_out && wsum && rowstats, "conv1x1_qkv_ln: bad arguments");
    GemmLn ln;
    ln.rowstats = rowstats; ln.wsum = wsum;
    return conv_gemm_launch(SDC_PREC_F16, 0, a, c, nullptr, 0, w_folded, nullptr, nullptr, kv_out, nullptr, 1, B, H, W, 3 * hidden, q_out, hidden, 0,
                            stream, &ln);
}

extern "C" int sdc_conv1x1_per_sample_ln(const void* a, int c, const void* w_per_sample, const float* bias, const float* gain, const void* residual,
                                         void* out, int B, int H, int W, int Cout, void* stream) {
    SDC_REQUIRE(gain && residual, "conv1x1_per_sample_ln: null gain / residual");
    GemmLn ln;
    ln.out_gain = gain;
    return conv_gemm_launch(SDC_PREC_F16, 0, a, c, nullptr, 0, w_per_sample, bias, residual, out, nullptr, 1, B, H, W, Cout, nullptr, 0, 1, stream, &ln);
}

extern "C" int sdc_conv1x1_per_sample(int prec, const void* a, int c, const void* w_per_sample, const float* bias, void* out,
                                      int operand_out, int B, int H, int W, int Cout, void* stream) {
    return conv_gemm_launch(prec, 0, a, c, nullptr, 0, w_per_sample, bias, nullptr, out, nullptr, operand_out, B, H, W, Cout, nullptr, 0, 1,
                            stream);
}
