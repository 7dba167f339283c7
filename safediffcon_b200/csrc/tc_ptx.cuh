// tcgen05 / TMA / mbarrier PTX wrappers and tensor-map helpers shared by the tensor-core kernels (sm_100a).
// Bit layouts follow the sm_100 UMMA shared-memory and instruction descriptors.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace sdc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a fully converged warp (elect.sync): keeps the surrounding control flow warp-uniform so that ptxas holds
// descriptors / addresses in uniform registers instead of wrapping every UTCHMMA / UTMALDG in an R2UR election loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(addr), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// FP16 operands (10-bit mantissa like TF32, but 2 bytes: twice the MMA rate and half the operand traffic), FP32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives on the mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart).
// Bit layout per the sm_100 UMMA descriptor: [0,14) start>>4, [16,30) LBO>>4 (unused for swizzled K-major, 1),
// [32,46) SBO>>4 = 64, [46,48) version = 1, [61,64) layout = 2 (SWIZZLE_128B).
// The start address may be any multiple of 128 bytes inside a 1024-byte-aligned, TMA-written SWIZZLE_128B region
// (a row-shifted view): measured on B200, the hardware derives the swizzle phase from the absolute shared-memory
// address, so the base-offset field [49,52) stays 0 (setting it to (start >> 7) & 7 produces wrong operands).
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = 64u | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// explicit shared-window accesses for the epilogue staging tile: the dynamic-smem base is re-aligned through an integer
// cast, after which the compiler no longer knows the address space and would emit generic LD.E / ST.E (long-scoreboard)
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
// TMA store of a shared-memory box to global memory (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(map), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------------------------------------- epilogue chunk
// One 32-row x 32-column accumulator chunk of an epilogue warp (lane = row):
//   tcgen05.ld -> registers (bias, residual, GroupNorm sums, rounding) -> swizzled shared staging -> ONE TMA store.
// The thread-level work is ~40 instructions per chunk (the earlier smem-transpose + per-row st.global epilogue was ~390
// and, with one epilogue warp per scheduler, issue-latency bound: 1x1 convolutions ran at 40% of HBM).
//   OUT_HALF = false: fp32 rows of 128 bytes, SWIZZLE_128B box {32, 32}; `round` stores TF32-rounded values.
//   OUT_HALF = true : fp16 rows of 64 bytes,  SWIZZLE_64B  box {32, 32}.
// stg: 1024-byte aligned shared address of a 4 KB staging buffer that no in-flight bulk store is still reading.
// res_row: this lane's residual row at column `col` (operand precision) or null.  Rows beyond the tensor are clipped by TMA;
// `row_ok` keeps them out of the statistics.
template <bool OUT_HALF, typename res_t>
__device__ __forceinline__ void epilogue_chunk(uint32_t taddr, uint32_t stg, const CUtensorMap* map_out, int col, int row0,
                                               bool row_ok, const float* __restrict__ bias, const res_t* __restrict__ res_row,
                                               bool round, float& s1, float& s2, int lane, int up_phase = -1, int up_w = 0,
                                               bool wait_stg = false, bool want_stats = true, const float* __restrict__ ln_wsum = nullptr,
                                               float ln_mu = 0.f, float ln_r = 1.f) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (ln_wsum) {
        // folded channel LayerNorm of the INPUT row: W (x - mu) r = r (W x - mu * sum_c W[., c]); ln_wsum points at this chunk's columns
        const float nmr = -ln_mu * ln_r;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(ln_wsum + j));
            v[j] = fmaf(nmr, w.x, ln_r * v[j]); v[j + 1] = fmaf(nmr, w.y, ln_r * v[j + 1]);
            v[j + 2] = fmaf(nmr, w.z, ln_r * v[j + 2]); v[j + 3] = fmaf(nmr, w.w, ln_r * v[j + 3]);
        }
    }
    if (bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
    }
    if (res_row && row_ok) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 b = load4_nc(res_row + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
    }
    if (row_ok && want_stats) {
        float a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { a1 += v[j]; a2 = fmaf(v[j], v[j], a2); }
        s1 += a1;
        s2 += a2;
    }
    if (wait_stg) {
        // single staging buffer per warp: the bulk store of this warp's previous chunk must have finished READING it (the TMEM
        // load and the arithmetic above already overlapped that read)
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    if constexpr (OUT_HALF) {
        const uint32_t rowa = stg + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __half2 h = __floats2half2_rn(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
                w[k] = *reinterpret_cast<const uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((uint32_t)j ^ sw) << 4)), "r"(w[0]), "r"(w[1]),
                         "r"(w[2]), "r"(w[3]) : "memory");
        }
    } else {
        const uint32_t rowa = stg + (uint32_t)lane * 128u, sw = (uint32_t)lane & 7u;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            if (round) o = make_float4(to_tf32(o.x), to_tf32(o.y), to_tf32(o.z), to_tf32(o.w));
            sts128(rowa + (((uint32_t)j ^ sw) << 4), o);
        }
    }
    fence_proxy_async();   // generic-proxy writes above -> visible to the async proxy (TMA) below
    __syncwarp();
    if (lane == 0) {
        // up_phase >= 0: rows are low-resolution pixels of one phase (a, b) of a fused nearest-upsample convolution; the output
        // map is the 5-D view {Cout, 2 (b), W, 2 (a), B*H} of the high-resolution tensor (see conv_gemm.cu, kind 3)
        if (up_phase >= 0) tma_store_5d(map_out, stg, col, up_phase & 1, row0 % up_w, up_phase >> 1, row0 / up_w);
        else tma_store_2d(map_out, stg, col, row0);
        bulk_commit();
    }
}

// Tail of the fp16 epilogue on its own: 32 fp32 values of this lane's row -> fp16 -> SWIZZLE_64B staging tile -> ONE TMA store
// of the 32 x 32 chunk at (col, row0).  The caller guarantees that no in-flight bulk store still reads `stg`.
__device__ __forceinline__ void stage_store_half(const float (&v)[32], uint32_t stg, const CUtensorMap* map_out, int col, int row0,
                                                 int lane) {
    const uint32_t rowa = stg + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const __half2 h = __floats2half2_rn(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
            w[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((uint32_t)j ^ sw) << 4)), "r"(w[0]), "r"(w[1]),
                     "r"(w[2]), "r"(w[3]) : "memory");
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map_out, stg, col, row0);
        bulk_commit();
    }
}

// Variant for the q columns of a LinearAttention qkv projection: the 32 accumulator columns of a chunk are exactly one head's
// logits of this lane's pixel, so q <- softmax_d(q) * 32^-0.5 (/root/reference/1D/model/unet.py:206,209) is applied in
// registers and stored in the operand precision (it is the A operand of the folded output projection).
template <bool OUT_HALF>
__device__ __forceinline__ void epilogue_chunk_qsoftmax(uint32_t taddr, uint32_t stg, const CUtensorMap* map_q, int col, int row0,
                                                        int lane, bool wait_stg = false, const float* __restrict__ ln_wsum = nullptr,
                                                        float ln_mu = 0.f, float ln_r = 1.f) {
    uint32_t r[32];
    tmem_ld32(taddr, r);
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    if (ln_wsum) {   // folded LayerNorm of the input row (see epilogue_chunk)
        const float nmr = -ln_mu * ln_r;
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(ln_wsum + j));
            v[j] = fmaf(nmr, w.x, ln_r * v[j]); v[j + 1] = fmaf(nmr, w.y, ln_r * v[j + 1]);
            v[j + 2] = fmaf(nmr, w.z, ln_r * v[j + 2]); v[j + 3] = fmaf(nmr, w.w, ln_r * v[j + 3]);
        }
    }
    float mx = v[0];
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
    // exp(v - mx) = ex2(v * log2e - mx * log2e): one packed FFMA2 + two MUFU per element pair, packed adds / scaling (sm_100 f32x2
    // arithmetic halves the FP32 issue slots of this chunk, which is issue bound: 4 epilogue warps per scheduler share them).
    // ex2.approx (2 ulp) on arguments <= 0: far inside the 2^-11 rounding of the fp16 / tf32 store below
    const float l2e = 1.4426950408889634f;
    const float2 k2 = make_float2(l2e, l2e), nm2 = make_float2(-mx * l2e, -mx * l2e);
    float2 den2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float2 a = __ffma2_rn(make_float2(v[j], v[j + 1]), k2, nm2);
        float2 e;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(a.x));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(a.y));
        den2 = __fadd2_rn(den2, e);
        v[j] = e.x; v[j + 1] = e.y;
    }
    const float sc = 0.17677669529663687f / (den2.x + den2.y);
    const float2 sc2 = make_float2(sc, sc);
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        const float2 o = __fmul2_rn(make_float2(v[j], v[j + 1]), sc2);
        v[j] = o.x; v[j + 1] = o.y;
    }
    if (wait_stg) {
        if (lane == 0) bulk_wait_read<0>();
        __syncwarp();
    }
    if constexpr (OUT_HALF) {
        const uint32_t rowa = stg + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __half2 h = __floats2half2_rn(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
                w[k] = *reinterpret_cast<const uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + (((uint32_t)j ^ sw) << 4)), "r"(w[0]), "r"(w[1]),
                         "r"(w[2]), "r"(w[3]) : "memory");
        }
    } else {
        const uint32_t rowa = stg + (uint32_t)lane * 128u, sw = (uint32_t)lane & 7u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            sts128(rowa + (((uint32_t)j ^ sw) << 4),
                   make_float4(to_tf32(v[4 * j]), to_tf32(v[4 * j + 1]), to_tf32(v[4 * j + 2]), to_tf32(v[4 * j + 3])));
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
        tma_store_2d(map_q, stg, col, row0);
        bulk_commit();
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ---------------------------------------------------------------------------------------------- CTA-pair (cta_group::2)
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA loads issued by either CTA of a pair; the transaction bytes are credited to the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                                int c4) {
    asm volatile("cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem, 256 x N over the CTA pair] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by the leader only
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Operand precision of the tensor-core kernels.  HALF = false: TF32 (fp32 containers, 32 elements per 128-byte K block,
// UMMA K = 8); HALF = true: FP16 (64 elements per K block, UMMA K = 16).  Both advance 32 bytes per MMA along K, so the
// shared-memory tiles, swizzle and descriptors are byte-identical; only the instruction kind / descriptor formats differ.
template <bool HALF>
struct Operand {
    static constexpr int kBytes = HALF ? 2 : 4;
    static constexpr int kBK = 128 / kBytes;   // elements per 128-byte K block
    // instruction descriptor: c_format F32 (bit 4), a/b format F16 = 0 or TF32 = 2 (bits 7, 10), N >> 3 at 17, M >> 4 at 24
    __device__ __forceinline__ static uint32_t idesc(int n, int m) {
        return (1u << 4) | ((HALF ? 0u : 2u) << 7) | ((HALF ? 0u : 2u) << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
    }
    template <bool PAIR>
    __device__ __forceinline__ static void mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc_, uint32_t accumulate) {
        if constexpr (HALF) {
            if constexpr (PAIR) umma_f16_2sm(tmem_d, adesc, bdesc, idesc_, accumulate); else umma_f16(tmem_d, adesc, bdesc, idesc_, accumulate);
        } else {
            if constexpr (PAIR) umma_tf32_2sm(tmem_d, adesc, bdesc, idesc_, accumulate); else umma_tf32(tmem_d, adesc, bdesc, idesc_, accumulate);
        }
    }
};
// arrive on the same-offset mbarrier of BOTH CTAs once all previously issued pair MMAs have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader CTA's copy of a barrier (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}

// ---------------------------------------------------------------------------------------------- host helpers
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

inline int encode_tmap(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, bool half = false, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = get_encode();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return SDC_ERR_CUDA; }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank); return SDC_ERR_CUDA; }
    return SDC_OK;
}

// output map of the TMA-store epilogue: out[M, Cout] row-major, box = 32 rows x 32 columns
inline int encode_out_tmap(CUtensorMap* map, const void* out, int64_t M, int Cout, bool out_half) {
    cuuint64_t dims[2] = {(cuuint64_t)Cout, (cuuint64_t)M};
    cuuint64_t str[1] = {(cuuint64_t)Cout * (out_half ? 2 : 4)};
    cuuint32_t box[2] = {32, 32};
    return encode_tmap(map, out, 2, dims, str, box, out_half, out_half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace sdc
