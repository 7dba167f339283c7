// Shared helpers for the safediffcon_b200 C-ABI library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/safediffcon_b200.h"

namespace sdc {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

#define SDC_REQUIRE(cond, ...)             \
    do {                                   \
        if (!(cond)) {                     \
            sdc::set_error(__VA_ARGS__);   \
            return SDC_ERR_ARG;            \
        }                                  \
    } while (0)

#define SDC_CUDA(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            sdc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return SDC_ERR_CUDA;                                                            \
        }                                                                                   \
    } while (0)

// call after every kernel launch: counts it and converts launch errors into a status code
#define SDC_LAUNCHED()                      \
    do {                                    \
        sdc::g_launches.fetch_add(1);       \
        SDC_CUDA(cudaGetLastError());       \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// round-to-nearest (ties away) fp32 -> tf32, kept in an fp32 container (low 13 mantissa bits zero)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// 4-element loads / stores of activations in either operand precision (fp32 containers or fp16)
__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __half* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float4 load4_nc(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 load4_nc(const __half* p) {
    const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// store as a tensor-core operand: TF32-rounded fp32 container, or fp16 (both round to nearest, 10-bit mantissa)
__device__ __forceinline__ void store_operand4(float* p, float4 v) {
    *reinterpret_cast<float4*>(p) = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}
__device__ __forceinline__ void store_operand4(__half* p, float4 v) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ float to_operand(float v, float) { return to_tf32(v); }
__device__ __forceinline__ __half to_operand(float v, __half) { return __float2half_rn(v); }
template <bool HALF> struct ActT { using type = float; };
template <> struct ActT<true> { using type = __half; };

// Philox4x32-10 counter RNG (Salmon et al. 2011), written out here so the stream is ours and reproducible
// independent of torch/cuRAND versions.
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ Philox(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    __device__ __forceinline__ uint4 operator()(uint4 c) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
            uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
            c = make_uint4(hi1 ^ c.y ^ a, lo1, hi0 ^ c.w ^ b, lo0);
            a += 0x9E3779B9u;
            b += 0xBB67AE85u;
        }
        return c;
    }
};

// four N(0,1) draws from one Philox block (Box-Muller on (0,1] uniforms)
__device__ __forceinline__ float4 normal4(const Philox& ph, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
    uint4 r = ph(make_uint4(c0, c1, c2, c3));
    const float s = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 1.0f) * s, u1 = (float)r.y * s;
    float u2 = ((float)r.z + 1.0f) * s, u3 = (float)r.w * s;
    float m0 = sqrtf(-2.0f * __logf(u0)), m1 = sqrtf(-2.0f * __logf(u2));
    float s0, c0f, s1, c1f;
    __sincosf(6.283185307179586f * u1, &s0, &c0f);
    __sincosf(6.283185307179586f * u3, &s1, &c1f);
    return make_float4(m0 * c0f, m0 * s0, m1 * c1f, m1 * s1);
}

}  // namespace sdc
