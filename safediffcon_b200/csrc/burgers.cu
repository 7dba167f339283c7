// Burgers rollout + scoring kernels.
//
// Reference behaviour: /root/reference/1D/data/generate_burgers.py:113-299 (explicit Euler, conservative
// central differences, Dirichlet-0 ghost cells) and /root/reference/1D/utils/metrics.py:29-94.
//
// Mapping: one warp owns one trajectory for all ~10,000 steps.  The s interior points live in registers,
// PPL = s/32 contiguous points per lane; the only inter-lane traffic per step is two shuffles for the halo.
// The forcing row of the current interval sits in registers and is reloaded once per snapshot interval;
// snapshots are written as coalesced 16-byte stores.  Nothing but u0, f (read once) and the snapshots
// (written once) touches HBM: 4*s*(1+nt+nt+1) bytes per trajectory.  The kernel is bound by the FP32 pipe
// and the per-step dependency chain, not by HBM (AI ~ 1600 flop/B) -- see DESIGN.md.
//
// STRICT mode reproduces the reference's fp32 op order with explicitly rounded mul/add (no FMA contraction):
//   us = u*u ; tr = (-a)*us[i-1] + a*us[i+1] ; di = (d*u[i-1] + d2*u[i]) + d*u[i+1]
//   u += dt*(((-0.5)*tr + di) + f)
// Products shared between neighbouring points (a*us[i], d*u[i]) are computed once; (-a)*x == -(a*x) exactly.
#include "common.cuh"
#include <math.h>

namespace sdc {

struct BurgersArgs {
    const float* u0; int64_t u0_stride; int64_t u0_div;     // u0 row of trajectory n: u0 + (n / u0_div) * u0_stride
    const float* f;  int64_t f_stride;  int64_t f_mod;      // f rows: f + (n % f_mod) * f_stride + k*s
    float* out;                                             // [N, nt+1, s] or null
    const float* target_final;                              // [N, s] or null
    float* J; int32_t* pts; int32_t* tms; int32_t* flag;    // per-trajectory scores or null
    int64_t N; int nt; int rec; int tail;                   // rec steps per interval, tail = extra steps after nt*rec
    float a, d, d2, dt, u_bound;
};

template <int PPL, bool STRICT>
__device__ __forceinline__ void euler_step(float (&u)[PPL], const float (&fk)[PPL], int lane, float a, float d, float d2,
                                           float dt) {
    float ul = __shfl_up_sync(0xffffffffu, u[PPL - 1], 1);
    float ur = __shfl_down_sync(0xffffffffu, u[0], 1);
    if (lane == 0) ul = 0.f;
    if (lane == 31) ur = 0.f;
    if constexpr (STRICT) {
        float pa[PPL + 2], qd[PPL + 2];  // a*u^2 and d*u at i-1 .. i+PPL
        pa[0] = __fmul_rn(a, __fmul_rn(ul, ul));
        qd[0] = __fmul_rn(d, ul);
        pa[PPL + 1] = __fmul_rn(a, __fmul_rn(ur, ur));
        qd[PPL + 1] = __fmul_rn(d, ur);
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            pa[i + 1] = __fmul_rn(a, __fmul_rn(u[i], u[i]));
            qd[i + 1] = __fmul_rn(d, u[i]);
        }
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            float tr = __fadd_rn(-pa[i], pa[i + 2]);
            float di = __fadd_rn(__fadd_rn(qd[i], __fmul_rn(d2, u[i])), qd[i + 2]);
            float rhs = __fadd_rn(__fadd_rn(__fmul_rn(-0.5f, tr), di), fk[i]);
            u[i] = __fadd_rn(u[i], __fmul_rn(dt, rhs));
        }
    } else {
        float us[PPL + 2], uu[PPL + 2];
        uu[0] = ul; uu[PPL + 1] = ur;
        us[0] = ul * ul; us[PPL + 1] = ur * ur;
#pragma unroll
        for (int i = 0; i < PPL; ++i) { uu[i + 1] = u[i]; us[i + 1] = u[i] * u[i]; }
        const float ha = -0.5f * a;
#pragma unroll
        for (int i = 0; i < PPL; ++i) {
            float rhs = fmaf(d2, uu[i + 1], fk[i]);
            rhs = fmaf(d, uu[i] + uu[i + 2], rhs);
            rhs = fmaf(ha, us[i + 2] - us[i], rhs);
            u[i] = fmaf(dt, rhs, uu[i + 1]);
        }
    }
}

template <int PPL>
__device__ __forceinline__ void load_row(float (&v)[PPL], const float* p, int lane) {
    if constexpr (PPL == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p) + lane);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else if constexpr (PPL == 2) {
        float2 t = __ldg(reinterpret_cast<const float2*>(p) + lane);
        v[0] = t.x; v[1] = t.y;
    } else {
#pragma unroll
        for (int i = 0; i < PPL; ++i) v[i] = __ldg(p + lane * PPL + i);
    }
}
template <int PPL>
__device__ __forceinline__ void store_row(const float (&v)[PPL], float* p, int lane) {
    if constexpr (PPL == 4) {
        reinterpret_cast<float4*>(p)[lane] = make_float4(v[0], v[1], v[2], v[3]);
    } else if constexpr (PPL == 2) {
        reinterpret_cast<float2*>(p)[lane] = make_float2(v[0], v[1]);
    } else {
#pragma unroll
        for (int i = 0; i < PPL; ++i) p[lane * PPL + i] = v[i];
    }
}

template <int PPL>
__device__ __forceinline__ void count_row(const float (&v)[PPL], float bound, int& pts, int& tms) {
    int c = 0;
#pragma unroll
    for (int i = 0; i < PPL; ++i) c += (fabsf(v[i]) > bound) ? 1 : 0;
    c = warp_sum_i(c);
    pts += c;
    tms += (c > 0) ? 1 : 0;
}

// Diagnostics (the reference has no failure detection; a diverged rollout silently yields NaN metrics):
// number of rollouts whose FINAL state holds a non-finite value, per device, since the last reset.  NaN/Inf stay data -- the
// outputs are exactly the reference's -- this only makes them countable without a pass over the trajectories.
__device__ unsigned long long g_nonfinite_rollouts = 0ull;

template <int PPL, bool STRICT>
__global__ void __launch_bounds__(256) burgers_rollout_kernel(BurgersArgs p) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= p.N) return;
    constexpr int S = PPL * 32;
    const bool score = (p.J != nullptr) || (p.pts != nullptr) || (p.tms != nullptr) || (p.flag != nullptr);
    float u[PPL], fk[PPL];
    load_row<PPL>(u, p.u0 + (n / p.u0_div) * p.u0_stride, lane);
    float* on = p.out ? p.out + n * (int64_t)(p.nt + 1) * S : nullptr;
    if (on) store_row<PPL>(u, on, lane);
    int pts = 0, tms = 0;
    if (score) count_row<PPL>(u, p.u_bound, pts, tms);
    const float* fn = p.f + (n % p.f_mod) * p.f_stride;
    for (int k = 0; k < p.nt; ++k) {
        load_row<PPL>(fk, fn + (int64_t)k * S, lane);
#pragma unroll 2
        for (int j = 0; j < p.rec; ++j) euler_step<PPL, STRICT>(u, fk, lane, p.a, p.d, p.d2, p.dt);
        if (on) store_row<PPL>(u, on + (int64_t)(k + 1) * S, lane);
        if (score) count_row<PPL>(u, p.u_bound, pts, tms);
    }
    {
        bool bad = false;
#pragma unroll
        for (int i = 0; i < PPL; ++i) bad |= !isfinite(u[i]);
        if (__any_sync(0xffffffffu, bad) && lane == 0) atomicAdd(&g_nonfinite_rollouts, 1ull);
    }
    // reference snapshot k is taken after step (k+1)*rec; steps beyond nt*rec reuse the last forcing row
    // and are not recorded -- the final-state score below therefore uses snapshot nt, like the reference.
    if (score && lane == 0) {
        if (p.pts) p.pts[n] = pts;
        if (p.tms) p.tms[n] = tms;
        if (p.flag) p.flag[n] = tms > 0;
    }
    if (p.J && p.target_final) {
        float tg[PPL];
        load_row<PPL>(tg, p.target_final + n * S, lane);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < PPL; ++i) { float dl = tg[i] - u[i]; acc += dl * dl; }
        acc = warp_sum(acc);
        if (lane == 0) p.J[n] = acc / (float)S;
    }
}

template <int PPL>
__global__ void __launch_bounds__(256) burgers_score_kernel(const float* traj, const float* target_final, float bound,
                                                            int64_t N, int nt1, float* J, int32_t* pts_o, int32_t* tms_o,
                                                            int32_t* flag_o) {
    const int lane = threadIdx.x & 31;
    const int64_t n = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    constexpr int S = PPL * 32;
    const float* t = traj + n * (int64_t)nt1 * S;
    float v[PPL];
    int pts = 0, tms = 0;
    for (int r = 0; r < nt1; ++r) {
        load_row<PPL>(v, t + (int64_t)r * S, lane);
        count_row<PPL>(v, bound, pts, tms);
    }
    if (lane == 0) {
        if (pts_o) pts_o[n] = pts;
        if (tms_o) tms_o[n] = tms;
        if (flag_o) flag_o[n] = tms > 0;
    }
    if (J && target_final) {
        float tg[PPL];
        load_row<PPL>(tg, target_final + n * S, lane);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < PPL; ++i) { float dl = tg[i] - v[i]; acc += dl * dl; }
        acc = warp_sum(acc);
        if (lane == 0) J[n] = acc / (float)S;
    }
}

static int launch_rollout(BurgersArgs& a, int s, double visc, double T, double dt, int strict, void* stream) {
    SDC_REQUIRE(a.N >= 0 && a.nt > 0, "burgers: N >= 0 and nt > 0 required");
    SDC_REQUIRE(s % 32 == 0 && s >= 32 && s <= 256 && (s / 32 == 1 || s / 32 == 2 || s / 32 == 4 || s / 32 == 8),
                "burgers: s=%d unsupported (need s in {32,64,128,256})", s);
    SDC_REQUIRE(dt > 0 && T > 0, "burgers: T, dt must be positive");
    if (a.N == 0) return SDC_OK;
    SDC_REQUIRE(a.u0 && a.f, "burgers: null input pointer");
    const double dx = 1.0 / (double)(s + 1);
    a.a = (float)(1.0 / (2.0 * dx));
    a.d = (float)(visc * 1.0 / (dx * dx));
    a.d2 = (float)(visc * -2.0 / (dx * dx));
    a.dt = (float)dt;
    const int steps = (int)ceil(T / dt);
    a.rec = steps / a.nt;
    a.tail = steps - a.rec * a.nt;
    SDC_REQUIRE(a.rec > 0, "burgers: fewer steps (%d) than forcing intervals (%d)", steps, a.nt);
    SDC_REQUIRE(a.tail == 0, "burgers: steps=%d not a multiple of nt=%d (the reference indexes past f)", steps, a.nt);
    const int wpb = 8;
    dim3 grid((unsigned)((a.N + wpb - 1) / wpb)), block(wpb * 32);
    cudaStream_t st = as_stream(stream);
#define SDC_ROLL(PPL_)                                                                  \
    if (strict) burgers_rollout_kernel<PPL_, true><<<grid, block, 0, st>>>(a);          \
    else burgers_rollout_kernel<PPL_, false><<<grid, block, 0, st>>>(a);
    switch (s / 32) {
        case 1: SDC_ROLL(1) break;
        case 2: SDC_ROLL(2) break;
        case 4: SDC_ROLL(4) break;
        default: SDC_ROLL(8) break;
    }
#undef SDC_ROLL
    SDC_LAUNCHED();
    return SDC_OK;
}

}  // namespace sdc

using namespace sdc;

extern "C" int sdc_burgers_nonfinite_rollouts(int reset, int64_t* count_host) {
    unsigned long long v = 0ull;
    SDC_CUDA(cudaMemcpyFromSymbol(&v, g_nonfinite_rollouts, sizeof(v)));   // synchronises with the legacy default stream only
    if (count_host) *count_host = (int64_t)v;
    if (reset) { v = 0ull; SDC_CUDA(cudaMemcpyToSymbol(g_nonfinite_rollouts, &v, sizeof(v))); }
    return SDC_OK;
}

extern "C" int sdc_burgers_solve_free(const float* u0, const float* f, float* out, int64_t N, int s, int nt, double visc,
                                      double T, double dt, int strict, void* stream) {
    BurgersArgs a{};
    a.u0 = u0; a.u0_stride = s; a.u0_div = 1;
    a.f = f; a.f_stride = (int64_t)nt * s; a.f_mod = N > 0 ? N : 1;
    a.out = out; a.N = N; a.nt = nt;
    SDC_REQUIRE(out != nullptr || N == 0, "burgers_solve_free: null output");
    return launch_rollout(a, s, visc, T, dt, strict, stream);
}

extern "C" int sdc_burgers_solve_cartesian(const float* u0, const float* f, float* out, int64_t Nu0, int64_t Nf, int s,
                                           int nt, double visc, double T, double dt, int strict, void* stream) {
    BurgersArgs a{};
    a.u0 = u0; a.u0_stride = s; a.u0_div = Nf > 0 ? Nf : 1;
    a.f = f; a.f_stride = (int64_t)nt * s; a.f_mod = Nf > 0 ? Nf : 1;
    a.out = out; a.N = Nu0 * Nf; a.nt = nt;
    SDC_REQUIRE(out != nullptr || a.N == 0, "burgers_solve_cartesian: null output");
    return launch_rollout(a, s, visc, T, dt, strict, stream);
}

extern "C" int sdc_burgers_control_score(const float* diffused, int pad, const float* target_final, float u_bound,
                                         float* out, int64_t N, int s, int nt, double visc, double T, double dt,
                                         int strict, float* J, int32_t* pts, int32_t* tms, int32_t* flag, void* stream) {
    SDC_REQUIRE(pad >= nt + 1, "control_score: pad=%d < nt+1", pad);
    BurgersArgs a{};
    const int64_t sample = 3LL * pad * s;
    a.u0 = diffused; a.u0_stride = sample; a.u0_div = 1;
    a.f = diffused ? diffused + (int64_t)pad * s : nullptr; a.f_stride = sample; a.f_mod = N > 0 ? N : 1;
    a.out = out; a.N = N; a.nt = nt;
    a.target_final = target_final; a.J = J; a.pts = pts; a.tms = tms; a.flag = flag; a.u_bound = u_bound;
    SDC_REQUIRE(J == nullptr || target_final != nullptr, "control_score: J requested without target_final");
    return launch_rollout(a, s, visc, T, dt, strict, stream);
}

extern "C" int sdc_burgers_score(const float* traj, const float* target_final, float u_bound, int64_t N, int nt1, int s,
                                 float* J, int32_t* pts, int32_t* tms, int32_t* flag, void* stream) {
    SDC_REQUIRE(s % 32 == 0 && (s / 32 == 1 || s / 32 == 2 || s / 32 == 4 || s / 32 == 8),
                "burgers_score: s=%d unsupported", s);
    SDC_REQUIRE(N >= 0 && nt1 > 0, "burgers_score: bad sizes");
    if (N == 0) return SDC_OK;
    SDC_REQUIRE(traj != nullptr, "burgers_score: null trajectory");
    SDC_REQUIRE(J == nullptr || target_final != nullptr, "burgers_score: J requested without target_final");
    const int wpb = 8;
    dim3 grid((unsigned)((N + wpb - 1) / wpb)), block(wpb * 32);
    cudaStream_t st = as_stream(stream);
    switch (s / 32) {
        case 1: burgers_score_kernel<1><<<grid, block, 0, st>>>(traj, target_final, u_bound, N, nt1, J, pts, tms, flag); break;
        case 2: burgers_score_kernel<2><<<grid, block, 0, st>>>(traj, target_final, u_bound, N, nt1, J, pts, tms, flag); break;
        case 4: burgers_score_kernel<4><<<grid, block, 0, st>>>(traj, target_final, u_bound, N, nt1, J, pts, tms, flag); break;
        default: burgers_score_kernel<8><<<grid, block, 0, st>>>(traj, target_final, u_bound, N, nt1, J, pts, tms, flag); break;
    }
    SDC_LAUNCHED();
    return SDC_OK;
}
