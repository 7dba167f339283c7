// On-device synthetic Burgers data: the fields of the reference generator
// make_data_varying_f (/root/reference/1D/data/generate_burgers.py:338-418) evaluated from its O(N) random parameters, and
// the dataset tensor assembly of BurgersDataset._process_data (/root/reference/1D/data/burgers.py:104-142).
//
// The reference draws a handful of scalars per instance from numpy's global RNG and then spends its time in float64
// exp() over [N, t, s] arrays on the host.  Here the host still draws the scalars (same RNG stream, same order -- the
// python layer does that), and the device evaluates the fields in float64 with the reference's operation order:
//   u0[n, x]   = a1 exp(-0.5 (x - l1)^2 / s1^2) + a2 exp(-0.5 (x - l2)^2 / s2^2)
//   f[n, j, x] = float32( sum_k (amp_k * (exp(-0.5 (x - lx_k)^2 / sx_k^2) * mask_x)) * (comp * exp(-0.5 (t_j - lt_k)^2 / st_k^2)) )
// so that results agree with numpy to the last float64 ulp of exp() (bit-identical after the float32 cast except at
// rounding ties).  Both kernels are write-bandwidth bound: 4 bytes per output element, coalesced.
#include "common.cuh"
#include <math.h>

namespace sdc {

// pu0: [N, 6] = (loc1, amp1, sig1, loc2, amp2, sig2); xg: [s] grid (float32 values of torch.linspace, promoted)
__global__ void __launch_bounds__(128) burgers_u0_kernel(const double* __restrict__ pu0, const float* __restrict__ xg,
                                                         double* __restrict__ u0_f64, float* __restrict__ u0_f32, int s) {
    const int64_t n = blockIdx.x;
    const double* p = pu0 + n * 6;
    for (int i = threadIdx.x; i < s; i += blockDim.x) {
        const double x = (double)xg[i];
        const double d1 = x - p[0], d2 = x - p[3];
        // explicit _rn intrinsics: no FMA contraction, numpy rounds every product and sum separately
        const double g1 = __dmul_rn(p[1], exp(-0.5 * (d1 * d1) / (p[2] * p[2])));
        const double g2 = __dmul_rn(p[4], exp(-0.5 * (d2 * d2) / (p[5] * p[5])));
        const double v = __dadd_rn(g1, g2);
        if (u0_f64) u0_f64[n * s + i] = v;
        if (u0_f32) u0_f32[n * s + i] = (float)v;
    }
}

// pf: [N, terms, 5] = (amp, loc_x, sig_x, loc_t, sig_t); tg: [t] time nodes; one CTA per (instance, time row)
__global__ void __launch_bounds__(128) burgers_f_kernel(const double* __restrict__ pf, const float* __restrict__ xg,
                                                        const float* __restrict__ tg, float* __restrict__ f, int s, int t,
                                                        int terms, double amp_compensate, int mask_mode, float alpha) {
    const int64_t n = blockIdx.x / t;
    const int j = blockIdx.x % t;
    extern __shared__ double et[];   // exp_time of this row per term
    if ((int)threadIdx.x < terms) {
        const double* p = pf + (n * terms + threadIdx.x) * 5;
        const double dt = (double)tg[j] - p[3];
        et[threadIdx.x] = amp_compensate * exp(-0.5 * (dt * dt) / (p[4] * p[4]));
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s; i += blockDim.x) {
        const double x = (double)xg[i];
        const double mask = (mask_mode == 0 || i < s / 4 || i >= 3 * s / 4) ? 1.0 : 0.0;
        double acc = 0.0;
        for (int k = 0; k < terms; ++k) {
            const double* p = pf + (n * terms + k) * 5;
            const double dx = x - p[1];
            const double es = exp(-0.5 * (dx * dx) / (p[2] * p[2])) * mask;
            const double term = __dmul_rn(__dmul_rn(p[0], es), et[k]);
            acc = (k == 0) ? term : __dadd_rn(acc, term);
        }
        float v = (float)acc;
        if (alpha != 1.0f) v = fminf(fmaxf(v * alpha, -10.0f), 10.0f);
        f[(n * t + j) * s + i] = v;
    }
}

// state[n, 0, :nt1] = u / scaler, state[n, 1, :nt] = f / scaler, state[n, 2, :nt1] = (max u^2 | u^2) / scaler, zero padding
__global__ void __launch_bounds__(256) dataset_states_kernel(const float* __restrict__ u, const float* __restrict__ f,
                                                             float* __restrict__ state, int nt1, int nt, int pad, int s,
                                                             float scaler, int use_max) {
    const int64_t n = blockIdx.x;
    const float* un = u + n * nt1 * s;
    const float* fn = f + n * nt * s;
    float* st = state + n * 3 * pad * s;
    __shared__ float red[8];
    __shared__ float smax;
    float m = -INFINITY;
    bool has_nan = false;
    if (use_max) {
        for (int i = threadIdx.x; i < nt1 * s; i += blockDim.x) {
            const float q = un[i] * un[i];
            has_nan |= (q != q);
            m = fmaxf(m, q);
        }
        if (has_nan) m = NAN;   // torch.amax propagates NaN
        // NaN-propagating max over the warp and the CTA
        for (int o = 16; o > 0; o >>= 1) {
            const float other = __shfl_xor_sync(0xffffffffu, m, o);
            m = (m != m || other != other) ? NAN : fmaxf(m, other);
        }
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = red[0];
            for (int w = 1; w < 8; ++w) t = (t != t || red[w] != red[w]) ? NAN : fmaxf(t, red[w]);
            smax = t;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < pad * s; i += blockDim.x) {
        const int row = i / s;
        const float uv = row < nt1 ? un[i] : 0.f;
        st[i] = uv / scaler;
        st[pad * s + i] = (row < nt ? fn[i] : 0.f) / scaler;
        st[2 * pad * s + i] = (row < nt1 ? (use_max ? smax : uv * uv) : 0.f) / scaler;
    }
}

}  // namespace sdc

using namespace sdc;

extern "C" int sdc_burgers_fields(const double* params_u0, const double* params_f, const float* x_grid, const float* t_grid,
                                  double* u0_f64, float* u0_f32, float* f, int64_t Nu0, int64_t Nf, int s, int t, int terms,
                                  double amp_compensate, int partial_control, float alpha, void* stream) {
    SDC_REQUIRE(s > 0 && t > 0 && terms > 0 && terms <= 128 && Nu0 >= 0 && Nf >= 0 && Nu0 < (1LL << 31) && Nf * t < (1LL << 31),
                "burgers_fields: bad sizes");
    SDC_REQUIRE(x_grid && (Nu0 == 0 || (params_u0 && (u0_f64 || u0_f32))) && (Nf == 0 || (params_f && t_grid && f)),
                "burgers_fields: null pointer");
    SDC_REQUIRE(partial_control == 0 || partial_control == 1, "burgers_fields: partial_control must be 0 (none) or 1 (front_rear_quarter)");
    if (Nu0 > 0) {
        burgers_u0_kernel<<<(unsigned)Nu0, 128, 0, as_stream(stream)>>>(params_u0, x_grid, u0_f64, u0_f32, s);
        SDC_LAUNCHED();
    }
    if (Nf > 0) {
        burgers_f_kernel<<<(unsigned)(Nf * t), 128, terms * sizeof(double), as_stream(stream)>>>(params_f, x_grid, t_grid, f, s, t, terms,
                                                                                               amp_compensate, partial_control, alpha);
        SDC_LAUNCHED();
    }
    return SDC_OK;
}

extern "C" int sdc_dataset_states(const float* u_traj, const float* f, float* state, int64_t N, int nt1, int nt, int pad, int s,
                                  float scaler, int use_max_safety, void* stream) {
    SDC_REQUIRE(N >= 0 && N < (1LL << 31) && nt1 > 0 && nt > 0 && pad >= nt1 && pad >= nt && s > 0, "dataset_states: bad sizes");
    if (N == 0) return SDC_OK;
    SDC_REQUIRE(u_traj && f && state, "dataset_states: null pointer");
    dataset_states_kernel<<<(unsigned)N, 256, 0, as_stream(stream)>>>(u_traj, f, state, nt1, nt, pad, s, scaler, use_max_safety);
    SDC_LAUNCHED();
    return SDC_OK;
}
