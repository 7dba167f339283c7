// Fused reverse-diffusion step.
//
// One launch replaces ~25 ATen kernels + an autograd graph + a host sync per step of the reference:
//   model_predictions   /root/reference/1D/model/diffusion.py:226-286
//   safety guidance     /root/reference/1D/utils/guidance.py:58-86   (closed-form gradient)
//   DDIM update         diffusion.py:500-510      DDPM update  diffusion.py:288-306
//   condition writes    diffusion.py:336-366
// One CTA per sample (the guidance statistic is a per-sample reduction over the safety channel); float4
// coalesced traffic: reads x_t, eps (+ noise) and writes x_{t-1} = 4 x 24.6 KB per sample per step.
// All arithmetic is explicitly rounded fp32 mul/add/div in the reference's association so that, given the same
// eps and noise, the result is bit-identical to the CPU reference except for the summation order of the
// guidance mean (which only feeds a threshold).
#include "common.cuh"

namespace sdc {

struct StepArgs {
    const float* x; const float* eps; const float* noise; float* out; float* x0_out; float* eps_out;
    const sdc_step_coef* coef; int step; const int32_t* counter;
    const sdc_chain_state* state;  // device-resident (step, seed, sample_offset): captured-graph replays
    sdc_guidance g; float g_unit;  // g_unit = fp32(w_score*scaler/(nt*W)) for mode 1
    const float* grad; const float* u_init; const float* u_final; const float* w_gt;
    int cond_idx; int pad_writes; int clip_denoised; int sampler;
    uint64_t seed; int64_t sample_offset; int H; int W;
};

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }
// torch.clamp propagates NaN; fminf/fmaxf drop it -> keep NaN explicitly
__device__ __forceinline__ float clamp1_nan(float v) { return (v != v) ? v : clamp1(v); }

__device__ __forceinline__ float apply_condition(float v, int c, int h, int w, int64_t b, const StepArgs& p) {
    if (c == 0) {
        if (h == 0 && p.u_init) v = p.u_init[b * p.W + w];
        if (h == p.cond_idx && p.u_final) v = p.u_final[b * p.W + w];
        if (p.pad_writes && h > p.cond_idx) v = 0.f;
    } else if (c == 1) {
        if (p.w_gt) v = p.w_gt[(b * p.H + h) * p.W + w];
        if (p.pad_writes && h >= p.cond_idx) v = 0.f;
    } else {
        if (p.pad_writes && h >= p.cond_idx) v = 0.f;
    }
    return v;
}

__global__ void __launch_bounds__(256) reverse_step_kernel(StepArgs p) {
    const int64_t b = blockIdx.x;
    const int HW = p.H * p.W;
    const int n4 = 3 * HW / 4;
    const sdc_step_coef cf = p.coef[p.state ? p.state->step : (p.counter ? *p.counter : p.step)];
    const bool ddim = p.sampler == SDC_SAMPLER_DDIM;
    const bool clip = ddim;  // clip_x_start is set by ddim_sample only
    const float4* x4 = reinterpret_cast<const float4*>(p.x) + b * n4;
    const float4* e4 = reinterpret_cast<const float4*>(p.eps) + b * n4;
    __shared__ float red_a[8];
    __shared__ int red_c[8];
    __shared__ float s_gval, s_max;

    // ---- pass 1: per-sample safety statistic of the first x0 estimate (modes 1, 2) ----
    float gval = 0.f, vmax = 0.f;
    if (p.g.mode == 1 || p.g.mode == 2) {
        const int lo4 = 2 * HW / 4, hi4 = lo4 + p.g.nt * p.W / 4;
        float acc = (p.g.mode == 1) ? 0.f : -INFINITY;
        for (int i = lo4 + threadIdx.x; i < hi4; i += blockDim.x) {
            float4 xv = x4[i], ev = e4[i];
            float v[4] = {xv.x, xv.y, xv.z, xv.w}, e[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float x0 = __fsub_rn(__fmul_rn(cf.c1, v[j]), __fmul_rn(cf.c2, e[j]));
                if (clip) x0 = clamp1_nan(x0);
                float sv = __fmul_rn(x0, p.g.scaler);
                acc = (p.g.mode == 1) ? acc + sv : fmaxf(acc, sv);
            }
        }
        acc = (p.g.mode == 1) ? warp_sum(acc) : warp_max(acc);
        if ((threadIdx.x & 31) == 0) red_a[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = red_a[0];
            for (int w = 1; w < (blockDim.x >> 5); ++w) t = (p.g.mode == 1) ? t + red_a[w] : fmaxf(t, red_a[w]);
            float stat = (p.g.mode == 1) ? t / (float)(p.g.nt * p.W) : t;
            float margin = __fsub_rn(__fadd_rn(stat, p.g.Q), p.g.u_bound_sq);
            float on = margin > 0.f ? 1.f : (margin == 0.f ? 0.5f : 0.f);  // torch.maximum splits the gradient at ties
            s_gval = on;
            s_max = t;
        }
        __syncthreads();
        const float on = s_gval;
        vmax = s_max;
        if (p.g.mode == 1) {
            gval = __fmul_rn(__fmul_rn(on, p.g_unit), cf.sched);
        } else {
            // amax backward spreads the gradient evenly over the tied maxima: count them
            int cnt = 0;
            for (int i = lo4 + threadIdx.x; i < hi4; i += blockDim.x) {
                float4 xv = x4[i], ev = e4[i];
                float v[4] = {xv.x, xv.y, xv.z, xv.w}, e[4] = {ev.x, ev.y, ev.z, ev.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x0 = __fsub_rn(__fmul_rn(cf.c1, v[j]), __fmul_rn(cf.c2, e[j]));
                    if (clip) x0 = clamp1_nan(x0);
                    cnt += (__fmul_rn(x0, p.g.scaler) == vmax) ? 1 : 0;
                }
            }
            cnt = warp_sum_i(cnt);
            __syncthreads();
            if ((threadIdx.x & 31) == 0) red_c[threadIdx.x >> 5] = cnt;
            __syncthreads();
            int tot = 0;
            for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red_c[w];
            // autograd order: (w_score*on) / count, then * scaler, then * sched
            gval = __fmul_rn(__fmul_rn(__fdiv_rn(__fmul_rn(p.g.w_score, on), (float)max(tot, 1)), p.g.scaler), cf.sched);
        }
    }

    // ---- pass 2: elementwise update ----
    const Philox ph(p.state ? p.state->seed : p.seed);
    const int64_t gs = (p.state ? p.state->sample_offset : p.sample_offset) + b;
    const float4* z4 = p.noise ? reinterpret_cast<const float4*>(p.noise) + b * n4 : nullptr;
    const float4* g4 = (p.g.mode == 3 && p.grad) ? reinterpret_cast<const float4*>(p.grad) + b * n4 : nullptr;
    float4* o4 = reinterpret_cast<float4*>(p.out) + b * n4;
    const bool need_noise = !cf.is_last;
    const bool write_cond = !cf.is_last;
    const int lo = 2 * HW, hi = lo + p.g.nt * p.W;  // element range carrying the safety gradient
    for (int i = threadIdx.x; i < n4; i += blockDim.x) {
        float4 xv = x4[i], ev = e4[i];
        float4 zv = make_float4(0.f, 0.f, 0.f, 0.f), gv = zv;
        if (need_noise) zv = z4 ? z4[i] : normal4(ph, (uint32_t)i, (uint32_t)gs, (uint32_t)cf.t, (uint32_t)(gs >> 32));
        if (g4) gv = g4[i];
        float v[4] = {xv.x, xv.y, xv.z, xv.w}, e[4] = {ev.x, ev.y, ev.z, ev.w}, z[4] = {zv.x, zv.y, zv.z, zv.w};
        float gg[4] = {gv.x, gv.y, gv.z, gv.w}, r[4], x0o[4], eo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = i * 4 + j;
            const float c1x = __fmul_rn(cf.c1, v[j]);
            float en = e[j];
            if (p.g.mode != 0) {
                float x0a = __fsub_rn(c1x, __fmul_rn(cf.c2, en));
                if (clip) x0a = clamp1_nan(x0a);
                float g = 0.f;
                if (p.g.mode == 1) g = (idx >= lo && idx < hi) ? gval : 0.f;
                else if (p.g.mode == 2) g = (idx >= lo && idx < hi && __fmul_rn(x0a, p.g.scaler) == vmax) ? gval : 0.f;
                else g = __fmul_rn(gg[j], cf.sched);
                en = __fadd_rn(en, g);
            }
            float x0 = __fsub_rn(c1x, __fmul_rn(cf.c2, en));
            if (clip) {
                x0 = clamp1_nan(x0);
                en = __fdiv_rn(__fsub_rn(c1x, x0), cf.c2);
            }
            float o;
            if (ddim) {
                o = cf.is_last ? x0
                               : __fadd_rn(__fadd_rn(__fmul_rn(x0, cf.k_x0), __fmul_rn(cf.k_eps, en)), __fmul_rn(cf.k_noise, z[j]));
            } else {
                if (p.clip_denoised) x0 = clamp1_nan(x0);
                float mean = __fadd_rn(__fmul_rn(cf.k_x0, x0), __fmul_rn(cf.k_eps, v[j]));
                o = __fadd_rn(mean, __fmul_rn(cf.k_noise, z[j]));
            }
            if (write_cond) {
                const int c = idx / HW, rem = idx - c * HW;
                o = apply_condition(o, c, rem / p.W, rem % p.W, b, p);
            }
            r[j] = o; x0o[j] = x0; eo[j] = en;
        }
        o4[i] = make_float4(r[0], r[1], r[2], r[3]);
        if (p.x0_out) reinterpret_cast<float4*>(p.x0_out)[b * n4 + i] = make_float4(x0o[0], x0o[1], x0o[2], x0o[3]);
        if (p.eps_out) reinterpret_cast<float4*>(p.eps_out)[b * n4 + i] = make_float4(eo[0], eo[1], eo[2], eo[3]);
    }
}

__global__ void __launch_bounds__(256) write_conditions_kernel(StepArgs p, float* x) {
    const int64_t b = blockIdx.x;
    const int HW = p.H * p.W, n = 3 * HW;
    for (int idx = threadIdx.x; idx < n; idx += blockDim.x) {
        const int c = idx / HW, rem = idx - c * HW;
        float v = x[b * n + idx];
        float o = apply_condition(v, c, rem / p.W, rem % p.W, b, p);
        if (o != v || (o != o) != (v != v)) x[b * n + idx] = o;
    }
}

__global__ void __launch_bounds__(256) fill_normal_kernel(float* x, int64_t per4, uint64_t seed, int64_t sample_offset,
                                                          int32_t t_tag) {
    const int64_t b = blockIdx.y;
    const Philox ph(seed);
    const int64_t gs = sample_offset + b;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 z = normal4(ph, (uint32_t)i, (uint32_t)gs, (uint32_t)t_tag, (uint32_t)(gs >> 32));
        reinterpret_cast<float4*>(x)[b * per4 + i] = z;
    }
}

__global__ void advance_counter_kernel(int32_t* c) { *c += 1; }

// state <- (step, seed, offset) when `set`, else state->step += 1; then t_index[0..B) <- coef[min(step, n_steps-1)].t
__global__ void __launch_bounds__(256) chain_state_kernel(sdc_chain_state* state, int set, int32_t step, uint64_t seed,
                                                          int64_t sample_offset, const sdc_step_coef* coef, int n_steps,
                                                          int32_t* t_index, int64_t B) {
    __shared__ int32_t s_t;
    if (threadIdx.x == 0) {
        int32_t st = set ? step : state->step + 1;
        if (set) { state->seed = seed; state->sample_offset = sample_offset; state->reserved = 0; }
        state->step = st;
        s_t = coef[st < n_steps ? st : n_steps - 1].t;
    }
    __syncthreads();
    if (t_index) for (int64_t i = threadIdx.x; i < B; i += blockDim.x) t_index[i] = s_t;
}

}  // namespace sdc

using namespace sdc;

static int reverse_step_launch(int sampler, const float* x, const float* eps, const float* noise, float* out,
                               float* x0_out, float* eps_out, const sdc_step_coef* coef, int step,
                               const int32_t* step_counter, const sdc_chain_state* state, const sdc_guidance* guidance,
                               const float* grad, const float* u_init, const float* u_final, const float* w_gt, int cond_idx,
                               int pad_writes, int clip_denoised, uint64_t seed, int64_t sample_offset, int64_t B, int H, int W,
                               void* stream) {
    SDC_REQUIRE(sampler == SDC_SAMPLER_DDIM || sampler == SDC_SAMPLER_DDPM, "reverse_step: unknown sampler %d", sampler);
    SDC_REQUIRE(B >= 0 && H > 0 && W > 0 && W % 4 == 0, "reverse_step: need W %% 4 == 0 (got H=%d W=%d)", H, W);
    SDC_REQUIRE(B < (1LL << 31), "reverse_step: batch too large");
    if (B == 0) return SDC_OK;
    SDC_REQUIRE(x && eps && out && coef, "reverse_step: null pointer");
    StepArgs p{};
    p.x = x; p.eps = eps; p.noise = noise; p.out = out; p.x0_out = x0_out; p.eps_out = eps_out;
    p.coef = coef; p.step = step; p.counter = step_counter; p.state = state;
    if (guidance) p.g = *guidance; else { p.g.mode = 0; p.g.nt = 0; p.g.scaler = 1.f; }
    SDC_REQUIRE(p.g.mode >= 0 && p.g.mode <= 3, "reverse_step: bad guidance mode %d", p.g.mode);
    if (p.g.mode == 1 || p.g.mode == 2) {
        SDC_REQUIRE(p.g.nt > 0 && p.g.nt <= H && (p.g.nt * W) % 4 == 0, "reverse_step: bad guidance nt=%d", p.g.nt);
        p.g_unit = (float)((double)p.g.w_score * (double)p.g.scaler / (double)(p.g.nt * W));
    }
    SDC_REQUIRE(p.g.mode != 3 || grad != nullptr, "reverse_step: mode 3 needs a gradient tensor");
    p.grad = grad; p.u_init = u_init; p.u_final = u_final; p.w_gt = w_gt;
    p.cond_idx = cond_idx; p.pad_writes = pad_writes; p.clip_denoised = clip_denoised; p.sampler = sampler;
    p.seed = seed; p.sample_offset = sample_offset; p.H = H; p.W = W;
    reverse_step_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(p);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_reverse_step(int sampler, const float* x, const float* eps, const float* noise, float* out,
                                float* x0_out, float* eps_out, const sdc_step_coef* coef, int step,
                                const int32_t* step_counter, const sdc_guidance* guidance, const float* grad,
                                const float* u_init, const float* u_final, const float* w_gt, int cond_idx, int pad_writes,
                                int clip_denoised, uint64_t seed, int64_t sample_offset, int64_t B, int H, int W,
                                void* stream) {
    return reverse_step_launch(sampler, x, eps, noise, out, x0_out, eps_out, coef, step, step_counter, nullptr, guidance, grad,
                               u_init, u_final, w_gt, cond_idx, pad_writes, clip_denoised, seed, sample_offset, B, H, W, stream);
}

extern "C" int sdc_reverse_step_state(int sampler, const float* x, const float* eps, const float* noise, float* out,
                                      float* x0_out, float* eps_out, const sdc_step_coef* coef, const sdc_chain_state* state,
                                      const sdc_guidance* guidance, const float* grad, const float* u_init,
                                      const float* u_final, const float* w_gt, int cond_idx, int pad_writes, int clip_denoised,
                                      int64_t B, int H, int W, void* stream) {
    SDC_REQUIRE(state != nullptr, "reverse_step_state: null chain state");
    return reverse_step_launch(sampler, x, eps, noise, out, x0_out, eps_out, coef, 0, nullptr, state, guidance, grad, u_init,
                               u_final, w_gt, cond_idx, pad_writes, clip_denoised, 0, 0, B, H, W, stream);
}

extern "C" int sdc_chain_state_set(sdc_chain_state* state, int32_t step, uint64_t seed, int64_t sample_offset,
                                   const sdc_step_coef* coef, int n_steps, int32_t* t_index, int64_t B, void* stream) {
    SDC_REQUIRE(state && coef && n_steps > 0 && step >= 0 && B >= 0, "chain_state_set: bad arguments");
    chain_state_kernel<<<1, 256, 0, as_stream(stream)>>>(state, 1, step, seed, sample_offset, coef, n_steps, t_index, B);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_chain_state_advance(sdc_chain_state* state, const sdc_step_coef* coef, int n_steps, int32_t* t_index,
                                       int64_t B, void* stream) {
    SDC_REQUIRE(state && coef && n_steps > 0 && B >= 0, "chain_state_advance: bad arguments");
    chain_state_kernel<<<1, 256, 0, as_stream(stream)>>>(state, 0, 0, 0, 0, coef, n_steps, t_index, B);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_write_conditions(float* x, const float* u_init, const float* u_final, const float* w_gt, int cond_idx,
                                    int pad_writes, int64_t B, int H, int W, void* stream) {
    SDC_REQUIRE(B >= 0 && H > 0 && W > 0 && B < (1LL << 31), "write_conditions: bad sizes");
    if (B == 0) return SDC_OK;
    SDC_REQUIRE(x != nullptr, "write_conditions: null pointer");
    StepArgs p{};
    p.u_init = u_init; p.u_final = u_final; p.w_gt = w_gt; p.cond_idx = cond_idx; p.pad_writes = pad_writes;
    p.H = H; p.W = W;
    write_conditions_kernel<<<(unsigned)B, 256, 0, as_stream(stream)>>>(p, x);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_fill_normal(float* x, int64_t B, int64_t per_sample, uint64_t seed, int64_t sample_offset,
                               int32_t t_tag, void* stream) {
    SDC_REQUIRE(B >= 0 && per_sample > 0 && per_sample % 4 == 0 && B < 65536, "fill_normal: per_sample %% 4 == 0, B < 65536");
    if (B == 0) return SDC_OK;
    SDC_REQUIRE(x != nullptr, "fill_normal: null pointer");
    const int64_t per4 = per_sample / 4;
    dim3 grid((unsigned)((per4 + 255) / 256 > 64 ? 64 : (per4 + 255) / 256), (unsigned)B);
    fill_normal_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, per4, seed, sample_offset, t_tag);
    SDC_LAUNCHED();
    return SDC_OK;
}

extern "C" int sdc_advance_counter(int32_t* counter, void* stream) {
    SDC_REQUIRE(counter != nullptr, "advance_counter: null pointer");
    advance_counter_kernel<<<1, 1, 0, as_stream(stream)>>>(counter);
    SDC_LAUNCHED();
    return SDC_OK;
}
