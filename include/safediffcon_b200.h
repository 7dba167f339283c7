/* safediffcon_b200 -- C ABI of the B200-native SafeDiffCon 1D-Burgers hot path.
 *
 * The reference (AI4Science-WestlakeU/safediffcon) has no FFI layer: its hot path is Python calling PyTorch.
 * This header is the drop-in boundary a maintainer binds with ctypes; every entry point
 * names the reference interface it replaces.  Conventions:
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer (fp32, contiguous) unless it is
 *     marked "host"; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, >0 = SDC_ERR_* below; sdc_last_error() returns a thread-local message.
 *     NaN/Inf in the data are data, not errors (they propagate exactly as in the reference);
 *   - no entry point synchronises the stream or allocates device memory unless documented; all are
 *     CUDA-graph capturable; the library keeps no global mutable state apart from the last-error string
 *     and handles created by sdc_unet_create.
 * There is NO CPU fallback: without a CUDA device every compute entry point returns SDC_ERR_CUDA.
 */
#ifndef SAFEDIFFCON_B200_H
#define SAFEDIFFCON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDC_OK 0
#define SDC_ERR_ARG 1     /* bad shape / unsupported size / null pointer */
#define SDC_ERR_CUDA 2    /* CUDA runtime or driver error (message in sdc_last_error) */
#define SDC_ERR_STATE 3   /* handle used before weights were packed, workspace too small, ... */

int sdc_version(void);
const char* sdc_last_error(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
int64_t sdc_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Burgers rollout.  Replaces burgers_numeric_solve_free (1D/data/generate_burgers.py:207-299):
 * explicit Euler, conservative central differences, Dirichlet-0 ghosts, dx = 1/(s+1),
 * steps = ceil(T/dt), forcing row k = step / (steps/nt), snapshot after every (steps/nt) steps.
 * u0:[N,s]  f:[N,nt,s]  out:[N,nt+1,s] (row 0 = u0).  s must be a multiple of 32, s <= 256.
 * strict != 0: reference op order, no FMA contraction -> bit-identical to the fp32 CPU reference.
 * strict == 0: FMA-contracted fast mode (within 1e-5 relative of the reference on dataset-like inputs).
 */
int sdc_burgers_solve_free(const float* u0, const float* f, float* out, int64_t N, int s, int nt,
                           double visc, double T, double dt, int strict, void* stream);

/* Diagnostics: number of rollouts (all solver entry points, current device) whose final state held NaN/Inf since the last
 * reset; *count_host receives it (host pointer, may be NULL).  The data path is untouched: non-finite values propagate exactly
 * as in the reference, this only counts them.  Call after synchronising the stream(s) the rollouts ran on. */
int sdc_burgers_nonfinite_rollouts(int reset, int64_t* count_host);

/* Cartesian variant, replaces burgers_numeric_solve (generate_burgers.py:113-205):
 * u0:[Nu0,s]  f:[Nf,nt,s]  out:[Nu0,Nf,nt+1,s]. */
int sdc_burgers_solve_cartesian(const float* u0, const float* f, float* out, int64_t Nu0, int64_t Nf, int s, int nt,
                                double visc, double T, double dt, int strict, void* stream);

/* Scoring, replaces the reductions of evaluate_samples / calculate_safety_metrics (1D/utils/metrics.py:29-34,
 * 77-92).  traj:[N,nt1,s]  target_final:[N,s] (the last row of u_target).  Per trajectory:
 * J[n] = mean_x (target_final - traj[n,nt1-1])^2 ; exceed_points[n] = #{|traj|>u_bound} ;
 * exceed_times[n] = #{rows with any exceed} ; exceed_flag[n] = any.  Any output pointer may be NULL. */
int sdc_burgers_score(const float* traj, const float* target_final, float u_bound, int64_t N, int nt1, int s,
                      float* J, int32_t* exceed_points, int32_t* exceed_times, int32_t* exceed_flag, void* stream);

/* control_trajectories + evaluate_samples in one launch (1D/utils/metrics.py:42-65 then :29-94):
 * diffused:[N,3,pad,s] UNSCALED model output (u0 = diffused[:,0,0,:], f = diffused[:,1,:nt,:]); rolls out and
 * scores against target_final without a second pass over the trajectory.  out may be NULL (scores only). */
int sdc_burgers_control_score(const float* diffused, int pad, const float* target_final, float u_bound, float* out,
                              int64_t N, int s, int nt, double visc, double T, double dt, int strict,
                              float* J, int32_t* exceed_points, int32_t* exceed_times, int32_t* exceed_flag,
                              void* stream);

/* ------------------------------------------------------------------------------------------------
 * Reverse-diffusion step.  Replaces, fused into one launch, GaussianDiffusion.model_predictions
 * (1D/model/diffusion.py:226-286), the shipped safety guidance (1D/utils/guidance.py:58-86, closed-form
 * gradient), the DDIM update (diffusion.py:500-521) or the DDPM p_mean_variance/p_sample update (:288-306),
 * and the condition / pad writes (:336-366).  State layout [B,3,H,W] fp32 NCHW, W a multiple of 4.
 */
typedef struct {
    float c1;          /* sqrt_recip_alphas_cumprod[t] */
    float c2;          /* sqrt_recipm1_alphas_cumprod[t] */
    float k_x0;        /* DDIM: sqrt(alpha_next)            DDPM: posterior_mean_coef1[t] */
    float k_eps;       /* DDIM: sqrt(1-alpha_next-sigma^2)  DDPM: posterior_mean_coef2[t] (multiplies x_t) */
    float k_noise;     /* DDIM: sigma                       DDPM: exp(0.5*posterior_log_variance_clipped[t]) */
    float sched;       /* J_scheduler(t) multiplier on the guidance gradient (1 if none) */
    int32_t is_last;   /* DDIM: time_next < 0 -> output = x0, no condition writes.  DDPM: t == 0 -> no noise */
    int32_t t;         /* diffusion time of this step (feeds the in-kernel RNG counter) */
} sdc_step_coef;

typedef struct {
    int32_t mode;          /* 0 = none, 1 = safety mean (use_max_safety=True), 2 = safety amax, 3 = gradient supplied */
    float Q;               /* conformal quantile added to the score */
    float u_bound_sq;      /* fp32(u_bound**2) */
    float w_score;         /* guidance weight */
    float scaler;          /* SCALER (10) */
    int32_t nt;            /* valid rows of the safety channel (11) */
} sdc_guidance;

#define SDC_SAMPLER_DDIM 0
#define SDC_SAMPLER_DDPM 1
/* DDPM second p_sample of the guidance_u0=False path: eps supplied is used as is (no guidance) */

/* x:[B,3,H,W] current x_t (read), eps: model output (read), noise: pre-generated N(0,1) or NULL -> Philox
 * (seed, global sample index = sample_offset+b, t), out: x_{t-1} (may alias x).  coef: DEVICE pointer to a table
 * of sdc_step_coef; the step used is step_counter ? *step_counter : step (device counter => one captured graph
 * serves every step).  u_init/u_final:[B,W] or NULL, w_gt:[B,H,W] or NULL, grad:[B,3,H,W] only for mode 3.
 * x0_out / eps_out: optional [B,3,H,W] outputs (pred_x_start, pred_noise).  cond_idx = index of the uT row (10);
 * pad_writes != 0 applies set_pad_condition. */
int sdc_reverse_step(int sampler, const float* x, const float* eps, const float* noise, float* out,
                     float* x0_out, float* eps_out, const sdc_step_coef* coef, int step, const int32_t* step_counter,
                     const sdc_guidance* guidance /* host */, const float* grad,
                     const float* u_init, const float* u_final, const float* w_gt, int cond_idx, int pad_writes,
                     int clip_denoised, uint64_t seed, int64_t sample_offset, int64_t B, int H, int W, void* stream);

/* set_condition + set_pad_condition on x in place (diffusion.py:336-366); used once before the first step. */
int sdc_write_conditions(float* x, const float* u_init, const float* u_final, const float* w_gt, int cond_idx,
                         int pad_writes, int64_t B, int H, int W, void* stream);

/* x <- N(0,1) from the library's Philox stream (replaces torch.randn(shape), diffusion.py:375,464). */
int sdc_fill_normal(float* x, int64_t B, int64_t per_sample, uint64_t seed, int64_t sample_offset, int32_t t_tag,
                    void* stream);

/* *counter += 1 on the stream (advances the captured-graph step). */
int sdc_advance_counter(int32_t* counter, void* stream);

/* Device-resident state of one reverse chain.  A whole reverse step (denoiser + sdc_reverse_step_state +
 * sdc_chain_state_advance) captured ONCE in a CUDA graph then serves every step of every chain of that shape: the
 * step index, the Philox seed and the global sample offset are read from device memory instead of kernel arguments
 * (the reference's python loop, 1D/model/diffusion.py:380-449,470-523, re-issues ~600 launches per step). */
typedef struct {
    int32_t step;          /* row of the sdc_step_coef table used by the next reverse step */
    int32_t reserved;
    uint64_t seed;         /* Philox key */
    int64_t sample_offset; /* global index of sample 0 of this shard */
} sdc_chain_state;

/* state (device) <- (step, seed, sample_offset); t_index[0..B) (device int32, may be NULL) <- coef[step].t, the
 * per-sample diffusion-time index the denoiser's FiLM lookup reads. */
int sdc_chain_state_set(sdc_chain_state* state, int32_t step, uint64_t seed, int64_t sample_offset,
                        const sdc_step_coef* coef, int n_steps, int32_t* t_index, int64_t B, void* stream);
/* state->step += 1; t_index[0..B) <- coef[min(step, n_steps-1)].t */
int sdc_chain_state_advance(sdc_chain_state* state, const sdc_step_coef* coef, int n_steps, int32_t* t_index, int64_t B,
                            void* stream);
/* sdc_reverse_step with (step, seed, sample_offset) taken from *state on the device. */
int sdc_reverse_step_state(int sampler, const float* x, const float* eps, const float* noise, float* out,
                           float* x0_out, float* eps_out, const sdc_step_coef* coef, const sdc_chain_state* state,
                           const sdc_guidance* guidance /* host */, const float* grad,
                           const float* u_init, const float* u_final, const float* w_gt, int cond_idx, int pad_writes,
                           int clip_denoised, int64_t B, int H, int W, void* stream);
/* adds n to sdc_launch_count(): kernels replayed from a captured graph are accounted by the host layer */
void sdc_count_launches(int64_t n);

/* ------------------------------------------------------------------------------------------------
 * Synthetic data (the step BEFORE the path).  Replaces the float64 numpy field evaluation of make_data_varying_f
 * (1D/data/generate_burgers.py:338-418) and the per-item tensor assembly of BurgersDataset._process_data
 * (1D/data/burgers.py:104-142).  The O(N) random scalars are drawn on the host from numpy's RNG in the reference's order.
 */
/* params_u0:[Nu0,6] = (loc1, amp1, sig1, loc2, amp2, sig2); params_f:[Nf,terms,5] = (amp, loc_x, sig_x, loc_t, sig_t), all
 * DEVICE float64; x_grid:[s], t_grid:[t] device float32 (torch.linspace values).  Outputs: u0 as float64 and/or float32
 * [Nu0,s] (either may be NULL), f:[Nf,t,s] float32 = float32(sum of terms) (* alpha, clamped to +-10 when alpha != 1).
 * partial_control: 0 = none, 1 = 'front_rear_quarter' mask (the caller doubles amp_compensate as the reference does). */
int sdc_burgers_fields(const double* params_u0, const double* params_f, const float* x_grid, const float* t_grid,
                       double* u0_f64, float* u0_f32, float* f, int64_t Nu0, int64_t Nf, int s, int t, int terms,
                       double amp_compensate, int partial_control, float alpha, void* stream);
/* state:[N,3,pad,s] = (u, f, safety)/scaler from u_traj:[N,nt1,s], f:[N,nt,s]; safety = u^2, or its per-sample maximum
 * broadcast when use_max_safety; rows beyond nt1 / nt are zero. */
int sdc_dataset_states(const float* u_traj, const float* f, float* state, int64_t N, int nt1, int nt, int pad, int s,
                       float scaler, int use_max_safety, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conformal calibration.  Replaces calculate_guidance/get_weight (1D/inference/guidance.py:9-46), the
 * nonconformity score of ConformalCalculator.get_conformal_scores (1D/inference/conformal.py:74-85),
 * normalize_weights (guidance.py:48-66) and calculate_quantile (conformal.py:95-118).
 */
/* stat[b] = red(scaler * x[b,2,:nt,:]) with red = mean (use_mean != 0) or amax.  x:[B,3,H,W]. */
int sdc_safety_stat(const float* x, float* stat, int use_mean, float scaler, int nt, int64_t B, int H, int W,
                    void* stream);
/* score[b] = |stat(pred) - stat(state)| ; weight[b] = exp(-w_score*max(stat(state)+Q-u_bound_sq,0)) and, if
 * Q2 is finite, multiplied by the same expression at Q2 (InfFT_Q).  pred,state:[B,3,H,W] in model units. */
int sdc_conformal_scores(const float* pred, const float* state, float* score, float* weight, const sdc_guidance* g
                         /* host; mode 1 or 2 */, float Q2, int64_t B, int H, int W, void* stream);
/* In place on w[n]: inf -> largest finite; out[i] = n*w[i]/sum(w), or 1 when the sum is 0; if scores != NULL,
 * scores[i] *= out[i] (the weighted scores the reference returns).  Single-CTA deterministic reduction. */
int sdc_normalize_weights(float* w, float* out, float* scores, int64_t n, void* stream);
/* rank-th order statistic (0-based, ascending; NaNs sort last like torch.sort) of scores[n] by radix select on the
 * fp32 bit patterns: *value_out = that value (bit-exact), *index_out = the index a STABLE sort would pick
 * (lowest index first among equal values).  workspace: >= sdc_kth_select_workspace(n) bytes of device memory. */
int64_t sdc_kth_select_workspace(int64_t n);
int sdc_kth_select(const float* scores, int64_t n, int64_t rank, float* value_out, int64_t* index_out,
                   void* workspace, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Denoiser.  Replaces Unet2D.forward (1D/model/unet.py:263-426) for the 2-D (time x space) U-Net with
 * channels-in = 3, groups = 1, 4 heads x 32.  Convolutions run as implicit GEMMs on tcgen05 (TF32 operands
 * rounded to nearest, FP32 accumulation in TMEM) fed by TMA; GroupNorm/SiLU/attention/LayerNorm are fused
 * bandwidth kernels.  See include/safediffcon_b200_unet.h for the handle API.
 */

#ifdef __cplusplus
}
#endif
#endif
