/* safediffcon_b200 -- C ABI of the WHOLE denoiser: one entry point per network pass.
 *
 * Replaces Unet2D.forward (/root/reference/1D/model/unet.py:382-426) for inference: eps = Unet2D(x_t, t).  The handle owns
 * the packed tensor-core weights, the per-timestep FiLM table and the launch schedule (~120 kernels of
 * safediffcon_b200_unet.h per evaluation: tcgen05 convolutions with fused GroupNorm / attention epilogues), so a caller
 * without Python -- or Python through ONE ctypes call instead of ~150 -- can evaluate the network.  Conventions as in
 * safediffcon_b200.h (device pointers, int status, `stream` = cudaStream_t as void*).
 *
 *   sdc_unet* net;
 *   sdc_unet_create(&net, 128, (int[]){1,2,4,8}, 4, 3, 3, SDC_PREC_F16, 10000.f, 1000);
 *   for (i < sdc_unet_param_count(net)) ptr[i] = <device fp32 tensor named sdc_unet_param_name(net, i)>;   // state_dict keys
 *   sdc_unet_pack_weights(net, ptr, n, stream);                     // again after every optimiser / EMA update
 *   ws = cudaMalloc(sdc_unet_workspace_bytes(net, B, 16, 128));
 *   sdc_unet_forward(net, x, NULL, t, eps, B, 16, 128, ws, bytes, stream);
 *
 * sdc_unet_forward allocates nothing, never synchronises and is CUDA-graph capturable at any batch size: every activation
 * lives in the caller's workspace (deterministic first-fit layout per (B, H, W)).
 */
#ifndef SAFEDIFFCON_B200_PLAN_H
#define SAFEDIFFCON_B200_PLAN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdc_unet sdc_unet;

/* Architecture of Unet2D(dim, dim_mults, channels, out_dim, resnet_block_groups = 1, 4 heads x 32) (unet.py:268-380).
 * prec: SDC_PREC_F16 (dim % 64 == 0; fp16 operands and compact fp16 intermediates) or SDC_PREC_TF32 (dim % 32 == 0).
 * table_timesteps: integer diffusion times [0, table_timesteps) get their FiLM rows from a table built by
 * sdc_unet_pack_weights (the time MLP is batch independent, unet.py:310-315,152-155).  Touches no CUDA state. */
int sdc_unet_create(sdc_unet** out, int dim, const int* dim_mults, int n_mults, int channels, int out_dim, int prec,
                    float sinusoidal_theta, int table_timesteps);
void sdc_unet_destroy(sdc_unet* net);

/* The parameters the handle expects, in the order of Unet2D.named_parameters(); names are the reference's state_dict keys
 * ("init_conv.weight", "downs.0.0.block1.proj.weight", ...), shapes the reference's (OIHW conv weights). */
int sdc_unet_param_count(const sdc_unet* net);
const char* sdc_unet_param_name(const sdc_unet* net, int i);
int64_t sdc_unet_param_numel(const sdc_unet* net, int i);

/* Pack / repack: params[i] = device pointer to the contiguous fp32 tensor of parameter i (never written).  Rounds the conv
 * weights to the operand precision in the K-major tcgen05 layout, copies the small parameters, rebuilds the FiLM table.  The
 * first call allocates the handle's device storage (cudaMalloc: ~0.6 GB for dim 128); later calls refresh it IN PLACE, so CUDA
 * graphs captured around sdc_unet_forward stay valid across weight updates. */
int sdc_unet_pack_weights(sdc_unet* net, const float* const* params, int n_params, void* stream);

/* Bytes of workspace sdc_unet_forward needs for this problem size (0 on error). */
int64_t sdc_unet_workspace_bytes(const sdc_unet* net, int B, int H, int W);

/* eps[B, out_dim, H, W] = Unet2D(x[B, channels, H, W], t), NCHW fp32 in and out.  Diffusion times: t_index != NULL -> device
 * int32[B] per-sample integer times (read by the kernels: a captured graph follows in-place updates); else the batch-uniform
 * integer t_uniform.  Times must lie in [0, table_timesteps).  H % 8 == 0, W % 8 == 0, W | 128 (see sdc_conv_gemm).
 * nonfinite (optional): device uint32 counter, += number of non-finite entries of eps (FP16-range guard). */
int sdc_unet_forward(sdc_unet* net, const float* x, const int32_t* t_index, int t_uniform, float* eps, int B, int H, int W,
                     void* workspace, int64_t workspace_bytes, uint32_t* nonfinite, void* stream);

/* Backward-data pass: eps = Unet2D(x, t) AND grad_x = d<eps, grad_eps>/dx in one call (the VJP with respect to the denoiser input
 * that the reference obtains by autograd when a guidance callable differentiates eps_theta(x_t, t),
 * /root/reference/1D/model/diffusion.py:254-262; parameters get no gradient).  Forward with every normalisation input kept, then
 * the reverse walk: GroupNorm / LayerNorm / attention backward kernels and the same tcgen05 convolutions in TF32 with the
 * transposed, tap-flipped weights.  Needs SDC_UNET_BACKWARD set before the last sdc_unet_pack_weights.  x, grad_eps, eps, grad_x:
 * [B, C, H, W] fp32 NCHW.  Workspace: sdc_unet_backward_workspace_bytes (~42 MB per sample for dim 128: all records stay live
 * until the reverse walk consumes them).  Allocates nothing, never synchronises. */
int64_t sdc_unet_backward_workspace_bytes(const sdc_unet* net, int B, int H, int W);
int sdc_unet_backward_data(sdc_unet* net, const float* x, const int32_t* t_index, int t_uniform, const float* grad_eps, float* eps,
                           float* grad_x, int B, int H, int W, void* workspace, int64_t workspace_bytes, void* stream);

/* The FiLM table built by sdc_unet_pack_weights: table[t * cols + c], t < rows = table_timesteps, cols = sum of 2 * Cout over the
 * ResnetBlocks in execution order (scale | shift per block; unet.py:152-155,166-175).  rows / cols receive the shape; if `out`
 * (device, rows * cols floats) is not NULL the table is copied into it on `stream`. */
int sdc_unet_film_table(const sdc_unet* net, float* out, int* rows, int* cols, void* stream);

/* Schedule switches, both EXPERIMENTAL and off by default (correct -- tests -- but measured slower than the separate kernels on
 * B200).  SDC_UNET_FUSE_LN: FP16 mode, levels with <= 256 channels -- the PreNorm LayerNorm is folded into the
 * qkv projection and the output LayerNorm + residual into the per-sample projection (sdc_conv1x1_qkv_ln,
 * sdc_conv1x1_per_sample_ln); value 2 folds the PreNorm only (output LayerNorm as a separate kernel).  SDC_UNET_FUSE_GN: conv + GroupNorm in one kernel on the 16x128 level (sdc_conv3x3_row_gn). */
#define SDC_UNET_FUSE_LN 1
#define SDC_UNET_FUSE_GN 2
/* SDC_UNET_FILM_TC (default 1): build the FiLM table with ONE tcgen05 TF32 GEMM over operands split into high + low parts
 * (relative error ~1e-6, 0.2 ms) instead of the fp32 CUDA-core loop (9 ms for dim 128); applies to the next sdc_unet_pack_weights. */
#define SDC_UNET_FILM_TC 3
/* SDC_UNET_BACKWARD (default 0): sdc_unet_pack_weights also packs the data-gradient weights (transposed, tap-flipped, TF32; a
 * second device allocation of the size of the fp32 conv weights) that sdc_unet_backward_data needs. */
#define SDC_UNET_BACKWARD 4
int sdc_unet_set_flag(sdc_unet* net, int flag, int value);

/* Per-launch profile of the NEXT forward calls: when enabled, every launch is bracketed by CUDA events on `stream` (adds two
 * event records per launch; do not enable inside a graph capture).  After synchronising, read entry i: kernel family name,
 * milliseconds, algorithmic bytes moved (HBM roofline) and FLOPs (tensor roofline) of that launch.  bench.py builds its
 * roofline objects from these. */
int sdc_unet_profile_enable(sdc_unet* net, int enable);
int sdc_unet_profile_count(const sdc_unet* net);
int sdc_unet_profile_entry(const sdc_unet* net, int i, const char** name, float* ms, double* bytes, double* flops);

#ifdef __cplusplus
}
#endif
#endif
