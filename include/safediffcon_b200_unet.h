/* safediffcon_b200 -- C ABI of the denoiser (Unet2D) building blocks.  See safediffcon_b200.h for conventions.
 *
 * Activation layout: NHWC ([B*H*W, C] row-major, "pixel rows").  Tensors that feed a tensor-core convolution
 * ("operands") are stored in the operand precision selected by `prec`, rounded to nearest by the producing kernel:
 *   SDC_PREC_TF32: fp32 containers holding TF32 values (10-bit mantissa), tcgen05.mma.kind::tf32;
 *   SDC_PREC_F16 : IEEE fp16 (same 10-bit mantissa, half the bytes, twice the tensor rate), tcgen05.mma.kind::f16.
 * Accumulation is always FP32; tensors that feed normalisation statistics or softmax (conv outputs ahead of
 * GroupNorm, qkv) stay fp32.  In the signatures below `const void*` operands are `float*` (TF32) or `__half*` (F16).
 * Weights are packed once per parameter version by sdc_pack_conv_weight.  The ops are exposed individually (each
 * is parity-tested against its torch counterpart); safediffcon_b200/unet.py strings them into Unet2D.forward
 * (reference: 1D/model/unet.py:382-426).
 */
#ifndef SAFEDIFFCON_B200_UNET_H
#define SAFEDIFFCON_B200_UNET_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SDC_PREC_TF32 0
#define SDC_PREC_F16 1

/* Repack an OIHW fp32 conv weight [Cout, Cin, kh, kw] into the K-major GEMM operand Wp[Cout, taps*Cin]
 * (K index = tap*Cin + cin, tap = ky*kw_ + kx), rounded to the operand precision.
 * kind 0: 1x1, 1: 3x3, 2: the 1x1 conv that follows the pixel-unshuffle of Downsample2d (unet.py:39-43):
 * Cin = 4*C with channel index c*4 + p1*2 + p2  ->  K index = (p1*2+p2)*C + c.
 * kind 3: the 3x3 conv that follows the nearest x2 upsample of Upsample2d (unet.py:33-37), packed for the fused
 * upsample convolution: Wp[4*Cout, 4*Cin], row = phase*Cout + co (phase = 2a + b = parity of the output pixel), K index =
 * (2r + s)*Cin + ci over the 2x2 low-resolution window; each entry is the sum of the 3x3 taps that read that input pixel. */
int sdc_pack_conv_weight(int prec, int kind, const float* w_oihw, void* w_packed, int Cout, int Cin, void* stream);

/* Implicit-GEMM convolution on tcgen05 (replaces nn.Conv2d 3x3 pad 1 / 1x1, unet.py:132,161,189-192,232-233,345,370,
 * and Downsample2d, unet.py:39-43).  Inputs: one or two NHWC operand tensors (a1 = second K segment of a channel
 * concat, c1 = 0 if absent) of spatial size H x W (kind 2: 2H x 2W).
 * out[B*H*W, Cout] = conv + bias (+ residual[B*H*W, Cout], an operand-precision tensor).
 * stats (optional): double[B][2], += (sum, sum of squares) of the fp32 results per sample (GroupNorm(1, C)).
 * operand_out: 1 = store `out` in the operand precision (it feeds another convolution), 0 = plain fp32.
 * kind 3 (nearest-upsample x2 followed by 3x3 pad 1, Upsample2d): H x W is the INPUT size, out is [B*2H*2W, Cout]; the
 * upsampled tensor is never materialised -- output pixel (2i+a, 2j+b) is a 2x2 convolution of the input with the phase-(a, b)
 * weights of sdc_pack_conv_weight(kind 3), i.e. 4/9 of the multiply-adds; single input, no residual / stats, W = 16 or W % 32 == 0.
 * Requirements: W | 128, (H*W) % 32 == 0, input channels % 32 (TF32) / % 64 (F16) == 0, Cout % 32 == 0. */
int sdc_conv_gemm(int prec, int kind, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                  const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout,
                  void* stream);

/* Same contract as sdc_conv_gemm(kind = 3x3) for the full-resolution level: W == 128 and Cout <= 128.  One CTA computes
 * two image rows and loads the activation halo once per 128-byte channel chunk (the 9 taps are shifted UMMA descriptor
 * views of it), cutting L2->SMEM operand traffic ~2.8x.  W == 64 (H % 8 == 0, at least one 8-row item per SM pair): a CTA pair
 * computes eight image rows from one 6-row x 64-pixel activation box per channel chunk and HORIZONTAL tap (TMA does that
 * shift), the vertical taps being row-shifted descriptor views of the box (csrc/conv_row64.cu): half the operand traffic of
 * the generic kernel.  Returns -1 (and does nothing) when the shape is not eligible. */
int sdc_conv3x3_row(int prec, const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias,
                    const void* residual, void* out, double* stats, int operand_out, int B, int H, int W, int Cout,
                    void* stream);

/* sdc_conv3x3_row (FP16 mode, fp16 output) FUSED with the GroupNorm(1, C) + FiLM + SiLU (+ residual) that follows it in
 * Block.forward / ResnetBlock.forward (unet.py:138-147,177-180) -- the separate sdc_gn_silu pass over HBM disappears.
 * Deferred epilogue on the CTA-pair kernel: work items are dealt round-robin, an epilogue warp first reduces its accumulators to
 * a partial (sum, sum of squares) of conv + bias and publishes it in its own 8-byte slot of the sample, waits until all 128 slots
 * of the sample are filled while the tensor pipe computes the next item, then normalises the fp32 accumulators straight from
 * TMEM and stores fp16.  sync_slots: B * SDC_GN_SLOT_BYTES bytes, EVERY BYTE 0xFF on entry (cudaMemsetAsync); stats: double[B][2],
 * receives the per-sample totals (fixed summation order: deterministic); gamma .. gn_residual as in sdc_gn_silu (gn_residual: fp16
 * or NULL).  Returns -1 (nothing done) when the shape is not eligible (needs FP16, H % 4 == 0, H <= 16, W == 128, Cout <= 128,
 * Cout % 32 == 0); the caller then issues sdc_conv3x3_row / sdc_conv_gemm followed by sdc_gn_silu. */
#define SDC_GN_SLOT_BYTES 1024
int sdc_conv3x3_row_gn(const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias, void* out,
                       double* stats, void* sync_slots, const float* gamma, const float* beta, const float* scale_shift,
                       const int32_t* t_index, int64_t ss_stride, const void* gn_residual, int B, int H, int W, int Cout, void* stream);

/* The network's last convolution fused with everything behind it (final_res_block.block2 + residual + final_conv,
 * unet.py:138-147,178-180,378,426): out[b, o, p] = sum_c head_w[o, c] * (silu(GN(conv)[b, p, c]) + gn_residual[b, p, c]) + head_b[o],
 * NCHW fp32.  Same kernel as sdc_conv3x3_row_gn; in pass 2 a lane owns a pixel, so the 1x1 head convolution is a per-lane dot
 * product over the normalised accumulator columns -- the activation is never stored or rounded.  head_cout <= 4.  out_nchw must
 * be ZEROED by the caller: the two warps that own the channel halves of a pixel add their partial dot products to it. */
int sdc_conv3x3_row_gn_head(const void* a0, int c0, const void* a1, int c1, const void* w_packed, const float* bias, double* stats,
                            void* sync_slots, const float* gamma, const float* beta, const void* gn_residual, const float* head_w,
                            const float* head_b, float* out_nchw, int head_cout, int B, int H, int W, int Cout, void* stream);

/* Stem: 7x7 pad 3 convolution of the NCHW model input (unet.py:326,392).  x:[B,Cin,H,W] NCHW fp32, w:[Cout,Cin,7,7] OIHW
 * (unpacked, fp32), out: NHWC operand [B*H*W, Cout].  FP32 CUDA-core kernel (0.3% of the FLOPs). */
int sdc_stem_conv7(int prec, const float* x, const float* w, const float* bias, void* out, int B, int Cin, int H, int W,
                   int Cout, void* stream);

/* Tensor-core stem, step 1: im2col of the NCHW fp32 input for the 7x7 pad-3 convolution.  a:[B*H*W, kp] operand rows,
 * columns [0, Cin*49) = high part of the patch (index ci*49 + ky*7 + kx), [kp/2, kp/2 + Cin*49) = low part
 * (x - high), zeros elsewhere.  Step 2 is sdc_conv_gemm(kind 0, c0 = kp) with the stem weight repeated in both ranges:
 * the product sees x to ~2^-22 with fp16 / tf32 operands. */
int sdc_stem_im2col(int prec, const float* x, void* a, int B, int Cin, int H, int W, int kp, void* stream);

/* The whole stem in ONE tcgen05 kernel (FP16 mode, W = 128, Cout = 128, Cin <= 3).  Per image row the kernel's own warps write ONE
 * operand tile T[w' = 0..133][(ci, ky)] = x[ci, h + ky - 3, w' - 3] (high | low fp16 split as in sdc_stem_im2col, zero = padding) into
 * shared memory in the K-major SWIZZLE_128B layout; the seven horizontal taps are row-shifted UMMA views of it, multiplied against
 * seven resident [Cout x 64] weight tiles -- no patch matrix in HBM (round 1: 1.3 GB written and read back per evaluation at
 * B = 1024) and none in shared memory either.  w_packed: sdc_pack_conv_weight(kind 0) of the [Cout, 320] matrix holding the
 * stem weight at columns 0 and 160.  out: NHWC fp16 [B*H*W, Cout] (+ bias).  Returns -1 (nothing done) for other shapes. */
int sdc_stem_conv7_tc(const float* x, const void* w_packed, const float* bias, void* out, int B, int Cin, int H, int W, int Cout, int kp,
                      void* stream);

/* GroupNorm(1, C) apply + FiLM + SiLU (+ residual), Block.forward (unet.py:138-147) and ResnetBlock's sum (:180):
 * y = silu(((x - mean_b) * rstd_b * gamma_c + beta_c) * (scale_bc + 1) + shift_bc) + res ; mean/rstd from
 * stats[b] = (sum, sumsq) over C*HW elements, eps 1e-5, biased variance.  scale_shift: [n_t, 2C] rows
 * (scale | shift) indexed by t_index[b] (NULL t_index = row 0 for every sample), or NULL for none.
 * x: a conv output -- fp32, or fp16 if x_operand (FP16 mode, compact intermediates: the statistics were taken from the fp32
 * accumulators, only the stored activations are rounded; the residual must then be fp16 too).
 * residual: operand precision if residual_operand else fp32.  y: operand. */
int sdc_gn_silu(int prec, const void* x, int x_operand, const double* stats, const float* gamma, const float* beta, const float* scale_shift,
                const int32_t* t_index, int64_t ss_stride, const void* residual, int residual_operand, void* y, int B, int HW,
                int C, void* stream);

/* The network's LAST GroupNorm + SiLU + residual fused with the 1x1 head convolution (final_res_block.block2 + final_conv,
 * unet.py:178-180,378,426): out[b, o, p] = sum_c head_w[o, c] * (silu(GN(x)[b, p, c]) + residual[b, p, c]) + head_b[o], NCHW fp32.
 * x: fp32 [B*HW, C] (C = 128), stats as in sdc_gn_silu, residual fp16 if residual_operand else fp32, Cout <= 4.  The activation
 * stays fp32 in registers (never stored, never rounded). */
int sdc_gn_silu_head(const float* x, const double* stats, const float* gamma, const float* beta, const void* residual,
                     int residual_operand, const float* head_w, const float* head_b, float* out, int B, int HW, int C, int Cout,
                     void* stream);

/* Channel LayerNorm (unet.py:53-63): y = (x - mean_c) * rsqrt(var_c + 1e-5) * g (+ residual), per pixel row.
 * x: operand precision if x_operand else fp32; residual and y: operand precision (TF32 mode: y is rounded to TF32
 * only when operand_out is set; F16 mode: y is always fp16). */
int sdc_channel_layernorm(int prec, const void* x, int x_operand, const float* g, const void* residual, void* y, int64_t M,
                          int C, int operand_out, void* stream);

/* LinearAttention core (unet.py:202-222) on fp32 qkv:[B*n, 384] rows (q | k | v, each heads*32 channels):
 * q <- softmax_d(q) * 32^-0.5 ; k <- softmax_n(k) ; ctx = k v^T ; out[B*n, 128] = ctx^T q, an operand.
 * n (pixels per sample) must be a multiple of 32.  workspace: >= sdc_linear_attention_workspace(B) bytes; on return it
 * holds per (sample, head) ctx[32][32] | kmax[32] | ksum[32], which sdc_linear_attention_bwd consumes. */
int64_t sdc_linear_attention_workspace(int B);
int sdc_linear_attention(int prec, const float* qkv, void* out, void* workspace, int B, int n, void* stream);

/* Fused LinearAttention path (same math as sdc_conv_gemm(qkv) + sdc_linear_attention + sdc_conv_gemm(to_out), without ever
 * materialising the [B*n, 128] attention tensor):
 *   1. sdc_conv1x1_qkv: the bias-free qkv projection (unet.py:189,203) whose epilogue applies q <- softmax_d(q) * 32^-0.5 to
 *      the q columns in registers; writes qs[B*H*W, hidden] (operand precision) and kv[B*H*W, 2*hidden] (k | v; fp32, or fp16 when
 *      kv_operand != 0, FP16 mode only -- pass the same flag to step 2).
 *   2. sdc_linear_attention_context: ctx = softmax_n(k) v^T per (sample, head) from rows k + i*ld, v + i*ld into the workspace.
 *   3. sdc_linear_attention_fold: per-sample folded projection Wf_b[Cout, 128] = W_out (x) ctx_b (operand precision).
 *   4. sdc_conv1x1_per_sample: out[B*H*W, Cout] = qs * Wf_b^T + bias (H*W must be a multiple of 128). */
int sdc_conv1x1_qkv(int prec, const void* a, int c, const void* w_packed, void* q_out, void* kv_out, int kv_operand, int B, int H, int W,
                    int hidden, void* stream);
int sdc_linear_attention_context(const void* k, const void* v, int ld, int kv_operand, void* workspace, int B, int n, void* stream);
int sdc_linear_attention_fold(int prec, const void* workspace, const float* w_out, void* w_folded, int B, int Cout, void* stream);
int sdc_conv1x1_per_sample(int prec, const void* a, int c, const void* w_per_sample, const float* bias, void* out, int operand_out,
                           int B, int H, int W, int Cout, void* stream);

/* PreNorm LayerNorm FOLDED into the qkv projection (FP16 mode; unet.py:65-76 + 189,203): W LN(x) = r (W_g x - mu W_g 1) with
 * W_g = W diag(g).  Three pieces:
 *   sdc_gn_silu_rowstats  = sdc_gn_silu(FP16, fp16 x in place, fp16 residual) that also emits rowstats[M][2] = per pixel row
 *                           (channel mean, 1 / sqrt(var + 1e-5)) of its fp16 OUTPUT (C = 128 or 256);
 *   sdc_pack_qkv_ln       : w_folded[Cout, Cin] = fp16(w * g), wsum[co] = sum_c float(w_folded[co, c]);
 *   sdc_conv1x1_qkv_ln    = sdc_conv1x1_qkv on the RAW rows with the folded weights; the epilogue applies r (acc - mu wsum).
 * The separate LayerNorm pass (one read + one write of the activations) disappears and the normalised rows are never rounded. */
int sdc_gn_silu_rowstats(const void* x, const double* stats, const float* gamma, const float* beta, const float* scale_shift,
                         const int32_t* t_index, int64_t ss_stride, const void* residual, void* y, float* rowstats, int B, int HW, int C,
                         void* stream);
int sdc_pack_qkv_ln(const float* w, const float* g, void* w_folded, float* wsum, int Cout, int Cin, void* stream);
int sdc_conv1x1_qkv_ln(const void* a, int c, const void* w_folded, const float* wsum, const float* rowstats, void* q_out, void* kv_out,
                       int B, int H, int W, int hidden, void* stream);

/* sdc_conv1x1_per_sample (FP16 mode) with LinearAttention's output LayerNorm and the Residual add in its epilogue
 * (to_out = Conv2d -> LayerNorm, unet.py:190-193, then Residual, :16-22): out = LN(a Wf_b^T + bias) * gain + residual, fp16.
 * A lane owns an output row in TMEM, so the channel statistics are a per-lane reduction over the accumulator columns (two TMEM
 * passes, no exchange).  Cout = 128 or 256 (the row must sit in one N tile). */
int sdc_conv1x1_per_sample_ln(const void* a, int c, const void* w_per_sample, const float* bias, const float* gain, const void* residual,
                              void* out, int B, int H, int W, int Cout, void* stream);

/* Full softmax attention core (unet.py:239-258) for n <= 32 tokens: out[B*n, 128], an operand. */
int sdc_attention(int prec, const float* qkv, void* out, int B, int n, void* stream);

/* Nearest-neighbour x2 upsample of an NHWC operand tensor (nn.Upsample(scale_factor=2), unet.py:33-37). */
int sdc_upsample2x(int prec, const void* x, void* y, int B, int H, int W, int C, void* stream);

/* Final 1x1 conv to the model's output channels, NHWC operand -> NCHW fp32 (unet.py:378,426).  w:[Cout, Cin] fp32. */
int sdc_head_conv1(int prec, const void* x, const float* w, const float* bias, float* out, int B, int HW, int Cin, int Cout,
                   void* stream);

/* Rows of a small dense layer: y[r, :] = act_in(x[r, :]) @ w[N, K]^T + b, act_in 0 = identity, 1 = SiLU, 2 = GELU(erf)
 * (time MLP, unet.py:310-315 and ResnetBlock.mlp, :152-155).  Batch independent: evaluated once per distinct t. */
int sdc_linear_rows(const float* x, const float* w, const float* b, float* y, int R, int K, int N, int act_in, void* stream);

/* Sinusoidal embedding (unet.py:81-95): emb[r] = (sin(t_r * f_k), cos(t_r * f_k)), f_k = exp(-k ln(theta)/(dim/2-1)). */
int sdc_sinusoidal_embedding(const float* t, float* emb, int R, int dim, float theta, void* stream);

/* Zero a double buffer (GroupNorm statistics) on the stream. */
int sdc_zero_f64(double* p, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Backward-data (VJP with respect to the denoiser input x_t; reference: autograd through Unet2D.forward when a
 * guidance callable differentiates eps_theta(x_t, t), 1D/model/diffusion.py:254-262).  Parameter gradients are not
 * produced.  Gradients are fp32 tensors; "operand" gradients (inputs of a dgrad convolution) are TF32-rounded.  The
 * dgrad of every convolution is sdc_conv_gemm / sdc_conv3x3_row in SDC_PREC_TF32 with weights packed by
 * sdc_pack_conv_weight_dgrad.
 * ------------------------------------------------------------------------------------------------------------- */

/* Pack an OIHW conv weight for the data-gradient convolution: Wt[Cin, taps*Cout] (TF32-rounded fp32),
 * kind 1: Wt[ci, tap'*Cout + co] = W[co, ci, 8 - tap'] (flipped taps); kind 0: Wt = W^T;
 * kind 2: Wt[p*C + c, co] = W[co, c*4 + p] (row order of the pixel shuffle that follows the 1x1 dgrad). */
int sdc_pack_conv_weight_dgrad(int kind, const float* w_oihw, float* w_packed, int Cout, int Cin, void* stream);

/* Backward of sdc_gn_silu w.r.t. x: dx = d/dx silu(GN(x) FiLM), TF32-rounded.  dy: gradient w.r.t. the activation
 * output (the residual branch is handled by the caller), x / stats / gamma / beta / scale_shift / t_index as in the
 * forward call.  sums: double[B][2] scratch (zeroed and filled here). */
int sdc_gn_silu_bwd(const float* dy, const float* x, const double* stats, const float* gamma, const float* beta,
                    const float* scale_shift, const int32_t* t_index, int64_t ss_stride, double* sums, float* dx, int B, int HW,
                    int C, void* stream);

/* Backward of the channel LayerNorm: dx = LN'(x)^T (dy * g) (+ add).  x: the forward input (fp16 if x_half else fp32). */
int sdc_channel_layernorm_bwd(const float* dy, const void* x, int x_half, const float* g, const float* add, float* dx, int64_t M,
                              int C, int operand_out, void* stream);

/* Backward of sdc_linear_attention: dqkv[B*n, 384] (TF32-rounded) from dout[B*n, 128], the forward qkv and the
 * forward workspace (ctx | kmax | ksum per head).  workspace: >= sdc_linear_attention_bwd_workspace(B) bytes. */
int64_t sdc_linear_attention_bwd_workspace(int B);
int sdc_linear_attention_bwd(const float* qkv, const float* dout, const void* fwd_workspace, void* workspace, float* dqkv, int B,
                             int n, void* stream);

/* Backward of sdc_attention (n <= 32 tokens): dqkv[B*n, 384], TF32-rounded. */
int sdc_attention_bwd(const float* qkv, const float* dout, float* dqkv, int B, int n, void* stream);

/* Backward of the pixel-unshuffle view: dx[b, 2h+p1, 2w+p2, c] = t[b, h, w, (2 p1 + p2) C + c] (+ add); H, W = low-res size. */
int sdc_pixel_shuffle_bwd(const float* t, const float* add, float* dx, int B, int H, int W, int C, int operand_out, void* stream);

/* Backward of sdc_upsample2x: dx[b, h, w, c] = sum of the 2x2 block of dy; H, W = low-res size. */
int sdc_upsample2x_bwd(const float* dy, float* dx, int B, int H, int W, int C, int operand_out, void* stream);

/* a += b over n floats (n % 4 == 0), optionally rounding the sum to TF32. */
int sdc_add_inplace(float* a, const float* b, int64_t n, int operand_out, void* stream);

/* Backward of sdc_head_conv1: dx[B*HW, Cin] = g[B, Cout, HW]^T w[Cout, Cin]. */
int sdc_head_conv1_bwd(const float* g, const float* w, float* dx, int B, int HW, int Cin, int Cout, int operand_out, void* stream);

/* col2im half of the 7x7 stem backward: dx[B, Cin, H, W] (NCHW) gathered from t[B*H*W, ld] = dY * W[Cout, Cin*49]
 * (computed by a 1x1 sdc_conv_gemm); column index ci*49 + ky*7 + kx. */
int sdc_stem_col2im(const float* t, float* dx, int B, int Cin, int H, int W, int ld, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Parameter gradients (what autograd accumulates into the U-Net parameters when the reference back-propagates the
 * inference-time fine-tuning loss through the last DDIM step, 1D/model/diffusion.py:524-551 and
 * 1D/inference/inference_ft.py:189-226, or the post-training loss, 1D/posttrain/post_train.py:206-260).  The activation
 * gradients dY come from the backward-data pass above; every output below is ACCUMULATED (+=) into a caller-zeroed
 * fp32 buffer with atomics.
 * ------------------------------------------------------------------------------------------------------------- */

/* Weight gradient of a convolution in the torch OIHW layout: dw[co, ci, ky, kx] += sum_p dy[p, co] * a[p + (ky,kx), ci].
 * kind as in sdc_conv_gemm (0 = 1x1, 1 = 3x3 pad 1, 2 = pixel-unshuffle + 1x1: dw[co, c*4 + 2 p1 + p2], input 2H x 2W);
 * a0 | a1: the NHWC input segments of the forward call (fp16 if a_half else fp32), dy: [B*H*W, Cout] fp32.
 * mma.sync TF32 with operands rounded to nearest, fp32 accumulation. */
int sdc_conv_wgrad(int kind, int a_half, const void* a0, int c0, const void* a1, int c1, const float* dy, float* dw,
                   int B, int H, int W, int Cout, void* stream);

/* The same weight gradient on tcgen05 (kind 0 / 1; Cout, c0, c1 multiples of 128; W in {16, 32, 64, 128}): the PIXEL axis is the
 * GEMM's K, both operands are read MN-major straight from their pixel-row layout (TMA boxes of dY and of the tap-shifted,
 * zero-padded activation window; tcgen05.mma.kind::tf32 with MN-major descriptors), fp32 accumulation in TMEM, pixel slices of a
 * tile added into dw with red.global.add.f32.  fp16 activations are first converted (exactly) to fp32 into `scratch`
 * (sdc_conv_wgrad_tc_scratch bytes, caller-owned).  Returns -1 (nothing done) for shapes it does not take: use sdc_conv_wgrad. */
int64_t sdc_conv_wgrad_tc_scratch(int a_half, int c0, int c1, int B, int H, int W);
int sdc_conv_wgrad_tc(int kind, int a_half, const void* a0, int c0, const void* a1, int c1, const float* dy, float* dw, int B, int H,
                      int W, int Cout, void* scratch, int64_t scratch_bytes, void* stream);

/* out[c] += sum over the M rows of x[M, C] (conv bias gradient from dY). */
int sdc_colsum(const float* x, float* out, int64_t M, int C, void* stream);

/* Per (sample, channel) sums of the GroupNorm+FiLM+SiLU backward: P[b, 0, c] += sum_p dz, P[b, 1, c] += sum_p dz * xhat with
 * dz = dy * silu'(z); arguments as sdc_gn_silu_bwd.  From these: d gamma_c = sum_b (1+sc_bc) P1, d beta_c = sum_b (1+sc_bc) P0,
 * d scale_bc = gamma_c P1 + beta_c P0, d shift_bc = P0. */
int sdc_gn_param_grad(const float* dy, const float* x, const double* stats, const float* gamma, const float* beta,
                      const float* scale_shift, const int32_t* t_index, int64_t ss_stride, float* P, int B, int HW, int C,
                      void* stream);

/* dg[c] += sum_rows dy[row, c] * xhat[row, c] for the channel LayerNorm (x: its forward input, fp16 if x_half). */
int sdc_channel_layernorm_gain_grad(const float* dy, const void* x, int x_half, float* dg, int64_t M, int C, void* stream);

/* Head 1x1 conv: dw[o, c] += sum g[b, o, p] x[(b,p), c], db[o] += sum g[b, o, p]  (g: NCHW fp32 gradient of the output). */
int sdc_head_conv1_wgrad(const float* g, const void* x, int x_half, float* dw, float* db, int B, int HW, int Cin, int Cout,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif
