/*
 * libaqualora_b200.so -- C ABI of the B200-native AquaLoRA hot paths.
 *
 * Every entry point replaces one piece of the reference's Python hot path (file:line cited per function,
 * relative to the AquaLoRA repository).  Conventions:
 *   - plain C, all pointers are DEVICE pointers owned by the caller unless a name ends in `_host`;
 *   - the library never allocates user-visible memory; scratch comes in through `ws` (query the size with
 *     the matching `*_workspace_bytes`);
 *   - every call is asynchronous and ordered on `stream` (a cudaStream_t passed as void*);
 *   - returns 0 (AQ_OK) or a negative AQ_ERR_* code, never throws; aq_last_error() returns a thread-local
 *     explanation of the last failure on the calling thread;
 *   - bf16 tensors are raw uint16 bfloat16 bit patterns, row-major, with explicit leading dimensions (in
 *     elements) where a view is allowed.
 */
#ifndef AQUALORA_B200_H_
#define AQUALORA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AQ_OK 0
#define AQ_ERR_BAD_SHAPE (-1)
#define AQ_ERR_BAD_DTYPE (-2)
#define AQ_ERR_BAD_ALIGN (-3)
#define AQ_ERR_ARCH (-4)
#define AQ_ERR_LAUNCH (-5)
#define AQ_ERR_WORKSPACE (-6)

int aq_version(void);               /* ABI version, bumped on any signature change */
int aq_arch(void);                  /* 100: the only architecture this library is built for (sm_100a) */
const char* aq_last_error(void);    /* thread-local, never NULL */
int aq_sm_count(void);              /* SM count of the current device, or <0 */
long long aq_launch_count(void);    /* kernels this library has launched in this process (bench.py: gpu_launches) */

/* ------------------------------------------------------------------------------------------------
 * (i) watermark-LoRA projection.  Replaces utils/lora_modules.py:9-26 (CustomLoRALinearLayerforward)
 * fused with the base op of utils/lora_modules.py:56-62 (CustomLoRACompatibleLinearforward):
 *
 *      H  = X Dn^T                       [M, r]
 *      Y  = X W^T + bias + (H (.) scale[row / tokens_per_sample, :]) Up^T        [M, dout]
 *
 * x      [M, din]   bf16, row stride ldx          w     [dout, din] bf16 (nn.Linear.weight)
 * bias   [dout]     bf16 or NULL                  down  [r, din]    bf16 (lora_layer.down.weight) or NULL
 * up     [dout, r]  bf16 (lora_layer.up.weight)   scale [M/tokens_per_sample, r] fp32 -- the EFFECTIVE
 *        diagonal: mapper(msg) for a tensor scale, the float multiplier broadcast for a float scale, times
 *        network_alpha/rank when set (lora_modules.py:15-25).
 * y      [M, dout]  bf16, row stride ldy          h_save [M, r] bf16 or NULL: H before scaling (kept for
 *        the backward).
 * down == NULL runs the plain base projection (lora_layer is None, lora_modules.py:57-59).
 * Constraints: din % 8 == 0, dout % 8 == 0, r % 8 == 0, r >= 8 (ranks above 64 run as ceil(r / 64) launches over 64-wide slices
 * of the LoRA operands, the later ones accumulating into Y), 16-byte aligned pointers.
 * One kernel: TMA-staged tiles, tcgen05 MMA with TMEM accumulators; H never leaves the SM on its way
 * into the second contraction.
 * ---------------------------------------------------------------------------------------------- */
int aq_lora_linear_fwd(const void* x, int64_t ldx, const void* w, const void* bias, const void* down,
                       const void* up, const float* scale, void* y, int64_t ldy, void* h_save, int64_t M,
                       int64_t tokens_per_sample, int din, int dout, int r, void* stream);
/* The same with a residual stream added in the tile epilogue: Y = X W^T + bias + LoRA + residual (residual [M, dout] bf16, row stride
 * ldres; NULL = none).  In a Transformer2D block the three `x + to_out(...)`, `x + ff(...)`, `res + proj_out(...)` adds
 * (scripts/lib/original_unet.py:779-806, :881-891) then cost no extra pass over the activations. */
int aq_lora_linear_fwd_residual(const void* x, int64_t ldx, const void* w, const void* bias, const void* down, const void* up,
                                const float* scale, const void* residual, int64_t ldres, void* y, int64_t ldy, void* h_save,
                                int64_t M, int64_t tokens_per_sample, int din, int dout, int r, void* stream);

/* Test / tuning hook: pin the column-tile width (64, 128, 160 or 192) and the number of column tiles per work
 * item for the calling thread's next aq_lora_linear_* launches; (0, 0) restores the built-in heuristics. */
int aq_lora_set_tuning(int block_n, int group_size);

/* Backward of the above (what autograd derives from utils/lora_modules.py:9-26,56-62; closed form in
 * DESIGN.md):
 *      dHs = G Up            dH = dHs (.) scale        dX  = G W + dH Dn
 *      dUp += G^T (H (.) scale)      dDn += dH^T X      dscale[b] += sum_n dHs (.) H
 * gy [M, dout] bf16 (ldgy); x [M, din] bf16 (ldx); h_save [M, r] bf16 from the forward.
 * w_t [din, dout] bf16 = W^T (cached once, W is frozen) -- NULL skips dX (inputs that need no gradient,
 * e.g. the text context of attn2.to_k / to_v); down_t [din, r] bf16 = Dn^T; up_t [r, dout] bf16 = Up^T.
 * gx [M, din] bf16 (ldgx) or NULL.  g_down [r, din], g_up [dout, r], g_scale [B, r] are fp32 and are
 * ACCUMULATED into (the caller zeroes the flat gradient buffer once per step); g_scale may be NULL.
 * ws: aq_lora_linear_bwd_workspace_bytes(M, r) bytes of scratch. */
/* Several projections of the SAME input in one launch: the 32 cross-attention to_k / to_v projections of a U-Net forward all
 * read encoder_hidden_states, to_q / to_k / to_v of a self-attention read the same tokens (what diffusers' AttnProcessor issues
 * as separate module calls, each the op sequence of utils/lora_modules.py:9-26,56-62).  proj[i]: the per-projection operands of
 * aq_lora_linear_fwd; x, scale, M, tokens_per_sample, din, r are shared.  down == NULL in every entry runs plain base
 * projections.  1 <= nproj <= 32. */
typedef struct aq_lora_projection {
  const void* w;        /* [dout, din] bf16 */
  const void* bias;     /* [dout] bf16 or NULL */
  const void* down;     /* [r, din] bf16 or NULL (all entries alike) */
  const void* up;       /* [dout, r] bf16 */
  void* y;              /* [M, dout] bf16, row stride ldy */
  int64_t ldy;
  void* h_save;         /* [M, r] bf16 or NULL */
  int dout;
} aq_lora_projection;
int aq_lora_linear_fwd_grouped(const void* x, int64_t ldx, const aq_lora_projection* proj, int nproj, const float* scale,
                               int64_t M, int64_t tokens_per_sample, int din, int r, void* stream);

size_t aq_lora_linear_bwd_workspace_bytes(int64_t M, int r);
int aq_lora_linear_bwd(const void* gy, int64_t ldgy, const void* x, int64_t ldx, const void* w_t,
                       const void* down_t, const void* up_t, const float* scale, const void* h_save, void* gx,
                       int64_t ldgx, float* g_down, float* g_up, float* g_scale, int64_t M,
                       int64_t tokens_per_sample, int din, int dout, int r, void* ws, size_t ws_bytes,
                       void* stream);
/* The two halves of aq_lora_linear_bwd separately.  `_bwd_dx` writes gx (and accumulates g_scale) and leaves dH / Hs in `ws`;
 * `aq_lora_wgrad_batch` then accumulates dUp / dDn of up to 64 layers in as few launches as possible (a layer's weight gradients feed
 * nothing downstream, so a caller queues the jobs of several layers -- each keeps its gy, x and ws alive -- and flushes once). */
typedef struct aq_wgrad_job {
  const void* gy; int64_t ldgy;   /* [M, dout] bf16 */
  const void* x; int64_t ldx;     /* [M, din] bf16 */
  const void* ws;                 /* the workspace aq_lora_linear_bwd_dx filled for this layer */
  float* g_down;                  /* [r, din] fp32, accumulated */
  float* g_up;                    /* [dout, r] fp32, accumulated */
  int64_t M;
  int din, dout, r;
} aq_wgrad_job;
int aq_lora_linear_bwd_dx(const void* gy, int64_t ldgy, const void* w_t, const void* down_t, const void* up_t, const float* scale,
                          const void* h_save, void* gx, int64_t ldgx, float* g_scale, int64_t M, int64_t tokens_per_sample, int din,
                          int dout, int r, void* ws, size_t ws_bytes, void* stream);
int aq_lora_wgrad_batch(const aq_wgrad_job* jobs, int njobs, void* stream);

/* Skinny weight-gradient contraction used by the backward:  C[i, j] (+)= sum_m P[m, i] * Q[m, j]
 * P [M, I] bf16 (ldp), Q [M, J] bf16 (ldq, J <= 64), C fp32 [I, J] (transpose_out = 0) or [J, I] (= 1),
 * accumulated with fp32 reductions in L2.  Exposed for tests and for create_wm_lora/merge tooling. */
int aq_wgrad_tn(const void* p, int64_t ldp, const void* q, int64_t ldq, float* c, int64_t ldc, int64_t M, int I,
                int J, int transpose_out, void* stream);

/* MapperNet (utils/models.py:98-115):  scale[b, :] = 1 + (msg[b, :] @ E) / sqrt(bits);  E [bits, r] fp32.
 * out_bf16_rounded != 0 rounds the result to bf16 precision (train/ppft_train.py:990 casts to weight dtype). */
int aq_mapper_fwd(const float* msg, const float* emb, float* scale, int B, int bits, int r, int out_bf16_rounded,
                  void* stream);
/* g_emb[i, :] += sum_b msg[b, i] * g_scale[b, :] / sqrt(bits) */
int aq_mapper_bwd(const float* msg, const float* g_scale, float* g_emb, int B, int bits, int r, void* stream);

/* fp32 master LoRA parameters -> the bf16 operand copies the kernels consume.  src [rows, cols] fp32;
 * dst [rows, cols] bf16 and (optional) dst_t [cols, rows] bf16. */
int aq_cast_transpose_bf16(const float* src, void* dst, void* dst_t, int rows, int cols, void* stream);
/* The same for many matrices in one launch (the 384 LoRA matrices after every optimizer step).  jobs: DEVICE table; entry i
 * covers 32x32 tiles [tile_begin, tile_begin + tiles_x * ceil(rows / 32)) of the launch, tiles_x = ceil(cols / 32); entries are
 * sorted by tile_begin, total_tiles = the sum.  dst or dst_t may be 0. */
typedef struct aq_cast_job {
  int64_t src;        /* const float* [rows, cols] */
  int64_t dst;        /* bf16 [rows, cols] or 0 */
  int64_t dst_t;      /* bf16 [cols, rows] or 0 */
  int64_t rows, cols, tile_begin, tiles_x;
} aq_cast_job;
int aq_cast_transpose_bf16_batched(const aq_cast_job* jobs, int njobs, int64_t total_tiles, void* stream);
/* bf16 [rows, cols] -> bf16 [cols, rows] (frozen W -> W^T cache for the dX contraction) */
int aq_transpose_bf16(const void* src, void* dst, int rows, int cols, void* stream);

/* Fused global-norm clip + AdamW over the flat fp32 LoRA parameter/gradient buffers
 * (train/ppft_train.py:1059-1068: clip_grad_norm_(1.0), AdamW step, zero_grad).
 * Pass 1 (aq_flat_sumsq) accumulates sum(g^2) into norm_sq[0] (caller zeroes it);  pass 2 applies
 * g *= min(1, max_norm / (sqrt(norm_sq * grad_scale^2) + 1e-6)) * grad_scale, the decoupled-weight-decay Adam
 * update, and zeroes g. */
int aq_flat_sumsq(const float* g, int64_t n, float* norm_sq, void* stream);
int aq_flat_clip_adamw(float* p, float* g, float* m, float* v, int64_t n, const float* norm_sq, float grad_scale,
                       float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                       int step, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (ii) latent watermark encoder.  SecretEncoder.forward / encode (utils/models.py:70-81, layers :57-64):
 *   c = bilinear_resize(Conv3x3(nearest_up(repeat4(view(SiLU(Linear(msg))))))), x_out = x + c.
 * msg [B, bits] fp32; w1 [base*base, bits], b1 [base*base] (secret_scaler.0); wc [4, 4, 3, 3], bc [4]
 * (secret_scaler.5); x, c_out, x_out [B, 4, H, W] fp32 NCHW (x / x_out may both be NULL: encode only).
 * base = 32, res = 64 in the reference.  ws: aq_secret_encoder_workspace_bytes(B, res). */
size_t aq_secret_encoder_workspace_bytes(int B, int res);
int aq_secret_encoder_fwd(const float* msg, const float* w1, const float* b1, const float* wc, const float* bc,
                          const float* x, float* c_out, float* x_out, int B, int bits, int base, int res, int H, int W,
                          void* ws, void* stream);

/* Backward of aq_secret_encoder_fwd (the encoder is trained through the VAE decoder, the noise layer and the message decoder:
 * train/latent_wm_pretrain.py:174-216).  g_c [B, 4, H, W] = dL/dc (add dL/dx_out to it first: x_out = x + c);
 * g_w1 [base*base, bits], g_b1 [base*base], g_wc [4, 4, 3, 3], g_bc [4] fp32 are ACCUMULATED into. */
size_t aq_secret_encoder_bwd_workspace_bytes(int B, int base, int res);
int aq_secret_encoder_bwd(const float* g_c, const float* msg, const float* w1, const float* b1, const float* wc, float* g_w1,
                          float* g_b1, float* g_wc, float* g_bc, int B, int bits, int base, int res, int H, int W, void* ws,
                          size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (ii) noise_layers distortion stack (utils/noise_layers/*).  x, y: [B, 3, H, W] fp32 NCHW in [-1, 1]; y != x.
 * Every quantity the reference layer samples internally is an explicit argument here (Noiser draws them on the
 * host, aqualora_b200/noise_layers.py), so results are reproducible against the CPU oracle.
 * ---------------------------------------------------------------------------------------------- */
/* JpegCompression.forward (jpeg_compression.py:130-162): zero-pad to x8, RGB->YUV, 8x8 DCT, keep the first
 * (25, 9, 9) zig-zag coefficients of (Y, U, V), IDCT, YUV->RGB, un-pad.  No quantisation. */
int aq_noise_jpeg(const float* x, float* y, int B, int H, int W, void* stream);
/* CropandResize.forward (noises.py:46-57): crop (top, left, crop_h, crop_w) -> bilinear (resize_h, resize_w) ->
 * bilinear (out_h, out_w), align_corners = False, no antialias; one box for the whole batch.  y: [B, 3, out_h, out_w]. */
int aq_noise_crop_resize(const float* x, float* y, int B, int H, int W, int top, int left, int crop_h, int crop_w,
                         int resize_h, int resize_w, int out_h, int out_w, void* stream);
/* GaussianBlur.forward (noises.py:67-70; kornia RandomGaussianBlur((ky, kx), sigma)): separable normalised Gaussian,
 * reflect border, sigmas [B] fp32 on the device (one per sample, both axes). */
int aq_noise_gauss_blur(const float* x, float* y, const float* sigmas, int B, int H, int W, int ky, int kx, void* stream);
/* GaussianNoise.forward (noises.py:80-85): y = x + std * N(0, 1) over n elements; the normals are element e ->
 * Philox4x32-10(key = seed, counter = offset + e / 4)[e % 4] through Box-Muller.  x == NULL writes std * noise only
 * (std = 1 gives the exact noise tensor, which tests hand to the oracle). */
int aq_noise_gauss_noise(const float* x, float* y, int64_t n, float std, uint64_t seed, uint64_t offset, void* stream);
/* ColorJitter.forward (noises.py:95-104; kornia ColorJiggle): params [B, 4] fp32 on the device = (brightness,
 * contrast, saturation, hue) per sample; order_host[4] = permutation of (0 brightness, 1 contrast, 2 saturation, 3 hue). */
int aq_noise_color_jiggle(const float* x, float* y, const float* params, const int* order_host, int B, int H, int W,
                          void* stream);

/* Input gradients of the noise layers (train/latent_wm_pretrain.py:186-190 back-propagates through `noiser(...)` into the
 * encoder).  gy: dL/dy of the matching forward call with the SAME parameters; gx: dL/dx, fully written.
 *   jpeg          linear and self-adjoint up to the colour matrices: the forward kernel with C1^T / C2^T
 *   crop_resize   adjoint of the fused double bilinear gather (fp32 atomics: run-to-run summation order varies)
 *   gauss_blur    adjoint of the reflect-border separable filter, as a gather
 *   gauss_noise   identity (gx = gy): no entry point
 *   color_jiggle  gx = J^T gy with the per-pixel 3 x 3 Jacobian of the executed arithmetic (forward-mode duals in the kernel) */
int aq_noise_jpeg_bwd(const float* gy, float* gx, int B, int H, int W, void* stream);
int aq_noise_crop_resize_bwd(const float* gy, float* gx, int B, int H, int W, int top, int left, int crop_h, int crop_w,
                             int resize_h, int resize_w, int out_h, int out_w, void* stream);
int aq_noise_gauss_blur_bwd(const float* gy, float* gx, const float* sigmas, int B, int H, int W, int ky, int kx, void* stream);
int aq_noise_color_jiggle_bwd(const float* x, const float* gy, float* gx, const float* params, const int* order_host, int B,
                              int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (ii) losses of the pretraining step (train/latent_wm_pretrain.py).
 * PRVL_loss (:42-50): max over all positions (batch included) of the 32 x 32 box mean, zero padding 16, of the channel-mean
 * |img1 - img2|; img [B, 3, H, W] fp32.  loss[0] receives the value; state (8 bytes, 8-byte aligned) keeps the arg-max for the
 * backward, which sends g_loss[0] / (1024 * 3) * sign(img1 - img2) to the winning window (g_img1 and / or g_img2, fully written).
 * binary_cross_entropy_with_logits (:200), mean reduction: loss[0] and, when g_logits != NULL, d loss / d logits. */
size_t aq_prvl_workspace_bytes(int B, int H, int W);
int aq_prvl_loss_fwd(const float* img1, const float* img2, float* loss, void* state, int B, int H, int W, void* ws, size_t ws_bytes,
                     void* stream);
int aq_prvl_loss_bwd(const float* img1, const float* img2, const void* state, const float* g_loss, float* g_img1, float* g_img2,
                     int B, int H, int W, void* stream);
int aq_bce_logits(const float* logits, const float* targets, float* loss, float* g_logits, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (ii) message decoder, train mode (`sec_decoder.train()`: train/latent_wm_pretrain.py:160, rob_enhance_finetune.py:980): the
 * batch-statistics BatchNorm2d (+ SiLU) of torchvision's Conv2dNormActivation over NHWC fp32 rows z [M = B*H*W, C].
 * fwd: mean / biased variance over the M rows; y = act(gamma (z - mean) rstd + beta), act = SiLU when `act` != 0; mean_rstd [C, 2]
 *      is written for the backward; running_mean / running_var (both or neither) are updated with `momentum` (unbiased variance).
 * bwd: gz fully written; g_gamma / g_beta (either may be NULL) are ACCUMULATED into.  ws: aq_bn_train_workspace_bytes(C).
 * C % 4 == 0, C <= 2048. */
size_t aq_bn_train_workspace_bytes(int C);
int aq_bn_train_fwd(const float* z, const float* gamma, const float* beta, float* running_mean, float* running_var, float* mean_rstd,
                    float* y, int64_t M, int C, float eps, float momentum, int act, void* ws, size_t ws_bytes, void* stream);
int aq_bn_train_bwd(const float* gy, const float* z, const float* gamma, const float* beta, const float* mean_rstd, float* gz,
                    float* g_gamma, float* g_beta, int64_t M, int C, int act, void* ws, size_t ws_bytes, void* stream);

/* Convolutions of the decoder's train path (csrc/decoder_train.cu): fp32, NHWC activations, weights in PyTorch's layouts.
 *   aq_dwconv_fwd / _bwd    depthwise k x k (k = 3, 5), stride 1 / 2, padding (k - 1) / 2, no bias: x [B, H, W, C], w [C, 1, k, k],
 *                           z [B, Ho, Wo, C]; bwd writes gx (may be NULL) and ACCUMULATES gw (may be NULL)
 *   aq_conv1x1_wgrad        gw [N, K] += gz^T [N, M] (x (.) se) [M, K]; se [M / hw, K] = the SE gate of each image or NULL.  (Forward
 *                           and input gradient of the pointwise convolutions are aq_conv1x1_tf32x3 with w resp. w^T.)
 *   aq_stem_conv_fwd / _bwd 3 x 3 stride 2 padding 1, 3 -> 32: x [B, 3, H, W] NCHW, w [32, 3, 3, 3], z [B, Ho, Wo, 32] NHWC;
 *                           bwd writes gx NCHW (may be NULL) and ACCUMULATES gw (may be NULL) */
int aq_dwconv_fwd(const float* x, const float* w, float* z, int B, int H, int W, int C, int k, int stride, void* stream);
int aq_dwconv_bwd(const float* gz, const float* x, const float* w, float* gx, float* gw, int B, int H, int W, int C, int k, int stride,
                  void* stream);
int aq_conv1x1_wgrad(const float* gz, const float* x, const float* se, float* gw, int64_t M, int K, int N, int hw, void* stream);
int aq_stem_conv_fwd(const float* x, const float* w, float* z, int B, int H, int W, void* stream);
int aq_stem_conv_bwd(const float* gz, const float* x, const float* w, float* gx, float* gw, int B, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * (ii) message decoder.  SecretDecoder.forward (utils/models.py:91-96 == evaluation/utils_eval.py:149-154):
 * torchvision EfficientNet-B1 (eval) with a Linear(1280, out_features) head; out_features = 2 * bits, viewed as
 * [B, bits, 2]; bit = argmax over the pair (evaluation/utils_eval.py:194).
 * x [B, 3, 512, 512] fp32 NCHW in [-1, 1] (resize other sizes first: aq_noise_crop_resize with a full-image crop is
 * the reference's bilinear F.interpolate).  packed: aq_effnetb1_packed_floats(out_features) fp32 values, BatchNorm
 * folded into the preceding conv, in execution order (layout: aqualora_b200/decoder.py:pack_state_dict).
 * logits [B, out_features] fp32; bits [B, out_features / 2] u8 or NULL.  fp32-faithful arithmetic: pointwise convolutions
 * as 3-term split-TF32 tensor-core products with fp32 accumulation, everything else fp32 FFMA.
 * ---------------------------------------------------------------------------------------------- */
size_t aq_effnetb1_packed_floats(int out_features);
size_t aq_effnetb1_workspace_bytes(int B);
int aq_effnetb1_fwd(const float* x, const float* packed, float* logits, unsigned char* bits, int B, int out_features,
                    void* ws, size_t ws_bytes, void* stream);

/* The two building blocks of the decoder's MBConv blocks (torchvision MBConv.forward as used by utils/models.py:88-96),
 * exported for kernel-level parity tests:
 * aq_conv1x1_tf32x3: y[m, n] = epi(sum_k (x[m, k] * se[m / hw, k]) * (w_hi + w_lo)[n, k] + bias[n]) over NHWC pixels
 *   x [M, K], y [M, N] fp32; w_hi / w_lo [N, K] = exact TF32 split of the BatchNorm-folded conv weight; se [M / hw, K] or NULL;
 *   epi: 0 none, 1 SiLU, 2 + residual[M, N], 3 SiLU then per-image column sums into y [M / hw, N] (y must be zeroed).
 * aq_depthwise_silu: y = SiLU(depthwise_k x k, stride s (x) + bias), NHWC, pad (k - 1) / 2; pooled [B, C] += sum over pixels
 *   x [B, H, H, C], w [k * k, C], y [B, Ho, Ho, C]. */
int aq_conv1x1_tf32x3(const float* x, const float* w_hi, const float* w_lo, const float* bias, const float* se,
                      const float* residual, float* y, int64_t M, int K, int N, int hw, int epi, void* stream);
int aq_depthwise_silu(const float* x, const float* w, const float* bias, float* y, float* pooled, int B, int H, int C, int k,
                      int stride, void* stream);
/* aq_expand_dw_fused: the two calls above as ONE kernel for the early blocks (expand 1x1 + SiLU -> depthwise + SiLU + squeeze sums;
 *   the 6x expanded map stays in shared memory): y [B, Ho, Ho, cexp] = SiLU(dw_k,s(SiLU(x W_e^T + b_e)) + b_d), pooled += sum.
 *   Supported (cin, cexp, k, stride): (16, 96, 3, 2), (24, 144, 3, 1), (24, 144, 5, 2) with Ho a multiple of 8 (16 for the stride-1 shape); other shapes
 *   return AQ_ERR_BAD_SHAPE (aq_effnetb1_fwd then runs the unfused pair).  Same arithmetic as aq_conv1x1_tf32x3 (3-term TF32
 *   split, fp32 accumulate) and aq_depthwise_silu (fp32 FMA). */
int aq_expand_dw_fused(const float* x, const float* w_hi, const float* w_lo, const float* b_e, const float* w_d, const float* b_d,
                       float* y, float* pooled, int B, int H, int cin, int cexp, int k, int stride, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Message -> deployable weights (scripts/create_wm_lora.py:23-41, scripts/merge_lora.py:98-120).
 * aq_lora_fold_down: out[i, :] = (down[i, :] * m[i]) * scale, down / out [r, cols] fp32, m [r] = mapper(msg) (the
 *   reference's diag_embed(m) @ down * scale for linear targets and down * m[:, None, None, None] * scale for 1x1 convs).
 * aq_lora_merge: w[dout, din] += coef * (up[dout, r] @ down[r, din]), fp32, coef = ratio * alpha / dim, r <= 64.
 * ---------------------------------------------------------------------------------------------- */
int aq_lora_fold_down(const float* down, const float* m, float* out, int r, int64_t cols, float scale, void* stream);
int aq_lora_merge(float* w, const float* up, const float* down, int dout, int din, int r, float coef, void* stream);

/* ------------------------------------------------------------------------------------------------
 * HBM-bound glue of the U-Net around the LoRA projections, on channels-last bf16 rows.
 *
 * GroupNorm [+ per-(sample, channel) add before it] [+ SiLU after it]: replaces `norm1 -> silu`,
 * `+ temb[:, :, None, None] -> norm2 -> silu` of ResnetBlock2D.forward (scripts/lib/original_unet.py:440-453),
 * Transformer2DModel.norm (:826) and conv_norm_out -> silu (:1416).
 *   x, y, dy, dx  [B, HW, C] bf16 (the channels_last storage of an NCHW tensor)      gamma, beta [C] bf16
 *   add_bc        [B, C] bf16 or NULL: x + add_bc[b, c] is what gets normalised (its gradient is not produced: the
 *                 time embedding is not trainable)
 *   mean_rstd     [B, G, 2] fp32 (mean, 1/sqrt(var + eps)), written by fwd, read by bwd
 *   ws            aq_group_norm_workspace_bytes(B, G) bytes.   C % 8 == 0, C % G == 0, G <= 128.
 * The backward returns dx only (gamma / beta belong to the frozen U-Net).
 * ---------------------------------------------------------------------------------------------- */
size_t aq_group_norm_workspace_bytes(int B, int G);
int aq_group_norm_nhwc_fwd(const void* x, const void* gamma, const void* beta, const void* add_bc, void* y, float* mean_rstd, int B,
                           int HW, int C, int G, float eps, int silu, void* ws, size_t ws_bytes, void* stream);
int aq_group_norm_nhwc_bwd(const void* dy, const void* x, const void* gamma, const void* beta, const void* add_bc,
                           const float* mean_rstd, void* dx, int B, int HW, int C, int G, float eps, int silu, void* ws, size_t ws_bytes,
                           void* stream);

/* GEGLU.forward (scripts/lib/original_unet.py:708-729): out = proj[:, :F] * gelu(proj[:, F:2F]) (erf gelu), and its
 * backward g_proj = [g_out * gelu(gate), g_out * h * gelu'(gate)].  proj [M, 2F] bf16 with row stride ldp; out, g_out
 * [M, F] and g_proj [M, 2F] contiguous bf16; F % 8 == 0. */
int aq_geglu_fwd(const void* proj, int64_t ldp, void* out, int64_t M, int F, void* stream);
int aq_geglu_bwd(const void* proj, int64_t ldp, const void* g_out, void* g_proj, int64_t M, int F, void* stream);

/* out = a + b + bias[c] over channels-last bf16 rows [rows, C]: the closing `input_tensor + hidden_states` of ResnetBlock2D.forward
 * (scripts/lib/original_unet.py:455-460) with the biases of conv2 / conv_shortcut folded in (the convolutions then run without
 * cuDNN's separate broadcast-bias pass).  C % 8 == 0. */
int aq_add_bias_rows(const void* a, const void* b, const void* bias, void* out, int64_t rows, int C, void* stream);

/* LayerNorm over contiguous bf16 token rows [M, C] (BasicTransformerBlock.norm1/2/3, scripts/lib/original_unet.py:732-806), eps
 * inside the square root, fp32 statistics, one rounding of the output.  mean_rstd [M, 2] fp32 is written by fwd (may be NULL when
 * no backward follows) and read by bwd; the backward returns dx only (gamma / beta are frozen).  C % 8 == 0, C <= 2048. */
int aq_layer_norm_fwd(const void* x, const void* gamma, const void* beta, void* y, float* mean_rstd, int64_t M, int C, float eps,
                      void* stream);
int aq_layer_norm_bwd(const void* dy, const void* x, const void* gamma, const float* mean_rstd, void* dx, int64_t M, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AQUALORA_B200_H_ */
