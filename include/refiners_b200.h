/*
 * refiners_b200 C ABI  -  the drop-in boundary of the B200-native hot path.
 *
 * The reference (finegrain-ai/refiners) has no FFI: its seam is the Python leaf-module
 * contract of refiners.fluxion.layers.  Each entry point below is
 * what a binding for that leaf would call; the comment on each names the reference code
 * it replaces (paths relative to /root/reference/src/refiners/).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.
 *   - every function returns 0 on success, <0 on error; rb200_last_error() gives the text.
 *   - never allocates device memory, never synchronises, never throws: work is enqueued on
 *     the cudaStream_t passed as `stream` (void* so that this header needs no CUDA include).
 *   - all tensor pointers are DEVICE pointers; leading dimensions / strides are in ELEMENTS.
 *   - dtype: element type of activations and weights (accumulation is always fp32).
 */
#ifndef REFINERS_B200_H
#define REFINERS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_ABI_VERSION 1

typedef enum { RB200_BF16 = 0, RB200_FP16 = 1, RB200_FP32 = 2 } rb200_dtype;

/* fused epilogues of rb200_linear / rb200_conv2d */
typedef enum {
  RB200_EPI_NONE = 0,
  RB200_EPI_GEGLU = 1, /* y[:, j] = v_j * gelu_erf(g_j); W rows must be packed by rb200_geglu_pack */
  RB200_EPI_GELU = 2,  /* exact (erf) GeLU   - fluxion/layers/activations.py:83-114 */
  RB200_EPI_SILU = 3   /* x * sigmoid(x)     - fluxion/layers/activations.py:31-41  */
} rb200_epilogue;

/* unary ops of rb200_unary */
typedef enum {
  RB200_UNARY_SILU = 0,
  RB200_UNARY_GELU = 1,
  RB200_UNARY_GELU_TANH = 2,
  RB200_UNARY_GELU_SIGMOID = 3, /* x * sigmoid(1.702 x) */
  RB200_UNARY_RELU = 4,
  RB200_UNARY_SIGMOID = 5
} rb200_unary_op;

/* One LoRA of a LoraAdapter: y += scale * (x down^T) up^T
 * (fluxion/adapters/lora.py:14-99 Lora = Chain(down, up, Multiply(scale)); :383-448 LoraAdapter) */
typedef struct {
  const void* down; /* [rank, K] row-major */
  const void* up;   /* [N, rank] row-major */
  float scale;
  int32_t rank;
} rb200_lora;

int rb200_abi_version(void);
const char* rb200_last_error(void);
/* SM count and compute capability of the current device; 0 on success */
int rb200_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t rb200_launch_count(void);
/* select kernel family: 0 = auto (tcgen05 where shapes allow), 1 = force the SIMT reference
 * kernels (parity triage).  Returns the previous value. */
int rb200_set_kernel_mode(int mode);

/* ---- Linear -------------------------------------------------------------------------------
 * Replaces fluxion/layers/linear.py:9-58 (torch.nn.Linear.forward -> aten::addmm), plus the
 * modules the reference composes around it:
 *   bias            linear.py:31-58
 *   LoRA            fluxion/adapters/lora.py:383-448 (Sum(target, *loras))
 *   residual        fluxion/layers/chain.py:901-927 (Residual)
 *   GEGLU           fluxion/layers/activations.py:136-160 + latent_diffusion/cross_attention.py:69-71
 *
 *   Y[M,N] = epi( X[M,K] W[N,K]^T + bias[N] + sum_i s_i (X down_i^T) up_i^T + residual[M,N] )
 *
 * LoRA operands are the packed form produced by rb200_lora_pack (r_pad = 0 disables).
 * `ws` must hold rb200_linear_workspace_bytes(M, r_pad, dtype) bytes when r_pad > 0.
 * With RB200_EPI_GEGLU, N is the PACKED width (2 * output width) and ldy/ldr refer to the
 * output of width N/2.
 */
size_t rb200_linear_workspace_bytes(int64_t M, int r_pad, int dtype);
int rb200_linear(void* stream, int dtype, const void* x, int64_t ldx, const void* w, int64_t ldw,
                 const void* bias, void* y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                 int r_pad, const void* lora_down_cat, const void* lora_up_cat,
                 const float* lora_colscale, const void* residual, int64_t ldr, int epilogue,
                 void* ws, size_t ws_bytes);

/* Pack n LoRAs for rb200_linear: down_cat[r_pad, K] (rows of all `down`, zero padded),
 * up_cat[N, r_pad] (columns of all `up`), colscale[r_pad] (scale of the owning LoRA).
 * r_pad = round_up(sum of ranks, 64). */
int rb200_lora_pack(void* stream, int dtype, int n_lora, const rb200_lora* loras, int64_t N,
                    int64_t K, void* down_cat, void* up_cat, float* colscale, int r_pad);

/* Interleave the value/gate halves of a GEGLU projection in groups of 16 rows so that one
 * epilogue tile sees both: w[2F, K] -> w_packed[2F, K], bias[2F] -> bias_packed[2F]. */
int rb200_geglu_pack(void* stream, int dtype, const void* w, const void* bias, void* w_packed,
                     void* bias_packed, int64_t F, int64_t K);

/* ---- Conv2d -------------------------------------------------------------------------------
 * Replaces fluxion/layers/conv.py:6-61 (torch.nn.Conv2d.forward -> cuDNN), zeros padding,
 * groups = 1, dilation = 1, plus:
 *   per-sample channel bias   latent_diffusion/range_adapter.py:47-86 (RangeAdapter2d)
 *   residual / shortcut sum   latent_diffusion/unet.py:27-51 (ResidualBlock is a Sum)
 * Activations are NHWC (channels-last) in memory: x[B,H,W,Cin], y[B,Ho,Wo,Cout].
 * w_packed is [R*S, Cout, Cin] (rb200_conv2d_pack_weight) - the K-major B operand per tap.
 *
 *   y[b,ho,wo,:] = sum_{r,s} x[b, ho*stride-pad+r, wo*stride-pad+s, :] W[r,s]^T
 *                  + bias + chan_bias[b,:] + residual[b,ho,wo,:]
 */
int rb200_conv2d_pack_weight(void* stream, int dtype, const void* w /* [Cout,Cin,R,S] */,
                             void* w_packed, int64_t Cout, int64_t Cin, int R, int S);
int rb200_conv2d(void* stream, int dtype, const void* x, const void* w_packed, const void* bias,
                 const void* chan_bias, const void* residual, void* y, int64_t B, int64_t H,
                 int64_t W, int64_t Cin, int64_t Cout, int R, int S, int stride, int pad,
                 int epilogue);

/* Zero-pad the channels of an image to a multiple of 8 so that narrow input convolutions (the UNet's
 * 4-channel latent conv, latent_diffusion/stable_diffusion_xl/unet.py:258-351; the 3-channel
 * ConditionEncoder stem, stable_diffusion_1/controlnet.py:16-68) run on the tensor-core path with
 * zero-extended weights: y[B,H,W,Cp] (dense NHWC) = x[B,C,H,W] addressed through element strides. */
int rb200_pad_channels(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C,
                       int Cp, int64_t sb, int64_t sc, int64_t sh, int64_t sw);

/* ---- GroupNorm (+SiLU) ----------------------------------------------------------------------
 * Replaces fluxion/layers/norm.py:52-92 (+ activations.py:31-41 when silu != 0).
 * x, y: NHWC [B, HW, C]; statistics per (sample, group) over HW * C/G elements in fp32.
 * ws: rb200_group_norm_workspace_bytes(B, HW, G) bytes. */
size_t rb200_group_norm_workspace_bytes(int64_t B, int64_t HW, int G);
int rb200_group_norm(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t HW,
                     int64_t C, int G, float eps, const void* gamma, const void* beta, int silu,
                     void* ws, size_t ws_bytes);

/* GroupNorm with FIXED statistics (tiled VAE inference): replaces FixedGroupNorm.compute_group_norm,
 * foundationals/latent_diffusion/auto_encoder.py:209-251.  stats: fp32 [B, G, 2] = (mean, 1/sqrt(var + eps))
 * per (sample, group).  frozen == 0: the statistics of x are computed, WRITTEN to stats and applied (the
 * adapter's first pass, on the downscaled image); frozen != 0: stats are read as they are and no statistics
 * pass runs (every later tile). */
int rb200_group_norm_fixed(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t HW,
                           int64_t C, int G, float eps, const void* gamma, const void* beta, int silu,
                           void* ws, size_t ws_bytes, float* stats, int frozen);

/* ---- LayerNorm ------------------------------------------------------------------------------
 * Replaces fluxion/layers/norm.py:14-49 (row LN over the last dim) and, applied to the NHWC
 * pixels of a map, fluxion/layers/norm.py:95-127 (LayerNorm2d). x, y: [rows, C] contiguous. */
int rb200_layer_norm(void* stream, int dtype, const void* x, void* y, int64_t rows, int64_t C,
                     float eps, const void* gamma, const void* beta);

/* ---- Elementwise ----------------------------------------------------------------------------
 * fluxion/layers/activations.py:31-160; GLU(GeLU): y[r, j] = x[r, j] * gelu(x[r, F + j]). */
int rb200_unary(void* stream, int dtype, const void* x, void* y, int64_t n, int op);
int rb200_geglu(void* stream, int dtype, const void* x, void* y, int64_t rows, int64_t F);
/* y = a + alpha * b (Residual / Sum of fluxion/layers/chain.py:867-927; Multiply of basics.py:379) */
int rb200_add(void* stream, int dtype, const void* a, const void* b, void* y, int64_t n, float alpha);

/* ---- Denoising-step glue --------------------------------------------------------------------
 * The elementwise part of one LatentDiffusionModel.forward (foundationals/latent_diffusion/model.py:137-159) with the
 * Euler solver (solvers/euler.py:63-100), two launches instead of one per arithmetic operator; every intermediate is
 * rounded to `dtype` exactly where the reference's operator-by-operator evaluation rounds it (bit-identical results).
 *   rb200_cfg_scale_input: y[i] (= y[n + i] when twice != 0: torch.cat((x, x)) of classifier-free guidance)
 *                            = x[i] / ((sigma^2 + 1) ^ 0.5),           sigma = *sigma (device scalar of `dtype`)
 *   rb200_cfg_euler:       noise = guided ? u + condition_scale * (c - u) : eps,  (u, c) = eps[i], eps[n + i]
 *                          y[i] = x[i] + noise * (sigmas[1] - sigmas[0])          (sigmas: two consecutive device scalars) */
int rb200_cfg_scale_input(void* stream, int dtype, const void* x, void* y, int64_t n, const void* sigma, int twice);
int rb200_cfg_euler(void* stream, int dtype, const void* x, const void* eps, void* y, int64_t n, const void* sigmas,
                    float condition_scale, int guided);

/* ---- SAM ViT data movement ------------------------------------------------------------------
 * Patch embedding (foundationals/segment_anything/image_encoder.py:9-34 PatchEncoder: Conv2d with
 * kernel = stride = P) is a GEMM over non-overlapping patches; rb200_patchify lays them out as rows:
 *   y[(b, ho, wo), (r, s, c)] = x[b, c, ho*P + r, wo*P + s]     y: [B*(H/P)*(W/P), P*P*C]
 * x is addressed through element strides (sb, sc, sh, sw), so NCHW and channels-last both work.
 * The matching weight is W[Cout, (r, s, c)]; the product goes through rb200_linear. */
int rb200_patchify(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t H, int64_t W,
                   int64_t C, int P, int64_t sb, int64_t sc, int64_t sh, int64_t sw);
/* Window partition / merge of image_encoder.py:202-237 (WindowPartition: pad + partition; WindowMerge:
 * merge + crop) on channels-last maps.
 *   merge == 0: x[B,H,W,C] -> y[B*nH*nW, ws, ws, C], nH = ceil(H/ws), nW = ceil(W/ws), zero padded
 *   merge != 0: x[B*nH*nW, ws, ws, C] -> y[B,H,W,C] (padding dropped) */
int rb200_window_partition(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W,
                           int C, int window, int merge);

/* ---- UNet skip-connection plumbing -------------------------------------------------------------
 * Channel concatenation of n <= 4 channels-last maps: y[p, :] = cat(src_0[p, :], ..., src_{n-1}[p, :]) for
 * `pixels` = B*H*W rows.  Replaces fluxion/layers/chain.py:930-964 (Concatenate, dim = 1) as used by
 * latent_diffusion/unet.py:66-79 (ResidualConcatenator).  srcs / channels are HOST arrays. */
int rb200_concat_channels(void* stream, int dtype, int n, const void* const* srcs, const int* channels,
                          void* y, int64_t pixels);
/* Nearest-neighbour resize of a channels-last map, x[B,H,W,C] -> y[B,Ho,Wo,C], source index
 * min(floor(dst * in / out), in - 1): fluxion/layers/sampling.py:13-38 (Interpolate, mode "nearest"),
 * the 2x upsampling of sampling.py:101-161 (Upsample). */
int rb200_resize_nearest(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C,
                         int Ho, int Wo);
/* k x k average pooling, stride k, of a channels-last map, x[B,H,W,C] -> y[B,H/k,W/k,C] (rows / columns beyond the last
 * full window are dropped, as torch.nn.AvgPool2d does): foundationals/latent_diffusion/t2i_adapter.py:17-19 (Downsample2d). */
int rb200_avg_pool2d(void* stream, int dtype, const void* x, void* y, int64_t B, int H, int W, int C, int k);

/* ---- Scaled dot-product attention -----------------------------------------------------------
 * Replaces fluxion/layers/attentions.py:115-202 (split heads, F.scaled_dot_product_attention,
 * merge heads).  q[B,Sq,H,D], k/v[B,Sk,H,D], o[B,Sq,H,D] addressed through (batch, seq)
 * strides in elements; heads are contiguous slices of the channel dim (stride D).
 * Optional second key/value set (IP-Adapter, latent_diffusion/image_prompt.py:237-309):
 *   o = softmax(q k^T * scale) v + scale2 * softmax(q k2^T * scale) v2
 */
int rb200_sdpa(void* stream, int dtype, const void* q, const void* k, const void* v, void* o,
               int64_t B, int H, int64_t Sq, int64_t Sk, int D, int64_t q_sb, int64_t q_ss,
               int64_t k_sb, int64_t k_ss, int64_t v_sb, int64_t v_ss, int64_t o_sb, int64_t o_ss,
               float scale, int is_causal, const void* k2, const void* v2, int64_t Sk2,
               int64_t k2_sb, int64_t k2_ss, int64_t v2_sb, int64_t v2_ss, float scale2);

/* ---- attention probabilities (self-attention guidance) ----------------------------------------
 * Replaces foundationals/latent_diffusion/self_attention_guidance.py:42-47
 * (SelfAttentionMap.compute_attention_scores):
 *   probs[b, h, i, j] = softmax_j( q[b, i, h*D:(h+1)*D] . k[b, j, h*D:(h+1)*D] * scale )
 * q: [B, Sq, H*D] with element strides (q_sb, q_ss), k likewise; probs: contiguous [B, H, Sq, Sk].
 * D % 8 == 0 (fp32: % 4), D <= 256, rows 16-byte aligned.  Logits and the softmax are fp32, the
 * stored probabilities are rounded once to `dtype`. */
int rb200_attention_probs(void* stream, int dtype, const void* q, const void* k, void* probs, int64_t B,
                          int H, int64_t Sq, int64_t Sk, int D, int64_t q_sb, int64_t q_ss,
                          int64_t k_sb, int64_t k_ss, float scale);

/* ---- StyleAligned shared attention -------------------------------------------------------------
 * Replaces one `StyleAligned` chain of foundationals/latent_diffusion/style_aligned.py:136-207 (Parallel(Identity,
 * ExtractReferenceFeatures) -> AdaIN :47-92 -> ScaleReferenceFeatures :95-133 -> Concatenate) applied to q, k or v of a
 * classifier-free-guidance batch.  x: [B, S, C] with element strides (x_sb, x_ss) - a slice of a fused q/k/v projection is
 * read in place; y: contiguous [B, S or 2S, C].  ref(b) = first image of b's half of the batch:
 *   y[b, s < S] = adain ? (x[b,s] - mean[b]) / (std[b] + eps) * std[ref(b)] + mean[ref(b)] : x[b,s]
 *   y[b, S + s] = x[ref(b), s] * (b == ref(b) ? 1 : scale)                   (rows present only when concatenate != 0)
 * mean / std: per (image, channel) over the S tokens, std unbiased; ws: rb200_style_aligned_workspace_bytes(B, C). */
size_t rb200_style_aligned_workspace_bytes(int64_t B, int64_t C);
int rb200_style_aligned(void* stream, int dtype, const void* x, void* y, int64_t B, int64_t S, int64_t C,
                        int64_t x_sb, int64_t x_ss, int adain, int concatenate, float scale, float eps,
                        void* ws, size_t ws_bytes);

/* ---- SAM decomposed relative-position attention ---------------------------------------------
 * Replaces foundationals/segment_anything/image_encoder.py:87-127 (RelativePositionAttention):
 *   logits = (q * d^-1/2) k^T + rel_h[:, :, None] + rel_w[:, None, :], softmax, @ v
 * qkv: [Bw, Hh, Ww, 3*heads*d]; rel_h_emb: [2*Hh-1, d]; rel_w_emb: [2*Ww-1, d]; o: [Bw,Hh,Ww,heads*d]
 * ws: rb200_sam_attention_workspace_bytes(...) */
size_t rb200_sam_attention_workspace_bytes(int64_t Bw, int Hh, int Ww, int heads, int d);
int rb200_sam_attention(void* stream, int dtype, const void* qkv, const void* rel_h_emb,
                        const void* rel_w_emb, void* o, int64_t Bw, int Hh, int Ww, int heads,
                        int d, void* ws, size_t ws_bytes);

#ifdef __cplusplus
}
#endif
#endif /* REFINERS_B200_H */
