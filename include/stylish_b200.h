/*
 * stylish_b200.h — C ABI of libstylish_b200.so (sm_100a kernels for the
 * Stylish-TTS forward hot path).
 *
 * The reference (Stylish-TTS/stylish-tts) is pure Python/PyTorch and has no FFI
 * of its own: the seam it offers is the nn.Module call protocol (SURVEY.md
 * §8b).  Each entry point below therefore replaces the *library call sequence*
 * the reference issues at the cited file:line (paths relative to
 * src/stylish_tts/train/).  INTEGRATION.md shows the ctypes binding.
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (torch allocates,
 *     the library never frees or retains them); `stream` is a cudaStream_t;
 *   - activations are fp32, channel-major (B, C, T) with T contiguous; where a
 *     tensor has `_bs` / `_cs` arguments they are the batch / channel strides in
 *     elements (so a slice of a wider concat buffer can be addressed);
 *   - return 0 on success, STY_ERR_* (<0) otherwise; sty_last_error() returns a
 *     thread-local message for the last failure;
 *   - no global mutable state; every call is re-entrant per stream; nothing
 *     synchronises the device.
 */
#ifndef STYLISH_B200_H
#define STYLISH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sty_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define STY_API __attribute__((visibility("default")))
#else
#define STY_API
#endif

#define STY_OK 0
#define STY_ERR_BAD_ARG (-1)
#define STY_ERR_WORKSPACE (-2)
#define STY_ERR_CUDA (-3)

/* activation codes */
#define STY_ACT_NONE 0
#define STY_ACT_RELU 1
#define STY_ACT_LEAKY02 2 /* LeakyReLU(0.2)                     ada_norm.py:148 */
#define STY_ACT_SNAKE 3   /* x + sin^2(a x)/a, a per channel    ada_norm.py:114, conv_next.py:77 */
#define STY_ACT_SWISH 4   /* x * sigmoid(x)                      conformer.py:33, SiLU duration_predictor.py:52 */
#define STY_ACT_GELU 5    /* exact (erf) GELU                    conv_next.py:115 */

STY_API int sty_version(void);
STY_API const char* sty_last_error(void);
/* number of SMs of the current device (grid sizing / diagnostics) */
STY_API int sty_device_sm_count(void);

/* ---- embedding ---------------------------------------------------------
 * out[b,c,t] = emb[tokens[b,t], c] * scale * (t < lengths[b])   (lengths NULL: no mask).
 * Replaces `self.emb(x) * sqrt(C); transpose(1,-1)` models/text_encoder.py:451-452; the
 * mask is the `x * x_mask` every consumer of the embedding applies (text_encoder.py:82,86). */
STY_API int sty_embed_fwd(const int64_t* tokens, const int64_t* lengths, const float* emb, float* out,
                          int B, int T, int C, int n_tokens, float scale, sty_stream_t stream);

/* ---- sequence mask -----------------------------------------------------------
 * out[b,t] = (t < lengths[b]) ? 1 : 0   (utils.py:54-58, as float like text_encoder.py:453) */
STY_API int sty_sequence_mask_fwd(const int64_t* lengths, float* out, int B, int T,
                                  sty_stream_t stream);

/* ---- generic Conv1d (stride 1) with fused prologue / epilogue ------------
 * y[b,co,t] = out_scale * mask_o[b,t] * act_o( bias[co] +
 *                 sum_{ci,k} w[ci,k,co] * xin[b,ci,t + k*dil - pad] )
 *             + res_scale * res[b,co,t]
 * xin[b,ci,u] = 0 outside [0,T), else act_i( in_scale[b,ci] * (x[b,ci,u]*mask_i[b,u])
 *                                            + in_shift[b,ci] )
 * Weights are PRE-PACKED as (CI, K, CO) (the reference stores (CO, CI, K));
 * `w_bs` != 0 selects a per-batch weight (used for `text_encoding @ alignment`).
 * Optional `out_sumsq[b,co] += sum_t y^2` (GRN statistics, conv_next.py:15-18).
 * `shuffle` = s > 1 stores co = c*s + r at y[b, c, t*s + r]
 * (einops "b (c s) t -> b c (t s)", generator.py:746).
 * Replaces F.conv1d / nn.Linear(+transpose) + the elementwise kernels around
 * them: text_encoder.py:79-86,195-222,325-330; ada_norm.py:109-120,176-192;
 * conv_next.py:80-93; conformer.py:84-95,173-187; generator.py:731-781,885. */
typedef struct sty_conv1d_args {
  const float* x;
  int64_t x_bs, x_cs;
  const float* w;
  int64_t w_bs;
  const float* bias; /* (CO) or NULL */
  float* y;
  int64_t y_bs, y_cs;
  const float* res; /* NULL or same indexing as y */
  int64_t r_bs, r_cs;
  const float* in_scale; /* (B,CI) or NULL */
  const float* in_shift; /* (B,CI) or NULL */
  const float* in_alpha; /* (CI): snake alpha of the prologue activation */
  const float* in_mask;  /* (B,T) or NULL: x is multiplied by it BEFORE scale / shift / activation; a NEGATIVE value
                            instead forces the activated value to 0 (zero gaps of end-to-end window layouts) */
  const float* out_mask; /* (B,T) or NULL */
  const float* out_alpha; /* (CO): snake alpha of the epilogue activation */
  float* out_sumsq;       /* (B,CO) accumulated with atomics, or NULL */
  int32_t B, CI, CO, T, K, dil, pad;
  int32_t in_act, out_act, shuffle;
  float out_scale, res_scale;
  /* Optional tensor-core path (tcgen05, "bf16x3" split precision): the same weights
   * pre-split into bf16 (hi, lo) in the UMMA K-major layout [K][2][CI/8][CO][8].
   * NULL selects the fp32 FMA kernel.  Used when CI%16==0, CO%16==0 and T>=128. */
  const void* w_split;
  /* Optional fused ConvNeXt front (tensor-core path, K == 1, CI <= 64): the conv input is
   * y = (1+gamma[b,c]) * LayerNorm_C(dwconv7(x) + dw_b)[c] + beta[b,c]   (conv_next.py:82-84)
   * dw_w (CI,7), dw_b (CI), gamma = dw_gb[b*dw_gb_bs + c], beta = dw_gb[b*dw_gb_bs + CI + c]. */
  const float* dw_w;
  const float* dw_b;
  const float* dw_gb;
  int64_t dw_gb_bs;
  float dw_eps;
  /* Optional `out_sum[b,co] += sum_t y` (with out_sumsq: the InstanceNorm statistics of the NEXT AdaIN are
   * accumulated by the conv that produces its input, ada_norm.py:109-140; tensor-core path: CO <= 64). */
  float* out_sum;
} sty_conv1d_args;
STY_API int sty_conv1d_fwd(const sty_conv1d_args* a, sty_stream_t stream);

/* ---- depthwise Conv1d ----------------------------------------------------
 * y[b,c,t] = act( post_scale[c] * (bias[c] + sum_k w[c,k] x[b,c,t+k-pad_left]) + post_shift[c] )
 * Replaces the depthwise k31 conv + eval BatchNorm1d + Swish of the conformer conv
 * module (conformer.py:176-186) and the weight-normed 1->1 k3 convs on F0 / N /
 * voiced (decoder.py:77-79). */
STY_API int sty_dwconv1d_fwd(const float* x, int64_t x_bs, int64_t x_cs, const float* w,
                     const float* bias, const float* post_scale, const float* post_shift,
                     float* y, int64_t y_bs, int64_t y_cs, int B, int C, int T, int K,
                     int pad_left, int act, sty_stream_t stream);

/* ---- ConvNeXt front: depthwise k7 + LayerNorm over C + adaptive affine ----
 * d = dwconv7(x)+bias;  y[b,c,t] = (1+gamma[b,c]) * LN_C(d)[c] + beta[b,c]
 * gamma = gb[b*gb_bs + c], beta = gb[b*gb_bs + C + c].   Replaces
 * conv_next.py:82-84 (+ AdaptiveLayerNorm ada_norm.py:203-211). */
STY_API int sty_dwconv_ln_fwd(const float* x, int64_t x_bs, const float* w, const float* bias,
                              const float* gb, int64_t gb_bs, float* y, int64_t y_bs, int B, int C,
                              int T, float eps, sty_stream_t stream);

/* ---- GeneratorConvNeXtBlock, fused, without the 4C-wide intermediate ------------------------------
 * y = x + W2 ( gs[b,:] * Snake(W1 AdaLN(LN_C(dwconv7(x))) + b1) ) + b2      (conv_next.py:80-93, GRN :7-18;
 * gs = 1 + grn_gamma * ||h||_t / (mean_j ||h||_t + 1e-6); GRN beta is folded into b2 by the caller).
 * Two passes over x (statistics, then output) on the tensor cores, fed by TMA: HBM traffic = read x twice +
 * write y.  Requirements: C = 32, J = 4C = 128, T >= 512, y != x, x / y rows 16-byte aligned with a pitch
 * >= T rounded up to 4 (x_cs, y_cs, x_bs, y_bs multiples of 4).  w1_split / w2_split: bf16 (hi, lo) packs in
 * the layout of sty_conv1d_args.w_split for K = 1 ([2][C/8][J][8] and [2][J/8][C][8]); gb: gamma | beta rows
 * of the AdaLN (B rows, stride gb_bs); sumsq, gs: (B, J) workspaces owned by the caller. */
STY_API int sty_convnext_fused_fwd(const float* x, int64_t x_bs, int64_t x_cs, float* y, int64_t y_bs, int64_t y_cs,
                                   const float* dw_w, const float* dw_b, const float* gb, int64_t gb_bs, float eps,
                                   const void* w1_split, const float* b1, const float* alpha,
                                   const float* grn_gamma, const void* w2_split, const float* b2, float* sumsq,
                                   float* gs, int B, int C, int J, int T, sty_stream_t stream);

/* ---- adversarial losses: discriminators.py:13-69, losses.py:166-373 -----------------
 * sty_leaky_s2d: y[n, c*s + p, u] = LeakyReLU_slope(x[n, c, s*u + p]) (0 past the end), x (N,C,W) -> y (N,C*s,ceil(W/s)):
 * the activation between the discriminator convs, fused with the space-to-depth rearrangement that turns a
 * stride-s convolution into a stride-1 one on s*C channels (s = 1: plain LeakyReLU).  _bwd: its gradient.
 * sty_sqdiff_sum_fwd: out[0] += sum_i (c - x[i])^2  (LSGAN terms).
 * sty_tprls_fwd: truncated pointwise relativistic least squares statistics of d = a - b:
 *   median = lower median of d (radix select, torch.median semantics), sums = [sum_mask (d-m)^2, count, sum_mask (d-m)]
 *   over the elements with a < b + m; no host synchronisation.  workspace: sty_tprls_workspace_bytes().
 * sty_tprls_bwd: da = coef[0]*mask*(d-m) + [d == m, one element]*coef[1], db = -da  (either may be NULL). */
STY_API int sty_leaky_s2d_fwd(const float* x, float* y, int64_t N, int C, int W, int s, float slope,
                              sty_stream_t stream);
STY_API int sty_leaky_s2d_bwd(const float* dy, const float* x, float* dx, int64_t N, int C, int W, int s, float slope,
                              sty_stream_t stream);
/* y[r, t] = (x ? x[r, t] : 1) * g[r] * mul over `rows` rows of T values (gate of ContextFreeDiscriminator,
 * discriminator.py:167-168; with x = NULL the broadcast of a per-row value) */
STY_API int sty_row_scale_fwd(const float* x, const float* g, float* y, int64_t rows, int T, float mul,
                              sty_stream_t stream);
/* y[r] = mul * sum_{j < P} x[r*P + j]  (P % 4 == 0): pooled value per window of the gapped waveform-discriminator
 * layout (AdaptiveAvgPool1d(1), discriminator.py:139) */
STY_API int sty_segment_sum_fwd(const float* x, float* y, int64_t rows, int P, float mul, sty_stream_t stream);
STY_API int sty_sqdiff_sum_fwd(const float* x, int64_t n, float c, float* out, sty_stream_t stream);
STY_API int64_t sty_tprls_workspace_bytes(void);
STY_API int sty_tprls_fwd(const float* a, const float* b, int64_t n, void* workspace, float* sums, float* median,
                          sty_stream_t stream);
STY_API int sty_tprls_bwd(const float* a, const float* b, int64_t n, const float* median, const float* coef, float* da,
                          float* db, int* flag, sty_stream_t stream);

/* ---- 64-wide attention, second generation (conformer.py:112-131; Transformer1d blocks of the style denoiser) --------
 * No mask / RoPE / dropout, head dim 64.  One pre-pass splits q (scaled), k, v into bf16 hi | lo planes stored as the
 * shared-memory image of each tile; the attention kernel then moves tiles with cp.async.bulk and issues the next tile's
 * S = Q K^T behind the current P V (tcgen05, bf16x3).  `workspace`: sty_attention64_workspace_bytes(B,H,T) bytes,
 * 16-byte aligned, owned by the caller (contents are scratch).
 * sty_attention64_fwd       : q, k, v (B, H*64, T) fp32 channel-major, batch stride qkv_bs; o (B, H*64, T), batch
 *                             stride o_bs; lse (B,H,T) log-sum-exp per query or NULL.
 * sty_attention64_tokens_fwd: qkv token-major rows [B*T, ld] (q | k | v at columns 0 / H*64 / 2*H*64); output bf16
 *                             hi | lo planes [2][M_pad][H*64] (rows >= B*T untouched). */
STY_API int64_t sty_attention64_workspace_bytes(int B, int H, int T);
STY_API int sty_attention64_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs, float* o, int64_t o_bs,
                                int B, int H, int T, float scale, float* lse, void* workspace, sty_stream_t stream);
/* backward of sty_attention64_fwd (needs its o and lse): two tcgen05 kernels over pre-split tiles — dQ (128 queries per
 * CTA, dQ accumulated in TMEM over all key tiles) and dK / dV (128 keys per CTA); dq, dk, dv (B, H*64, T) with batch
 * stride dqkv_bs.  workspace: sty_attention64_bwd_workspace_bytes(B,H,T), 16-byte aligned. */
STY_API int64_t sty_attention64_bwd_workspace_bytes(int B, int H, int T);
STY_API int sty_attention64_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                                const float* d_o, int64_t o_bs, const float* lse, float* dq, float* dk, float* dv,
                                int64_t dqkv_bs, int B, int H, int T, float scale, void* workspace, sty_stream_t stream);
STY_API int sty_attention64_tokens_fwd(const float* qkv, int64_t ld, void* out_split, int64_t M_pad, int B, int H, int T,
                                       float scale, void* workspace, sty_stream_t stream);

/* ---- spectrogram discriminator, non-tensor-core layers (discriminator.py:13-69) ---------------------------------
 * Row-channel images (B, Hp = bins + 2, C, W) with zero border rows, 32 hidden channels, LeakyReLU slope 0.1.
 * sty_disc_first_fwd  : h = Conv2d(1 -> 32, 3x9, pad (1,4))(y), y (B,bins,W) -> h (B,Hp,32,W) (border rows written 0);
 *                       w (32,1,3,9), bias (32) in the reference layout.  _dgrad: dy (B,bins,W) from dh (zero borders);
 *                       _wgrad: dw (32*27), db (32) (zeroed by the call, accumulated with atomics).
 * sty_disc_tail_fwd   : a = LeakyReLU(h); score (B,bins,W) = Conv2d(32 -> 1, 3x3, pad 1)(a) + b_score;
 *                       next_kind 0: nothing else | 1: next = a (B,Hp,32,W) | 2: next = space-to-depth of a along W,
 *                       (B,Hp,64,ceil(W/2)), channel 2c+p = a[c, 2u+p] (input of the following stride-(1,2) layer).
 * sty_disc_tail_bwd   : dh = LeakyReLU'(h) * (data gradient of the score conv from dscore (may be NULL) + d(next)),
 *                       border rows written 0.
 * sty_disc_score_wgrad: dw (32*9), db (1) of the score conv (zeroed by the call). */
STY_API int sty_disc_first_fwd(const float* y, const float* w, const float* bias, float* h, int B, int bins, int W,
                               sty_stream_t stream);
STY_API int sty_disc_first_dgrad(const float* dh, const float* w, float* dy, int B, int bins, int W, sty_stream_t stream);
STY_API int sty_disc_first_wgrad(const float* y, const float* dh, float* dw, float* db, int B, int bins, int W,
                                 sty_stream_t stream);
STY_API int sty_disc_tail_fwd(const float* h, const float* w_score, const float* b_score, float* score, float* next,
                              int next_kind, int B, int Hp, int W, sty_stream_t stream);
STY_API int sty_disc_tail_bwd(const float* h, const float* w_score, const float* dscore, const float* dnext, float* dh,
                              int next_kind, int B, int Hp, int W, sty_stream_t stream);
STY_API int sty_disc_score_wgrad(const float* h, const float* dscore, float* dw, float* db, int B, int Hp, int W,
                                 sty_stream_t stream);

/* ---- style-diffusion denoiser (BASELINE configs[3]; absent from the reference) ----
 * Token-major (rows = tokens) dense layers on TMA-fed tcgen05 GEMMs with bf16 hi|lo operand planes:
 * sty_split_planes_fwd : fp32 (n) -> bf16 planes out[0..n) = hi, out[n..2n) = lo   (weights, inputs)
 * sty_gemm_split_fwd   : C[M,N] = act(A W^T + bias) (+ res); A = planes [2][M][K], W = planes [2][N][K];
 *                        out fp32 (M,N) and / or out_split planes [2][M][N]; M % 128, N % 256, K % 64 == 0
 * sty_build_tokens_fwd : tok[b*T+t] = [scale * x[b] (Cx) | emb[b,t] (Ce)], rows >= B*T zero
 * sty_row_ln_split_fwd : hm = h + add[b] (written when hm != NULL); planes of LayerNorm_C(hm)*gamma+beta; C = 1024
 * sty_token_mean_fwd   : out[b] = mean_t x[b*T+t]
 * sty_attention_tokens_fwd : softmax(q k^T * scale) v per (batch, 64-wide head) on token-major q|k|v rows,
 *                        result as planes [2][M_pad][H*64] */
STY_API int sty_split_planes_fwd(const float* x, void* out, int64_t n, sty_stream_t stream);
STY_API int sty_gemm_split_fwd(const void* a_split, const void* w_split, const float* bias, const float* res,
                               float* out, void* out_split, int M, int N, int K, int act, sty_stream_t stream);
STY_API int sty_build_tokens_fwd(const float* x, const float* emb, float scale, float* tok, int B, int T, int Cx, int Ce,
                                 int M, sty_stream_t stream);
STY_API int sty_row_ln_split_fwd(const float* h, const float* add, const float* gamma, const float* beta, float eps,
                                 float* hm, void* out_split, int M, int M_real, int T, int C, sty_stream_t stream);
STY_API int sty_token_mean_fwd(const float* x, float* out, int B, int T, int C, sty_stream_t stream);
STY_API int sty_attention_tokens_fwd(const float* qkv, int64_t ld, void* out_split, int64_t M_pad, int B, int H, int T,
                                     float scale, sty_stream_t stream);

/* ---- LayerNorm over the channel axis of (B,C,T) --------------------------
 * v = x (+ res);  n = (v-mean_c)/sqrt(var_c+eps)
 * y = act( (g_plus_one ? 1+g : g) * n + b ) * mask[b,t]
 * g = gamma[b*g_bs + c], b = beta[b*g_bs + c]  (g_bs = 0: shared over the batch).
 * x and res share the batch stride x_bs; channel stride is T for x, res and y.
 * Replaces text_encoder.LayerNorm (text_encoder.py:24-33, eps 1e-4),
 * nn.LayerNorm via transposes (generator.py:756-778,887) and AdaptiveLayerNorm
 * (ada_norm.py:203-211, conformer.py:74-77,250). */
STY_API int sty_chan_layernorm_fwd(const float* x, const float* res, int64_t x_bs, const float* gamma,
                                   const float* beta, int64_t g_bs, int g_plus_one, float* y,
                                   int64_t y_bs, const float* mask, int B, int C, int T, float eps,
                                   int act, sty_stream_t stream);
/* same, explicit channel strides x_cs (x and res) / y_cs >= T (pitch-padded rows) */
STY_API int sty_chan_layernorm_pitched_fwd(const float* x, const float* res, int64_t x_bs, int64_t x_cs,
                                           const float* gamma, const float* beta, int64_t g_bs, int g_plus_one,
                                           float* y, int64_t y_bs, int64_t y_cs, const float* mask, int B, int C,
                                           int T, float eps, int act, sty_stream_t stream);

/* ---- InstanceNorm statistics folded with the style affine ----------------
 * mean/var over T per (b,c) (biased variance), then
 * scale[b,c] = (1+gamma[b,c]) / sqrt(var+eps),  shift[b,c] = beta[b,c] - mean*scale
 * so that AdaIN(x) = scale*x + shift, applied in the consumer conv's prologue.
 * gamma = gb[b*gb_bs + c], beta = gb[b*gb_bs + C + c].
 * Replaces AdaptiveInstance ada_norm.py:129-140. */
STY_API int sty_instnorm_affine_fwd(const float* x, int64_t x_bs, int64_t x_cs, const float* gb,
                            int64_t gb_bs, float* scale, float* shift, int B, int C, int T,
                            float eps, sty_stream_t stream);

/* ---- AdaIN affine from accumulated moments ------------------------------------------------------
 * mean = sum/T, var = max(sumsq/T - mean^2, 0) (biased); scale = (1+gamma)/sqrt(var+eps),
 * shift = beta - mean*scale, gamma/beta as in sty_instnorm_affine_fwd.  One-pass counterpart of
 * sty_instnorm_affine_fwd for tensors whose producer conv accumulated out_sum / out_sumsq. */
STY_API int sty_moments_affine_fwd(const float* sum, const float* sumsq, const float* gb, int64_t gb_bs,
                                   float* scale, float* shift, int B, int C, int T, float eps,
                                   sty_stream_t stream);

/* ---- small dense layer on vectors -----------------------------------------
 * out[b,j] = bias[j] + sum_i W[j,i] * s[b,i]      (all style FCs of a model packed
 * row-wise into one W: ada_norm.py:136,204 `self.fc(s)`). */
STY_API int sty_linear_rows_fwd(const float* s, const float* W, const float* bias, float* out, int B,
                        int I, int J, sty_stream_t stream);

/* ---- GRN scale -------------------------------------------------------------
 * gx[b,j] = sqrt(sumsq[b,j]); nx = gx / (mean_j gx + 1e-6)
 * scale[b,j] = 1 + gamma[j]*nx[b,j]      (conv_next.py:15-18: gamma*(x*nx)+beta+x;
 * the `beta` term is folded into the bias of the following pointwise conv). */
STY_API int sty_grn_scale_fwd(const float* sumsq, const float* gamma, float* scale, int B, int J,
                      sty_stream_t stream);

/* ---- rotary table ------------------------------------------------------------
 * cos/sin of t * 10000^(-2i/d_rot), i < d_rot/2, t < T  -> (T, d_rot/2) each.
 * text_encoder.py:112-137. */
STY_API int sty_rope_table(float* cos_out, float* sin_out, int T, int d_rot, float base,
                   sty_stream_t stream);

/* ---- multi-head attention core ----------------------------------------------
 * q,k,v: (B, H*D, T) channel-major (head h owns channels [h*D,(h+1)*D)).
 * o[b,h*D+j,t] = sum_u softmax_u( scale * <rope(q_t), rope(k_u)> + m(t,u) ) v[u][j]
 * m(t,u) = 0 if (t < len[b] and u < len[b]) else -1e4 (lengths==NULL: no mask).
 * rope tables (T, d_rot/2) or NULL.  Replaces arrange_heads + RoPE + mask +
 * F.scaled_dot_product_attention (text_encoder.py:233-272; conformer.py:112-131). */
STY_API int sty_attention_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs,
                      float* o, int64_t o_bs, const int64_t* lengths, const float* rope_cos,
                      const float* rope_sin, int d_rot, int B, int H, int D, int T, float scale,
                      sty_stream_t stream);

/* ---- multi-head attention core, any head size (D % 4 == 0, D <= 256) -------------------
 * Same semantics as sty_attention_fwd; q, k, v each have their own batch stride so that the
 * query and the key/value projections may come from different tensors
 * (DurationPredictor.compute_cross duration_predictor.py:58-67; ProsodyEncoder 2 heads x 160,
 * prosody_encoder.py:63-81). */
STY_API int sty_attention_generic_fwd(const float* q, int64_t q_bs, const float* k, const float* v,
                                      int64_t kv_bs, float* o, int64_t o_bs, const int64_t* lengths,
                                      const float* rope_cos, const float* rope_sin, int d_rot, int B,
                                      int H, int D, int T, float scale, sty_stream_t stream);

/* ---- duration head ------------------------------------------------------------------------
 * x: (B, NC, T) raw class scores.  out[b,t,:] = -|cumsum([x0, |x1|, |x2|, ...])| * (t < len[b])
 * as (B, T, NC).   duration_predictor.py:82-86 */
STY_API int sty_duration_head_fwd(const float* x, const int64_t* lengths, float* out, int B, int NC,
                                  int T, sty_stream_t stream);

/* ---- soft durations ----------------------------------------------------------------------------
 * pred (B,T,NC) -> dur[b,t] = sum_c softmax(pred)[c]*table[c] / (sum_c softmax + 1e-9) * (t < len[b]);
 * also total[0] = max_b round(sum_t dur[b,t]) (int32) — the frame count of the alignment.
 * DurationProcessor.prediction_to_duration / class_to_dur_soft utils.py:726-750,759 */
STY_API int sty_soft_duration_fwd(const float* pred, const int64_t* lengths, const float* table,
                                  float* dur, int32_t* total, int B, int T, int NC,
                                  sty_stream_t stream);

/* ---- soft alignment -------------------------------------------------------------------------------
 * duration (B,T) -> alignment (B,T,F): parabola window 1-(2x/(d+6))^2 around each token's centre,
 * kept on (lower-3, upper+3), clamped at 0, softmax over the TEXT axis (zeros outside the window
 * still receive weight e^0 — literal).  DurationProcessor.duration_to_alignment utils.py:752-791 */
STY_API int sty_alignment_fwd(const float* duration, float* alignment, int B, int T, int F,
                              sty_stream_t stream);

/* ---- batched matrix product ---------------------------------------------------
 * C[b] (M,N) = A[b] (M,K) @ Bm[b] (K,N), all row-major, batch strides in elements.
 * `text_encoding @ alignment` speech_predictor.py:60. */
STY_API int sty_bmm_fwd(const float* A, int64_t a_bs, const float* Bm, int64_t b_bs, float* C,
                int64_t c_bs, int B, int M, int N, int K, sty_stream_t stream);

/* ---- masked scale / row broadcast (glue of the predictors) ---------------------------------
 * scale_mask:     out[b,c,t] = x[b,c,t] * mask[b,t] * scale      (`x * x_mask`, prosody_encoder.py:70,79)
 * broadcast_rows: out[b,s,t] = v[b,s]  with batch stride out_bs  (style.unsqueeze(2).expand, :67) */
STY_API int sty_scale_mask_fwd(const float* x, const float* mask, float* out, int B, int C, int T,
                               float scale, sty_stream_t stream);
STY_API int sty_broadcast_rows_fwd(const float* v, float* out, int64_t out_bs, int B, int S, int T,
                                   sty_stream_t stream);

/* ---- gated linear unit over channels -------------------------------------------
 * y[b,c,t] = x[b,c,t] * sigmoid(x[b,c+C,t])   conformer.py:37-44 */
STY_API int sty_glu_fwd(const float* x, float* y, int B, int C, int T, sty_stream_t stream);

/* ---- harmonic source (SineGen + merge) --------------------------------------------
 * pitch, voiced: (B,F).  noise: (B, F*hop, H) standard normal draws (the reference
 * draws them with torch.randn, generator.py:440).  lin_w (H), lin_b (1).
 * work: (B,H,F) doubles scratch.  out: excitation (B, F*hop).
 * Replaces f0_upsamp + SourceModuleHnNSF + SineGen (generator.py:336-447,496-510,
 * 719-723).  The phase accumulation is carried in fp64 cycles. */
STY_API int sty_source_fwd(const float* pitch, const float* voiced, const float* noise,
                   const float* lin_w, const float* lin_b, double* work, float* out, int B,
                   int F, int hop, int H, float sample_rate, float sine_amp, float noise_std,
                   float voiced_threshold, sty_stream_t stream);

/* ---- conv-STFT of the excitation -------------------------------------------------------
 * wave (B,L) -> spec, phase (B, bins_keep, L/hop): replicate-pad n_fft/2, windowed DFT
 * with bases basis_re/basis_im (bins, n_fft), mag = sqrt(re^2+im^2+1e-14),
 * phase = atan2(im/mag, re/mag); the last frame and bins >= bins_keep are dropped.
 * stft.py:98-136 + generator.py:724-729. */
STY_API int sty_stft_fwd(const float* wave, const float* basis_re, const float* basis_im, float* spec,
                 float* phase, int B, int L, int n_fft, int hop, int bins_keep,
                 sty_stream_t stream);
/* same, outputs with explicit batch / channel strides (rows padded to 16 bytes so that the consumer convs can
 * fetch them with TMA) */
STY_API int sty_stft_pitched_fwd(const float* wave, const float* basis_re, const float* basis_im, float* spec,
                                 float* phase, int64_t out_bs, int64_t out_cs, int B, int L, int n_fft, int hop,
                                 int bins_keep, sty_stream_t stream);

/* ---- spectral head + conv-iSTFT + tanh ------------------------------------------------
 * logamp, real, imag: (B, bins, S) with batch strides logamp_bs / ri_bs (real and imag may be
 * the two halves of one fused conv output).  Per frame f in [0,S] (frame S replicates S-1):
 * mag = exp(logamp), ph = atan2(imag, real), re = mag*cos(ph), im = mag*sin(ph);
 * overlap-add with basis_re/basis_im (bins, n_fft) (already windowed and scaled),
 * wave = sum re*B_re - im*B_im, trimmed by n_fft/2 at both ends, then tanh.
 * out: (B, S*hop).  generator.py:782-799,896 + stft.py:138-187. */
STY_API int sty_istft_head_fwd(const float* logamp, int64_t logamp_bs, const float* real,
                               const float* imag, int64_t ri_bs, const float* basis_re,
                               const float* basis_im, float* out, int B, int S, int bins, int n_fft,
                               int hop, sty_stream_t stream);
/* same, inputs with explicit channel strides (pitch-padded rows) */
STY_API int sty_istft_head_pitched_fwd(const float* logamp, int64_t logamp_bs, int64_t logamp_cs, const float* real,
                                       const float* imag, int64_t ri_bs, int64_t ri_cs, const float* basis_re,
                                       const float* basis_im, float* out, int B, int S, int bins, int n_fft,
                                       int hop, sty_stream_t stream);

/* ---- framed real FFT front-end: STFT -> |X| / phase / mel / log, one kernel -------------------
 * torch.stft(center=True, pad_mode="reflect", onesided) semantics: frame f covers samples
 * [f*hop - n_fft/2, f*hop + n_fft/2) of the reflect-padded signal, times `window` (n_fft floats,
 * a shorter win_length already zero-padded to the centre as torch.stft does).
 *   mag   (B, K, n_frames), K = n_fft/2+1 : |X|^power                         (may be NULL)
 *   phase (B, K, n_frames)               : (|X| > phase_floor) * atan2(Im, Re) (may be NULL)
 *   mel   (B, n_mels, n_frames)          : post( sum_k fb[k,m] |X[k]|^power )  (may be NULL)
 *     mel_mode 0: raw   1: log1p(.)   2: (log(mel_eps + .) - mel_mean) / mel_std
 * The filterbank is sparse (triangles): filter m covers bins [fb_start[m], fb_start[m]+fb_len[m])
 * with weights fb_w[fb_off[m] ...]; `fbt_*` is the same matrix in CSR form by bin (backward only).
 * `twiddle`: n_fft/2 complex values exp(-2 pi i q / n_fft) as interleaved floats.
 * Replaces torchaudio MelSpectrogram + calculate_mel (train_context.py:155-169, utils.py:825-834)
 * and MultiSpectrogram.calculate_single (multi_spectrogram.py:40-55: torch.stft, abs, masked
 * angle, MelScale, log1p).  The backward accumulates into d_audio (which the caller zeroes):
 * d_audio[b, l] += d(mel)/d(audio) . d_mel + d(phase)/d(audio) . d_phase + d(mag)/d(audio) . d_mag. */
typedef struct sty_spectrogram_args {
  const float* audio;
  int64_t audio_bs;
  const float* window;
  const float* twiddle;
  float* mag;
  float* phase;
  float* mel;
  const int32_t* fb_start;
  const int32_t* fb_len;
  const int32_t* fb_off;
  const float* fb_w;
  const int32_t* fbt_ptr; /* (K+1) */
  const int32_t* fbt_mel;
  const float* fbt_w;
  int32_t B, L, n_fft, hop, n_frames, n_mels, power, mel_mode;
  float mel_eps, mel_mean, mel_std, phase_floor;
} sty_spectrogram_args;
STY_API int sty_spectrogram_fwd(const sty_spectrogram_args* a, sty_stream_t stream);
STY_API int sty_spectrogram_bwd(const sty_spectrogram_args* a, const float* d_mel, const float* d_phase,
                                const float* d_mag, float* d_audio, int64_t d_audio_bs,
                                sty_stream_t stream);

/* ---- log-energy of a normalised log-mel ------------------------------------------------------
 * out[b,f] = log( || exp(mel[b,:,f]*std + mean) ||_2 + 1e-9 )   (utils.py:73-85 log_norm/raw_energy,
 * stage_type.py:88-97) */
STY_API int sty_mel_energy_fwd(const float* mel, float* out, int B, int n_mels, int F, float mean,
                               float std, sty_stream_t stream);

/* ---- STFT losses ----------------------------------------------------------------------------------
 * l1_sums:    sums[0] += sum|t-p|, sums[1] += sum|t|   (spectral convergence, losses.py:27-28)
 * l1 bwd:     d_pred = coef[0] * sign(pred - target)   (coef on the device)
 * phase_loss: d = pred - target over (B,K,N), aw(x) = |x - 2 pi round(x / 2 pi)|, w_k = 2.5^(k/(K/2)):
 *             sums[0] += sum w_k aw(d), sums[1] += sum_{k<K-1} w_k aw(diff_k d), sums[2] += sum_{n<N-1}
 *             w_k aw(diff_n d)        (losses.py:41-84; the means are taken by the finalize step)
 * phase bwd:  d_pred of  coef[0] * (sums[0]/(BKN) + sums[1]/(B(K-1)N) + sums[2]/(BK(N-1)))
 * finalize:   out[0] = mel loss (mean over resolutions of sum|t-p|/(sum|t|+1e-6)), out[1] = multi-phase
 *             loss, out[2] = total = gm*out[0] + gp*out[1] with gm = w_mel/(out[0]+1e-9) when
 *             `normalize` (LossLog.backwards_loss loss_log.py:82-94) else w_mel (same for gp);
 *             out[4+r] = backward coefficient of l1 bwd at resolution r, out[4+n_res+r] = of phase bwd. */
STY_API int sty_l1_sums_fwd(const float* target, const float* pred, int64_t n, float* sums,
                            sty_stream_t stream);
STY_API int sty_l1_sums_bwd(const float* target, const float* pred, int64_t n, const float* coef,
                            float* d_pred, sty_stream_t stream);
STY_API int sty_phase_loss_fwd(const float* pred, const float* target, int B, int K, int N, float* sums,
                               sty_stream_t stream);
STY_API int sty_phase_loss_bwd(const float* pred, const float* target, int B, int K, int N,
                               const float* coef, float* d_pred, sty_stream_t stream);
STY_API int sty_stft_loss_finalize(const float* l1_sums, const float* phase_sums, const float* phase_counts,
                                   int n_res, float w_mel, float w_phase, int normalize, float* out,
                                   sty_stream_t stream);

/* =====================================================================================
 * Training path: backward kernels (the reference relies on ATen/cuDNN autograd formulas).
 * Data gradients of Conv1d reuse sty_conv1d_fwd with transposed, tap-reversed weights.
 * ===================================================================================== */

/* ---- dropout (training-mode stochastic regularisers) ---------------------------------------
 * Masks come from a stateless integer hash of (seed, site, element index) — see common.cuh drop_keep and its
 * numpy restatement oracle/dropout_oracle.py — so forward and backward regenerate the same mask and nothing
 * is stored.  `seed` is a DEVICE pointer to one uint64 read when the kernel runs (graph replays draw new masks);
 * `site` separates the dropout sites of one step.  NULL / p == 0 = no dropout.
 *   y[i]  = (res ? res[i] : 0) + scale * act(x[i]) * keep(i / group) / (1-p)      act: STY_ACT_NONE | STY_ACT_SWISH
 *   dx[i] = dy[i] * scale * act'(x[i]) * keep(i / group) / (1-p)
 * group = 1: nn.Dropout; group = T on (B,C,T): nn.Dropout1d; group = C*T: DropPath (conv_next.py:138-153).
 * Replaces nn.Dropout in ConvReluNorm / Encoder / FFN (text_encoder.py:63,204,323,352,388-391), FeedForward /
 * Attention / ConformerConvModule / ConformerBlock (conformer.py:90-92,144,187,245), AdaptiveDecoderBlock
 * (ada_norm.py:183-186), Dropout1d duration_predictor.py:79. */
typedef struct sty_dropout {
  const void* seed; /* device pointer to a uint64 */
  uint32_t site;
  float p;
} sty_dropout;
STY_API int sty_dropout_fwd(const float* x, const float* res, float* y, int64_t n, int64_t group, int act,
                            float scale, const sty_dropout* drop, sty_stream_t stream);
STY_API int sty_dropout_bwd(const float* x, const float* dy, float* dx, int64_t n, int64_t group, int act,
                            float scale, const sty_dropout* drop, sty_stream_t stream);
/* y[r,t] = act(scale[r]*x[r,t] + shift[r]) * keep(r*T+t)/(1-p), x (rows,T) contiguous: AdaIN affine + LeakyReLU +
 * Dropout ahead of the convs of AdaptiveDecoderBlock in train() mode (ada_norm.py:181-186).  Its backward is
 * sty_dropout_bwd (act NONE) followed by sty_prologue_bwd_reduce / _apply. */
STY_API int sty_affine_act_dropout_fwd(const float* x, const float* scale, const float* shift, float* y,
                                       int rows, int T, int act, const sty_dropout* drop, sty_stream_t stream);
/* attention with dropout on the probabilities (F.scaled_dot_product_attention(dropout_p=...),
 * text_encoder.py:270-275): element index of the mask = ((b*H + h)*T + query)*T + key. */
STY_API int sty_attention_drop_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs,
                                   float* o, int64_t o_bs, const int64_t* lengths, const float* rope_cos,
                                   const float* rope_sin, int d_rot, int B, int H, int D, int T, float scale,
                                   float* lse, const sty_dropout* drop, sty_stream_t stream);
STY_API int sty_attention_drop_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                                   const float* d_o, int64_t o_bs, const float* lse, const int64_t* lengths,
                                   const float* rope_cos, const float* rope_sin, int d_rot, float* dq, float* dk,
                                   float* dv, int64_t dqkv_bs, float* delta, int B, int H, int D, int T,
                                   float scale, const sty_dropout* drop, sty_stream_t stream);

/* ---- attention forward that also returns the row log-sum-exp (B,H,T) for the backward */
STY_API int sty_attention_lse_fwd(const float* q, const float* k, const float* v, int64_t qkv_bs,
                                  float* o, int64_t o_bs, const int64_t* lengths, const float* rope_cos,
                                  const float* rope_sin, int d_rot, int B, int H, int D, int T, float scale,
                                  float* lse, sty_stream_t stream);

/* ---- attention backward: dq, dk, dv (same layout / batch stride dqkv_bs) from dO; the
 * probabilities are recomputed from `lse`; `delta` (B,H,T) is scratch (= <dO, O> per row).
 * Backward of F.scaled_dot_product_attention + RoPE + mask (text_encoder.py:233-272,
 * conformer.py:112-131). */
STY_API int sty_attention_bwd(const float* q, const float* k, const float* v, int64_t qkv_bs, const float* o,
                              const float* d_o, int64_t o_bs, const float* lse, const int64_t* lengths,
                              const float* rope_cos, const float* rope_sin, int d_rot, float* dq, float* dk,
                              float* dv, int64_t dqkv_bs, float* delta, int B, int H, int D, int T, float scale,
                              sty_stream_t stream);

/* ---- attention backward for any head size through materialised probabilities (prosody encoder, 2 x 160) ----
 * heads_to_rows: x (B,H*D,T) [batch stride x_bs] -> y (B,H,T,D) * mul with RoPE on the first d_rot features;
 *                inverse != 0: x is (B,H,T,D), y the (B,H*D,T) tensor (batch stride x_bs), TRANSPOSED rotation.
 * attn_probs:    P[b,h,i,:] = softmax_j(<q_rows[i], k_rows[j]> + mask), mask as in sty_attention_fwd.
 * softmax_bwd:   dP <- P * (dP - sum_j P dP) row-wise, in place.
 * bmm_tn:        C[b] (M,N) = A[b] (K,M)^T @ Bm[b] (K,N).
 * Backward of text_encoder.py:233-272 as used by prosody_encoder.py:63-81 and duration_predictor.py:58-67. */
STY_API int sty_heads_to_rows(const float* x, int64_t x_bs, float* y, const float* rope_cos, const float* rope_sin,
                              int d_rot, int B, int H, int D, int T, float mul, int inverse, sty_stream_t stream);
STY_API int sty_attn_probs(const float* q_rows, const float* k_rows, const int64_t* lengths, float* P, int B, int H,
                           int D, int T, sty_stream_t stream);
STY_API int sty_softmax_bwd(const float* P, float* dP, int64_t rows, int T, sty_stream_t stream);
STY_API int sty_bmm_tn_fwd(const float* A, int64_t a_bs, const float* Bm, int64_t b_bs, float* C, int64_t c_bs,
                           int B, int M, int N, int K, sty_stream_t stream);

/* ---- Conv1d weight gradient ---------------------------------------------------------------
 * dw[co,ci,k] += sum_{b,t} g[b,co,t] * xin[b,ci,t + k*dil - pad],  g = out_scale*mask_o[b,t]*dy[b,co,t],
 * xin = the forward's prologue applied to x (see sty_conv1d_fwd).  dw is in the REFERENCE layout
 * (CO, CI, K) and is accumulated (the caller zeroes it).  K in {1,3,5,7,11,21}. */
typedef struct sty_conv1d_wgrad_args {
  const float* x;
  int64_t x_bs, x_cs;
  const float* dy;
  int64_t dy_bs, dy_cs;
  const float* in_scale;
  const float* in_shift;
  const float* in_alpha;
  const float* in_mask;
  const float* out_mask;
  float* dw;
  int32_t B, CI, CO, T, K, dil, pad, in_act;
  float out_scale;
  /* != 0: use the tcgen05 kernel (bf16x3 split precision, time as the contraction axis, taps folded
   * into the M side by an overlapping shared-memory descriptor) when CI % 8 == 0 and CO % 8 == 0;
   * 0 or ineligible shapes: fp32 FMA kernel (K in {1,3,5,7,11,21}). */
  int32_t tensor_cores;
} sty_conv1d_wgrad_args;
STY_API int sty_conv1d_wgrad(const sty_conv1d_wgrad_args* a, sty_stream_t stream);

/* out[c] += scale * sum_{b,t} mask[b,t]*x[b,c,t]      (bias gradients) */
STY_API int sty_channel_sum(const float* x, int64_t x_bs, int64_t x_cs, const float* mask, float* out, int B,
                            int C, int T, float scale, sty_stream_t stream);
/* out[r] = sum_t a[r,t]*b[r,t]                        (GRN: sum_t g*hb per (b,j)) */
STY_API int sty_row_dot(const float* a, const float* b, float* out, int64_t rows, int T, sty_stream_t stream);
/* mean[b,c], var[b,c] (biased) over T                 (InstanceNorm / BatchNorm statistics, training) */
STY_API int sty_row_moments(const float* x, int64_t x_bs, int64_t x_cs, float* mean, float* var, int B, int C,
                            int T, sty_stream_t stream);

/* ---- backward of the conv prologue  u = act(scale[b,c]*(x*mask) + shift[b,c]) ------------------------
 * given dxp = dL/du (B,C,T contiguous):  g_a = dxp*act'(a)
 * reduce: sums[b,c,:] = (sum_t g_a, sum_t g_a*(x*mask - center[b,c]), sum_t dxp * d snake/d alpha)
 * apply:  dx = g_a*scale*mask + c0[b,c] + c1[b,c]*x + add       (c0/c1: the statistics path of
 *         InstanceNorm / BatchNorm, ada_norm.py:129-140, conformer.py:183; add: a residual gradient) */
STY_API int sty_prologue_bwd_reduce(const float* dxp, const float* x, int64_t x_bs, int64_t x_cs,
                                    const float* scale, const float* shift, const float* alpha,
                                    const float* mask, const float* center, float* sums, int B, int C, int T,
                                    int act, sty_stream_t stream);
STY_API int sty_prologue_bwd_apply(const float* dxp, const float* x, int64_t x_bs, int64_t x_cs,
                                   const float* scale, const float* shift, const float* alpha,
                                   const float* mask, const float* c0, const float* c1, const float* add,
                                   int64_t add_bs, int64_t add_cs, float* dx, int64_t dx_bs, int64_t dx_cs,
                                   int B, int C, int T, int act, sty_stream_t stream);

/* ---- GRN + Snake backward (conv_next.py:15-18,86-88) ---------------------------------------------------
 * hb = snake(h; alpha_j); d_hb = g_u*gs[b,j] + kc[b,j]*hb; d_h = d_hb*(1+sin(2 alpha h));
 * dalpha[j] += sum_{b,t} d_hb * d snake/d alpha.   d_h may alias g_u. */
STY_API int sty_grn_snake_bwd(const float* g_u, const float* h, const float* gs, const float* kc,
                              const float* alpha, float* d_h, float* dalpha, int B, int J, int T,
                              sty_stream_t stream);
/* same with any activation `act` in place of Snake (alpha / dalpha unused unless act == STY_ACT_SNAKE):
 * AdaptiveConvNeXtBlock uses GELU (conv_next.py:96-141) */
STY_API int sty_grn_act_bwd(const float* g_u, const float* h, const float* gs, const float* kc, const float* alpha,
                            float* d_h, float* dalpha, int B, int J, int T, int act, sty_stream_t stream);

/* ---- backward of sty_chan_layernorm_fwd: dv (= dx = dres), dgb[b*dg_bs + c] += d gamma,
 * dgb[b*dg_bs + C + c] += d beta (dg_bs = 0: shared over the batch; the caller zeroes dgb). */
STY_API int sty_chan_layernorm_bwd(const float* x, const float* res, int64_t x_bs, const float* gamma,
                                   const float* beta, int64_t g_bs, int g_plus_one, const float* dy,
                                   const float* mask, float* dv, float* dgb, int64_t dg_bs, int B, int C, int T,
                                   float eps, int act, sty_stream_t stream);

/* ---- backward of the plain depthwise conv (no post-affine): dx (may be NULL), dw (C,K) and db (C)
 * accumulated (the caller zeroes them). */
STY_API int sty_dwconv1d_bwd(const float* dy, const float* x, int64_t x_bs, int64_t x_cs, const float* w,
                             float* dx, int64_t dx_bs, int64_t dx_cs, float* dw, float* db, int B, int C, int T,
                             int K, int pad_left, sty_stream_t stream);

STY_API int sty_glu_bwd(const float* x, const float* dy, float* dx, int B, int C, int T, sty_stream_t stream);
/* inverse of the conv epilogue's pixel shuffle: x[b, c*s+r, t] = y[b, c, t*s+r] */
STY_API int sty_unshuffle(const float* y, float* x, int B, int CO, int T, int s, sty_stream_t stream);
/* d_emb[tokens[b,t], c] += dx[b,c,t]*scale*(t < lengths[b])   (the caller zeroes d_emb) */
STY_API int sty_embed_bwd(const int64_t* tokens, const int64_t* lengths, const float* dx, float* d_emb, int B,
                          int T, int C, int n_tokens, float scale, sty_stream_t stream);
/* C[b] (M,N) = A[b] (M,K) @ Bt[b] (N,K)^T      (d text_encoding = d asr @ alignment^T) */
STY_API int sty_bmm_nt_fwd(const float* A, int64_t a_bs, const float* Bt, int64_t b_bs, float* C, int64_t c_bs,
                           int B, int M, int N, int K, sty_stream_t stream);
/* backward of sty_linear_rows_fwd: dW (J,I), dbias (J), ds (B,I) (ds may be NULL) */
STY_API int sty_linear_rows_bwd(const float* dh, const float* s, const float* W, float* dW, float* dbias,
                                float* ds, int B, int I, int J, sty_stream_t stream);
/* backward of sty_istft_head_fwd (out = its saved output): d_logamp (B,bins,S) contiguous; d_real / d_imag
 * (bins,S) planes with batch stride dri_bs (the two halves of the fused real|imag conv's gradient) */
STY_API int sty_istft_head_bwd(const float* dout, const float* out, const float* logamp, int64_t logamp_bs,
                               const float* real, const float* imag, int64_t ri_bs, const float* basis_re,
                               const float* basis_im, float* d_logamp, float* d_real, float* d_imag,
                               int64_t dri_bs, int B, int S, int bins, int n_fft, int hop, sty_stream_t stream);

/* ---- fused AdamW over a flat parameter arena (torch.optim.AdamW semantics, optimizers.py:106-117):
 * g is multiplied by grad_scale first (e.g. 1/world_size after a sum all-reduce). */
STY_API int sty_adamw_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                           float beta2, float eps, float weight_decay, int step, float grad_scale,
                           sty_stream_t stream);

/* =====================================================================================
 * Mel style encoder (mel_style_encoder.py:121-152): images are kept "row-channel",
 * X[b, r, c, w] contiguous (B, Hp = H+2, C, W) with zero border rows, so that an R x K Conv2d is
 * sty_conv1d_fwd on R consecutive rows seen as R*C stacked channels (overlapping strided view).
 * ===================================================================================== */
/* gradient of that overlapping view back to the image: dx[b,r,c,w] = sum_{k<R} g[b*Hp + r - k][k*C + c][w],
 * g: (B*Hp - (R-1), R*C, W) contiguous */
STY_API int sty_fold_rows(const float* g, float* dx, int R, int B, int Hp, int C, int W, sty_stream_t stream);
/* learned downsampling: depthwise Conv2d 3x3, stride 2, padding 1 (mel_style_encoder.py:28-38);
 * y: (B, Ho+2, C, Wo), Ho = (H-1)/2+1, Wo = (W-1)/2+1; w (C,9), bias (C) or NULL.
 * bwd: dx (may be NULL), dw (C,9) and db (C, may be NULL) accumulated (the caller zeroes them). */
STY_API int sty_dwconv3x3s2_fwd(const float* x, const float* w, const float* bias, float* y, int B, int Hp, int C,
                                int W, sty_stream_t stream);
STY_API int sty_dwconv3x3s2_bwd(const float* dy, const float* x, const float* w, float* dx, float* dw, float* db,
                                int B, int Hp, int C, int W, sty_stream_t stream);
/* F.avg_pool2d(x, 2) with the last column replicated when W is odd (mel_style_encoder.py:57-60); H even */
STY_API int sty_avgpool2_fwd(const float* x, float* y, int B, int Hp, int C, int W, sty_stream_t stream);
STY_API int sty_avgpool2_bwd(const float* dy, float* dx, int B, int Hp, int C, int W, sty_stream_t stream);
/* mean over rows [r0,r0+Rn) x columns [w0,w0+Wn): the valid region of the 5x5 conv followed by
 * AdaptiveAvgPool2d(1) (mel_style_encoder.py:140-141) -> (B,C); bwd fills dx (B,Hp,C,W) */
STY_API int sty_region_mean_fwd(const float* x, float* out, int B, int Hp, int C, int W, int r0, int Rn, int w0,
                                int Wn, sty_stream_t stream);
STY_API int sty_region_mean_bwd(const float* g, float* dx, int B, int Hp, int C, int W, int r0, int Rn, int w0,
                                int Wn, sty_stream_t stream);
/* same, with hyper = [learning rate, step count (>= 1)] in DEVICE memory: the launch can be captured in a CUDA
 * graph and replayed while the schedule / bias correction advance */
STY_API int sty_adamw_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* hyper,
                               float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                               sty_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* STYLISH_B200_H */
