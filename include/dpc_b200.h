/* dpc_b200.h — C-ABI of the B200-native DiffPhyCon sampling hot path (libdpc_b200.so).
 *
 * The reference has no FFI: its boundary for this path is the Python class surface
 * (SURVEY.md section 8(b)).  This library is what a maintainer would bind from those classes with ctypes
 * (see INTEGRATION.md).  Every entry point takes plain device pointers + sizes + a CUDA stream handle
 * (cudaStream_t passed as void*); there are no torch types in any signature.
 *
 * Conventions
 *   - all activation tensors are fp32.  "channels-last" means [B, F, H, W, C] (C contiguous); "reference layout"
 *     means the reference's [B, F, C, H, W] (diffusion_2d_smoke.py:703, conv3d.py:486).
 *   - every function returns 0 on success, otherwise a cudaError_t value (or -1 for an argument error);
 *     dpc_last_error() returns a static string describing the last failure on the calling thread.
 *   - nothing here synchronises the device; all work is enqueued on `stream`.
 *
 * Reference paths are relative to the reference repository root; "conv3d.py" abbreviates
 * model/video_diffusion_pytorch/video_diffusion_pytorch_conv3d.py and "smoke.py" diffusion/diffusion_2d_smoke.py.
 */
#ifndef DPC_B200_H
#define DPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPC_ABI_VERSION 1

int dpc_abi_version(void);
const char* dpc_last_error(void);
/* 1 if the current device is sm_100 (B200), else 0; negative on CUDA error. */
int dpc_device_is_sm100(void);

/* ---------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer on tensor cores (TF32 inputs, fp32 accumulate).
 * Replaces nn.Conv3d / nn.ConvTranspose3d / nn.Linear / 1x1 nn.Conv2d forward calls of
 * conv3d.py:159-163 (Up/Downsample), :189-204 (Block.proj), :214 (res_conv), :240-241, :288-289 (to_qkv/to_out),
 * :403 (init_conv), :471 (final 1x1x1).
 *
 * Input  : x1 [B,Fi,Hi,Wi,C1] channels-last, optionally concatenated along C with x2 [B,Fi,Hi,Wi,C2]
 *          (replaces torch.cat at conv3d.py:538, :545).
 * Weights: packed K-major [Npad][Kpad], k = tap*(C1+C2) + ci, tap = (dt*kh + dh)*kw + dw, zero padded
 *          (packing is done by the host mirror, diffphycon_b200/packing.py).
 * taps   : device int32 [ntaps][4] = (dt, dh, dw, (dt*Hi + dh)*Wi + dw).
 * Output : row m = ((b*Fo + fo)*Ho + ho)*Wo + wo is written to
 *          channels-last row ((b*Fo + fo)*Hfull + ho*oh_mul + oh_off)*Wfull + wo*ow_mul + ow_off   (out_layout 0)
 *          or to reference layout [B,Fo,Cout,Hfull,Wfull]                                            (out_layout 1).
 *          oh_mul/ow_mul = 2 express one parity class of ConvTranspose3d(1,4,4; stride (1,2,2); pad (0,1,1)).
 * Epilogue: + bias[Cout], + residual (same indexing as the output, channels-last only),
 *          GroupNorm partial statistics: gn_stats[b][g][0..1] += (sum, sum of squares) over the written values,
 *          g = n / (Cout / gn_groups)   (feeds dpc_groupnorm_silu; replaces the statistics pass of nn.GroupNorm).
 * ------------------------------------------------------------------------------------------------------- */
typedef struct dpc_conv_params {
  const float* x1; const float* x2;
  const float* w; const float* bias; const float* residual;
  const int32_t* taps;
  float* y;
  double* gn_stats;
  int32_t C1, C2;
  int32_t B, Fi, Hi, Wi;
  int32_t Fo, Ho, Wo;
  int32_t ntaps;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;
  int32_t Cout, Npad, Kpad;
  int32_t Hfull, Wfull, oh_mul, oh_off, ow_mul, ow_off;
  int32_t out_layout;
  int32_t gn_groups;
  int32_t precise;   /* 1 = 3xTF32 error-compensated product (near-fp32), 0 = plain TF32 */
  /* optional (dpc_conv3d_tcgen05, 1x1x1 only): the residual enters as silu(residual * res_scale[b][c] + res_shift[b][c]), i.e. a
   * GroupNorm-apply + SiLU folded into per-(sample, channel) coefficients by dpc_gn_fold: ResnetBlock's
   * block2-norm-act + res_conv(x) (conv3d.py:197-204, :229-230) in one pass.  [B][Cout] each; NULL = plain residual. */
  const float* res_scale; const float* res_shift;
} dpc_conv_params;

int dpc_conv_igemm(const dpc_conv_params* p, void* stream);

/* Conv3d 3x3x3 / pad 1, 1x1x1 (= Linear over channels-last rows), the 1x4x4 stride-(1,2,2) down-conv and one parity class of
 * the 1x4x4 stride-(1,2,2) ConvTranspose on the 5th-generation tensor cores: persistent CTAs, TMA-tiled operand staging
 * (3x3x3: one halo box per (dt, 32-channel chunk) shared by all nine in-plane taps; down-conv: four parity boxes fetched with
 * TMA traversal stride 2), tcgen05.mma kind::tf32, accumulators in TMEM (double-buffered when they fit).  When B*F is even the
 * 3x3x3 path runs as cta_group::2 clusters (two consecutive frames per CTA pair, M = 256 MMAs; DPC_TC_PAIR=0 disables).
 * Same arguments, weight packing and epilogue as dpc_conv_igemm (bias, residual for 1x1x1, GroupNorm(8) partial statistics
 * for 3x3x3).
 * Supported: channels-last output, C1 % 32 == 0, C2 % 32 == 0, 4 <= W <= 254, precise == 0, and
 *   ntaps == 27: unit stride, pad 1, no residual, Cout in {64,128,256,512} == Npad (512 = two 256-column launches),
 *                gn_groups in {0, 8};
 *   ntaps == 9 : the same kernel for 3x3 convolutions over IMAGES (Fi = Fo = 1, pt = 0, ph = pw = 1) — the 2-D networks of
 *                diffusion/diffusion_2d_jellyfish.py:189-204 and their dgrad convolutions;
 *   ntaps == 1 : unit stride, pad 0, no gn_stats, Cout == 64 or Cout % 128 == 0 (column tiles of 64/128/256, up to 512
 *                outputs per pass over the input);
 *   ntaps == 16: kernel 1x4x4, stride (1,2,2), pad (0,1,1), C2 == 0, Cout in {64,128,256} == Npad, no residual / gn_stats;
 *   ntaps == 4 : kernel 1x2x2 of ConvTranspose parity class (oh_off, ow_off), ph == 1-oh_off, pw == 1-ow_off,
 *                oh_mul == ow_mul == 2, same channel constraints.
 * Returns -2 (nothing launched) for any other shape so the host mirror can use dpc_conv_igemm (same numerics class). */
int dpc_conv3d_tcgen05(const dpc_conv_params* p, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * GroupNorm-apply + (scale+1, shift) + SiLU (+ residual) — conv3d.py:197-204 (Block.forward after proj),
 * :229-230 (ResnetBlock residual add).  y is the raw conv output [B, rows_per_sample, C] channels-last,
 * stats is the [B,G,2] double buffer filled by the conv epilogue.  scale_shift (nullable) is [B, ss_stride]
 * with scale at [ss_off .. ss_off+C) and shift at [ss_off+C .. ss_off+2C)   (conv3d.py:222-225).
 * out = silu(((y-mean)*rstd*gamma + beta) * (scale+1) + shift) + residual
 * ------------------------------------------------------------------------------------------------------- */
int dpc_groupnorm_silu(const float* y, const double* stats, const float* gamma, const float* beta,
                       const float* scale_shift, int64_t ss_stride, int64_t ss_off,
                       const float* residual, float* out,
                       int32_t B, int64_t rows_per_sample, int32_t C, int32_t groups, float eps, void* stream);

/* Channel LayerNorm, gain only, over C for every row (+ optional residual added after the gain):
 * use_rsqrt 0: (x-mean)/sqrt(var+eps)*gamma  — conv3d.py:165-174;
 * use_rsqrt 1: (x-mean)*rsqrt(var+eps)*gamma — the 2-D variants, model/burgers_1d/unet.py:60-70. */
/* Folds GroupNorm statistics into per-(sample, channel) affine coefficients: scale = rstd*gamma, shift = beta - mean*rstd*gamma
 * (stats as produced by the conv epilogues: [B][groups][2] doubles = sum, sum of squares over rows_per_sample*C/groups values). */
int dpc_gn_fold(const double* stats, const float* gamma, const float* beta, float* scale, float* shift, int32_t B,
                int64_t rows_per_sample, int32_t C, int32_t groups, float eps, void* stream);
/* GroupNorm statistics of a coarser grouping from a finer one: stats_out[b][g] += sum of the groups_in / groups_out consecutive
 * entries of stats_in (sum and sum of squares are additive over channels).  The tcgen05 conv epilogue produces GroupNorm(8)
 * statistics; the Burgers U-Nets use GroupNorm(1) blocks (model/burgers_1d/unet.py:95-111 with resnet_block_groups = 1). */
int dpc_gn_stats_merge(const double* stats_in, double* stats_out, int32_t B, int32_t groups_in, int32_t groups_out, void* stream);
int dpc_layernorm_channels(const float* x, const float* gamma, const float* residual, float* out, int64_t rows,
                           int32_t C, float eps, int32_t use_rsqrt, void* stream);

/* nn.Upsample(scale_factor=2, mode='nearest') on channels-last [BF,H,W,C] -> [BF,2H,2W,C] (model/burgers_1d/unet.py:40-44). */
int dpc_upsample_nearest2x(const float* x, float* out, int64_t BF, int32_t H, int32_t W, int32_t C, void* stream);

/* final_conv[1] — conv3d.py:427, Conv3d(dim, out_dim, 1) — for dim == 64 and out_dim in {2, 4, 6}: x [BF*HW][64] channels-last
 * -> out [BF][out_dim][HW] (the reference layout), w [out_dim][64] fp32, bias [out_dim] or NULL; fp32 FMAs.  Returns -2 for
 * other shapes (served by dpc_conv_igemm with out_layout = 1). */
int dpc_final_proj(const float* x, const float* w, const float* bias, float* out, int64_t BF, int32_t HW, int32_t C,
                   int32_t Cout, void* stream);
/* Reference layout [B,F,Ctot,H,W], channels [c0, c0+Cin) -> channels-last [B,F,H,W,Cpad] zero padded
 * (replaces the permute at conv3d.py:495 and the slice x[:, :, 3:5] at smoke.py:612). */
int dpc_pack_input(const float* x, float* out, int32_t B, int32_t F, int32_t Ctot, int32_t c0, int32_t Cin,
                   int32_t H, int32_t W, int32_t Cpad, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Attention blocks.  qkv is [rows, 3*heads*32] channels-last with q | k | v thirds, head-major inside each third
 * (conv3d.py:246, :311); out is [rows, heads*32].  dim_head is fixed at 32 (conv3d.py:362).
 * ------------------------------------------------------------------------------------------------------- */
/* Temporal softmax attention over frames per pixel with RoPE on q,k and T5 relative bias — conv3d.py:293-352,
 * rotary-embedding-torch 0.8.4 rotate_queries_or_keys.  rope_cos/rope_sin: [F][32] (angle table, interleaved pairs);
 * pos_bias: [heads][F][F] or NULL; use_rope bit 0: rotate q,k (0 skips the rotation); bit 1: pos_bias[h][i][j] depends on j - i only
 * (RelativePositionBias, conv3d.py:110-148) and is read through a [heads][2F-1] shared-memory table. F <= 64.  The two contractions run on tensor cores
 * (mma.sync m16n8k8; precise 0: TF32 operands, 1: 3xTF32 split products, fp32-class like the reference's einsum); frames are
 * padded to 32 (F <= 32) or 64 (32 < F <= 64) with the padded keys masked out. */
int dpc_temporal_attention(const float* qkv, const float* rope_cos, const float* rope_sin, const float* pos_bias,
                           float* out, int32_t B, int32_t F, int32_t HW, int32_t heads, int32_t use_rope,
                           int32_t precise, void* stream);
/* The whole temporal-attention residual block in one launch (dim 64, 4 heads, F = 32, HW % 4 == 0; returns -2 otherwise and
 * the caller uses the unfused entry points):  y = x + to_out(Attention(LayerNorm(x)))  — conv3d.py:165-174 (LayerNorm, gain
 * only), :293-352 (to_qkv, scale, RoPE, relative bias, softmax, to_out without bias), :153-157 (Residual).
 * x, y: [B, F, HW, 64] channels-last.  w_qkv: [384][64] = to_qkv.weight * LayerNorm gain (gain folded in), TF32-rounded;
 * w_out: [64][128] = to_out.weight, TF32-rounded.  pos_bias [4][32][32] must be a function of (j - i) only (T5 relative
 * bias, conv3d.py:74-112): the kernel reads its first row and first column.  Contractions: TF32 tcgen05 MMAs for the two
 * projections, fp32 for q k^T and P v. */
int dpc_temporal_block_fused(const float* x, const float* w_qkv, const float* w_out, const float* rope_cos,
                             const float* rope_sin, const float* pos_bias, float* y, int32_t B, int32_t F, int32_t HW,
                             int32_t C, int32_t heads, float eps, void* stream);
/* Softmax attention over the HW tokens of every frame (mid block) — conv3d.py:449-451 with Attention(:293-352),
 * no RoPE, no bias. */
int dpc_spatial_attention(const float* qkv, float* out, int32_t BF, int32_t HW, int32_t heads, void* stream);
/* Same attention with q k^T and P v on tensor cores (TF32 mma.sync, flash-attention style online softmax, fp32 accumulate);
 * HW % 64 == 0, returns -2 otherwise. */
int dpc_spatial_attention_mma(const float* qkv, float* out, int32_t BF, int32_t HW, int32_t heads, void* stream);
/* Spatial linear attention — conv3d.py:243-257: q softmax over d, k softmax over pixels, v NOT divided by HW.
 * ctx_ws: workspace of BF*heads*32*32 floats. */
int dpc_spatial_linear_attention(const float* qkv, float* ctx_ws, float* out, int32_t BF, int32_t HW, int32_t heads,
                                 void* stream);

/* Stem convolution init_conv — conv3d.py:392, Conv3d(channels, dim, (kt, kh, 7), padding (kt/2, kh/2, 3)) — on tcgen05 with
 * sliding-window (no-swizzle, overlapping core matrix) operand descriptors over the raw channels-last input.
 * x: [B,F,H,W,Cpad] (Cpad % 4 == 0, <= 16; padded channels zero), y: [B,F,H,W,N] channels-last, N in {32, 64, 128}.
 * w: [N][kt*kh*(Cpad/4)*32] with k = (((dt*kh + dh)*(Cpad/4) + plane)*8 + dw)*4 + c, channel = 4*plane + c, the dw = 7 column
 * zero (diffphycon_b200.packing.pack_stem_conv).  TF32 operands, fp32 accumulate.  Returns -2 for shapes it does not serve. */
int dpc_stem_conv_tcgen05(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t F, int32_t H,
                          int32_t W, int32_t Cpad, int32_t N, int32_t kt, int32_t kh, int32_t kw, void* stream);
/* The whole spatial-linear-attention residual block (dim 64, 4 heads, HW % 128 == 0; returns -2 otherwise and the caller
 * uses the unfused entry points):  y = x + to_out(SpatialLinearAttention(LayerNorm(x))) + b  — conv3d.py:165-174 (LayerNorm,
 * gain only), :232-257 (to_qkv, softmax of q over d, softmax of k over the pixels, context, to_out), :153-157 (Residual).
 * x, y: [BF, HW, 64] channels-last.  w_qkv: [384][64] = to_qkv.weight * LayerNorm gain, TF32-rounded; w_out: [64][128]
 * = to_out.weight; b_out: [64] or NULL.  Workspaces: ctx_ws BF*4*32*32 floats (receives the normalised, scaled context),
 * mt_ws BF*64*128 floats (context folded with to_out).  Three launches; contractions are TF32 tcgen05 MMAs, softmaxes and
 * the context normalisation fp32. */
int dpc_spatial_linear_block_fused(const float* x, const float* w_qkv, const float* w_out, const float* b_out,
                                   float* ctx_ws, float* mt_ws, float* y, int32_t BF, int32_t HW, int32_t C,
                                   int32_t heads, float eps, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Time embedding — conv3d.py:139-151 (SinusoidalPosEmb), :404-409 (time_mlp), :211-214, :222-224 (ResnetBlock.mlp).
 * freqs: [dim/2] table exp(-i*log(10000)/(dim/2-1)).  t_emb out: [B, 4*dim].
 * dpc_time_proj: out[b, j] = bias[j] + sum_k W[j,k] * silu(t_emb[b,k]) for the row-concatenated mlp weights of all
 * ResnetBlocks (W: [total, tdim]).
 * ------------------------------------------------------------------------------------------------------- */
int dpc_time_embed(const int64_t* t, const float* freqs, const float* w1, const float* b1, const float* w2,
                   const float* b2, float* hidden_ws, float* t_emb, int32_t B, int32_t dim, void* stream);
int dpc_time_proj(const float* t_emb, const float* W, const float* bias, float* out, int32_t B, int32_t tdim,
                  int32_t total, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Fused guidance + posterior update of one denoising step, reference layout [B,F,6,H,W].
 * smoke.py:610-656 (model_predictions), :659-666 (p_mean_variance), :671-699 (p_sample), :720 (re-impose init),
 * with the stock guidance of inference/inference_2d_smoke.py:30-44 in closed form.
 * ------------------------------------------------------------------------------------------------------- */
typedef struct dpc_step_coefs {
  float sqrt_recip_alphas_cumprod;    /* extract(sqrt_recip_alphas_cumprod, t) */
  float sqrt_recipm1_alphas_cumprod;
  float guidance_coef;                /* standard_fixed_ratio, or coeff_ratio*flip(betas)[t] ("standard-alpha") */
  float prior_coef;                   /* w_prob_exp - 1 */
  float w_energy;                     /* stock guidance_fn argument */
  float rescaler[6];                  /* dataset/data_2d.py:167 */
  /* DDPM posterior (used by dpc_ddpm_guided_step) */
  float posterior_mean_coef1, posterior_mean_coef2, sigma;   /* sigma = exp(0.5*posterior_log_variance_clipped) */
  int32_t add_noise;                  /* t > 0 */
  /* DDIM (used by dpc_ddim_guided_step) */
  float sqrt_alpha_next, c, ddim_sigma;
  int32_t last;                       /* time_next < 0: return x_start */
} dpc_step_coefs;

/* eps_joint [B,F,6,H,W]; eps_w [B,F,2,H,W] (scattered into channels 3:5); noise like x (may be NULL when
 * add_noise == 0); init [B,H,W] or NULL (NULL: do not re-impose x[:,0,0]); x_out like x (may alias x);
 * x_start_out nullable.
 * use_stock_guidance 0: `g` [B,F,6,H,W] is a user-supplied design_fn gradient evaluated on x_start (see
 * dpc_predict_x_start). */
int dpc_ddpm_guided_step(const float* x, const float* eps_joint, const float* eps_w, const float* noise,
                         const float* init, const float* g, int32_t use_stock_guidance,
                         const dpc_step_coefs* coefs, float* x_out, float* x_start_out,
                         int32_t B, int32_t F, int32_t H, int32_t W, void* stream);
int dpc_ddim_guided_step(const float* x, const float* eps_joint, const float* eps_w, const float* noise,
                         const float* init, const float* g, int32_t use_stock_guidance,
                         const dpc_step_coefs* coefs, float* x_out, float* x_start_out,
                         int32_t B, int32_t F, int32_t H, int32_t W, void* stream);
/* Table-driven form of the two entry points above for a step captured ONCE in a CUDA graph and replayed for the whole
 * schedule: dpc_sampler_prepare reads the device-side step counter *step_index, writes the batched time
 * tensor tt[0..B) = t_table[i] (the U-Nets' `time` argument) and *cur = c_table[i], then increments the counter;
 * dpc_guided_step_dev is dpc_ddpm/ddim_guided_step with the stock guidance, taking the per-step coefficients from device
 * memory (`coefs_dev` = cur).  `coefs_host` supplies what does not change from step to step (w_energy, rescaler).  `noise` must
 * be non-NULL; it is ignored when the step's coefficients say so (add_noise == 0 / last != 0).  x_out may alias x. */
int dpc_sampler_prepare(const int64_t* t_table, const dpc_step_coefs* c_table, int32_t* step_index, int32_t nsteps, int64_t* tt,
                        int32_t B, dpc_step_coefs* cur, void* stream);
int dpc_guided_step_dev(int32_t ddim, const float* x, const float* eps_joint, const float* eps_w, const float* noise,
                        const float* init, const dpc_step_coefs* coefs_host, const dpc_step_coefs* coefs_dev, float* x_out,
                        float* x_start_out, int32_t B, int32_t F, int32_t H, int32_t W, void* stream);
/* x_start = maybe_clip(sqrt_recip*x - sqrt_recipm1*eps) — smoke.py:576-580, :620-621; for user design_fn callables. */
/* Self-recurrence re-noising of the Burgers sampler — diffusion_1d_burgers.py:472-482 (recurrent_sample):
 * out = a * x + b * z with a = sqrt(alpha_t / alpha_{t-1}), b = sqrt(1 - alpha_t / alpha_{t-1}) (products and sum rounded like the
 * reference's tensor ops); z NULL at t == 0 (no noise).  out may alias x. */
int dpc_renoise(const float* x, const float* z, float a, float b, float* out, int64_t n, void* stream);
int dpc_predict_x_start(const float* x, const float* eps, float sqrt_recip, float sqrt_recipm1, int32_t clip,
                        float* x_start, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Post-sampling smoke rollout — dataset/apps/evaluate_solver.py:205-310 (solver), :118-147 (get_envolve) with
 * phi/flow.py:294-327, phi/math/nd.py:332-427, :602-614, phi/math/scipy_backend.py:58-77, :181-185,
 * phi/solver/sparse.py:27-119 and phi/solver/base.py:56-103 (plain CG, max|r| >= accuracy, <= max_iterations).
 * One launch runs all T-1 simulation steps in fp64 (the reference's NumPy precision); a thread-block cluster of 4 (while every
 * trajectory runs in one wave) or 2 CTAs per trajectory keeps the CG vectors on chip (x_ws is unused since round 2).
 * fluid_mask [127][127] int8 (1 fluid / 0 obstacle), velocity_mask [128][128][2] fp32 (component 0 = x faces),
 * init_velocity [B][128][128][2], init_density [B][nx][nx], c1/c2 [B][nt][nx][nx] (tiled in space and time on the fly,
 * es.py:223-227).  Workspaces: vel_ws B*2*128*128*2 doubles, x_ws B*127*127 doubles, dens_ws B*4*127*127 floats.
 * Outputs: densitys / zero_densitys [B][T][128][128] fp32, velocitys [B][T][128][128][2] fp64, smoke_out [B][T] fp64
 * (es.py:305), iterations [B][T] int32 (CG iterations of the step that produced frame t).
 * ------------------------------------------------------------------------------------------------------- */
int dpc_smoke_rollout(const int8_t* fluid_mask, const float* velocity_mask, const float* init_velocity,
                      const float* init_density, const float* c1, const float* c2, double* vel_ws, double* x_ws,
                      float* dens_ws, float* densitys, float* zero_densitys, double* velocitys, double* smoke_out,
                      int32_t* iterations, int32_t B, int32_t nt, int32_t nx, int32_t T, double dt, double accuracy,
                      int32_t max_iterations, void* stream);

/* Evaluation-stage sums of InferencePipeline.multi_evaluate (inference/inference_2d_smoke.py:384-416) taken from the rollout
 * outputs: pred [B][F][6][S][S] against data_current = (density, velocity x/y, masked sampled controls, smoke portion) at frames
 * t = f*T/F and pixels (y*128/S, x*128/S), frame 0 excluded.  The sampled controls count as zero inside [mask_lo, mask_hi)^2
 * (:322).  sums [B][12] doubles: 0..5 sum (pred_c - data_c)^2; 6..10 sum data_c^2 (c = 0..4); 11 sum of pred[:, F-1, 5]. */
int dpc_smoke_eval_sums(const float* pred, const float* densitys, const double* velocitys, const double* smoke_out,
                        double* sums, int32_t B, int32_t F, int32_t S, int32_t T, int32_t mask_lo, int32_t mask_hi, void* stream);

/* Burgers sampler, elementwise parts of diffusion/diffusion_1d_burgers.py:396-470 on [B,C,H,W] tensors (n elements,
 * `plane` = H*W): dpc_burgers_model_output combines the joint and prior network outputs (mode 0: eps1 - coef*eps2',
 * mode 1: (eps1 - coef*eps2')/beta, mode 2: (beta*eps1)'; ' zeroes channel 0, :403, :414) and predicts x_start (:425);
 * dpc_ddpm_posterior_step adds the guidance (eps + g*gscale, :431-434), re-predicts and optionally clamps x_start (:456-458)
 * and applies the posterior (:381-389, :464-470).  g / noise / x_start_out / pred_noise_out may be NULL. */
int dpc_burgers_model_output(const float* x, const float* eps1, const float* eps2, float* out, float* x_start, int32_t mode,
                             float coef, float beta, float sqrt_recip, float sqrt_recipm1, int32_t C, int64_t plane,
                             int64_t n, void* stream);
int dpc_ddpm_posterior_step(const float* x, const float* eps, const float* g, const float* noise, float* x_out,
                            float* x_start_out, float* pred_noise_out, float gscale, float sqrt_recip, float sqrt_recipm1,
                            int32_t clip, float coef1, float coef2, float sigma, int64_t n, void* stream);

/* Jellyfish sampler step, diffusion/diffusion_2d_jellyfish.py (jf.py).  State x [B,F,7,H,W] = [state 3, boundary 3, theta 1],
 * P = H*W; the diffused channels are {0,1,2,6}: eps / x_start / g / noise are [B,F,4,P], eps_w [B,F,1,P].
 * dpc_jelly_x_start: x_start = clamp?(sqrt_recip*x4 - sqrt_recipm1*eps) (jf.py:714, :744, :764).
 * dpc_jelly_step, one launch per sampling step: DDPM (ddim=0) pred = c1*x_start + c2*x4 + sigma*noise (jf.py:601-604, :789)
 *   then pred -= ga*g - gb*eps_w with eps_w broadcast over the 4 channels (jf.py:798-804); DDIM (ddim=1) pred_noise = eps +
 *   ga*g - gb*pad(eps_w) (eps_w on the theta channel only, jf.py:728-742), pred = c1*x_start + c2*pred_noise + sigma*noise
 *   (jf.py:925-927).  Then dtheta[b,f] = mean_hw(theta) - thetas_0[b] (update_bd, jf.py:810-814), the conditions of
 *   jf.py:858-864 (frames < cond_steps: state_0 and thetas_0; frames >= F-cond_steps: thetas_0), theta_mean [B,F] of the
 *   conditioned theta (jf.py:876); writes channels 0..2 and 6 of x_next and channel 6 of x_w (the prior model's input).
 *   g and noise may be NULL (no design_fn / t == 0); eps is read only when ddim=1.
 * dpc_jelly_write_bd: channels 3..5 of x_next and x_w <- pred_bd [B*F,3,P], frames < cond_steps and >= F-cond_steps <- bd_0
 *   [B,3,P] (jf.py:860-861). */
int dpc_jelly_x_start(const float* x, const float* eps, float* x_start, float sqrt_recip, float sqrt_recipm1, int32_t clip,
                      int64_t BF, int64_t P, void* stream);
int dpc_jelly_step(const float* x, const float* x_start, const float* eps, const float* eps_w, const float* g,
                   const float* noise, const float* state_0, const float* thetas_0, float* x_next, float* x_w, float* dtheta,
                   float* theta_mean, float ga, float gb, float c1, float c2, float sigma, int32_t ddim, int32_t B, int32_t F,
                   int32_t cond_steps, int64_t P, void* stream);
int dpc_jelly_write_bd(const float* pred_bd, const float* bd_0, float* x_next, float* x_w, int32_t B, int32_t F,
                       int32_t cond_steps, int64_t P, void* stream);

/* Burgers finite-difference rollout — dataset/apps/generate_burgers.py:207-299 (burgers_numeric_solve_free), stencils
 * of Diff_mat_1D (:95-110).  u0 [N][s], f [N][Nt][s] -> traj [N][Nt+1][s] (u0 followed by one record per force window).
 * t0,t1 = -/+ 1/(2dx), d0,d1,d2 = visc*(1,-2,1)/dx^2 as fp32 (host: generate_burgers.py:255-258), steps = ceil(T/dt). */
int dpc_burgers_rollout(const float* u0, const float* f, float* traj, int32_t N, int32_t s, int32_t Nt, int32_t steps,
                        float t0, float t1, float d0, float d1, float d2, float dt, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Jellyfish surrogate networks: the 2-D `Unet` boundary updater and `ForceUnet`
 * (diffusion/diffusion_2d_jellyfish.py:276-403, :406-481; cited as jf.py) forward and BACKWARD w.r.t. activations and the
 * time conditioning — what `force_fn` (inference/inference_2d_jellyfish.py:85-114) obtains through torch.autograd.grad.
 * Tensors are channels-last fp32 [N images][HW][C].  The dgrad convolutions are dpc_conv_igemm / dpc_conv3d_tcgen05 calls
 * with transposed, spatially flipped packed weights; the entry points below are everything else.
 * ------------------------------------------------------------------------------------------------------- */
/* LinearAttention forward (jf.py:206-225): like dpc_spatial_linear_attention with v scaled by v_scale (= 1/(h*w), jf.py:219)
 * and, if kstat != NULL, the softmax_n(k) statistics [BF*heads][32][2] = (max, sum) kept for the backward pass. */
int dpc_spatial_linear_attention_ex(const float* qkv, float* ctx_ws, float* kstat, float* out, int32_t BF, int32_t HW,
                                    int32_t heads, float v_scale, void* stream);
/* LinearAttention core backward: qkv [BF*HW][3*heads*32], ctx [BF*heads][32][32] and kstat from the forward, dout
 * [BF*HW][heads*32] -> dqkv (same layout as qkv); dctx_ws: BF*heads*32*32 floats; scale = 32^-0.5. */
int dpc_linattn2d_bwd(const float* qkv, const float* ctx, const float* kstat, const float* dout, float* dctx_ws, float* dqkv,
                      int32_t BF, int32_t HW, int32_t heads, float scale, float v_scale, void* stream);
/* softmax Attention core backward (jf.py:241-255): out = the forward's attention output [BF*HW][heads*32]; returns -2 when
 * HW exceeds what one CTA's shared memory holds (~780 tokens). */
int dpc_attention2d_bwd(const float* qkv, const float* out, const float* dout, float* dqkv, int32_t BF, int32_t HW,
                        int32_t heads, float scale, void* stream);
/* Backward of dpc_groupnorm_silu (Block.forward after proj, jf.py:196-204) w.r.t. the raw conv output y:
 * dy [B][rows][C]; dss (nullable, same layout / stride / offset as scale_shift) receives d scale and d shift per (sample,
 * channel); sums_ws: B*C*2 doubles. */
int dpc_gn_silu_bwd(const float* y, const double* stats, const float* gamma, const float* beta, const float* scale_shift,
                    int64_t ss_stride, int64_t ss_off, const float* dout, float* dy, double* sums_ws, float* dss, int32_t B,
                    int64_t rows_per_sample, int32_t C, int32_t groups, float eps, void* stream);
/* Backward of dpc_layernorm_channels w.r.t. x: dx = LN'(x; gamma) dy (+ add, nullable: the residual branch's gradient). */
int dpc_layernorm_channels_bwd(const float* x, const float* gamma, const float* dy, const float* add, float* dx, int64_t rows,
                               int32_t C, float eps, int32_t use_rsqrt, void* stream);
/* out = a + b (n % 4 == 0; out may alias a or b). */
int dpc_add(const float* a, const float* b, float* out, int64_t n, void* stream);
/* Backward of dpc_upsample_nearest2x: dy [N,2H,2W,C] -> dx [N,H,W,C] (sum of each 2x2 block). */
int dpc_sumpool2x2(const float* dy, float* dx, int64_t N, int32_t H, int32_t W, int32_t C, void* stream);
/* ForceUnet head (jf.py:478-479): out[n][o] = bias[o] + sum_c W[o][c] * mean_hw x[n][:][c], and its backward w.r.t. x. */
int dpc_mean_head(const float* x, const float* W, const float* bias, float* out, int64_t N, int32_t HW, int32_t C, int32_t O,
                  void* stream);
int dpc_mean_head_bwd(const float* dout, const float* W, float* dx, int64_t N, int32_t HW, int32_t C, int32_t O, void* stream);
/* time_mlp with a FLOAT time (the boundary updater is conditioned on theta, jf.py:313-318, inference_2d_jellyfish.py:99-102):
 * same as dpc_time_embed otherwise. */
int dpc_time_embed_f32(const float* t, const float* freqs, const float* w1, const float* b1, const float* w2, const float* b2,
                       float* hidden_ws, float* t_emb, int32_t B, int32_t dim, void* stream);
/* d(sum over all ResnetBlock scale/shift rows)/d t: backward through dpc_time_proj (w_proj [total][4*dim]), SiLU, time_mlp and
 * the sinusoidal embedding, given dss [B][total] and the forward's t_emb [B][4*dim] -> dt [B]. */
int dpc_time_mlp_bwd(const float* t, const float* freqs, const float* w1, const float* b1, const float* w2, const float* w_proj,
                     const float* t_emb, const float* dss, float* dt, int32_t B, int32_t dim, int32_t total, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DPC_B200_H */
