// Fused guidance + prior re-weighting + posterior (DDPM) / DDIM update of one denoising step, reference layout
// [B,F,6,H,W] fp32.  One read of x, eps_joint, eps_w, noise; one write of x_{t-1} (and optionally x_start).
// Follows smoke.py:610-656 (model_predictions), :659-666, :671-699 (p_sample), :720, :739-775 (ddim_sample) and the
// stock guidance of inference/inference_2d_smoke.py:30-44 in closed form.  Every product/sum is an explicit
// round-to-nearest intrinsic in the reference's evaluation order so the result is bit-comparable with PyTorch's
// unfused elementwise kernels (no FMA contraction).
#include "common.cuh"

namespace dpc {

struct StepGeom {
  int F, HW;
  float inv_hw_neg;     // -(1/HW): d(-mean_{h,w})/dx
  float energy_coef;    // fl(w_energy / (F*2*HW))
};

__device__ __forceinline__ float clamp1(float v) { return fminf(fmaxf(v, -1.0f), 1.0f); }

// guided x_start / pred_noise for one element.  c = channel, is_last_frame = (f == F-1)
template <bool CLIP>
__device__ __forceinline__ void guided_prediction(float x, float ej, float ew /*0 outside 3:5*/, float g_user,
                                                  bool stock, int c, bool is_last_frame, const dpc_step_coefs& k,
                                                  const StepGeom& gm, float& pred_noise, float& x_start) {
  const float srx = __fmul_rn(k.sqrt_recip_alphas_cumprod, x);
  float xs0 = __fsub_rn(srx, __fmul_rn(k.sqrt_recipm1_alphas_cumprod, ej));
  if (CLIP) xs0 = clamp1(xs0);
  float g;
  if (stock) {
    g = 0.0f;
    if (c == 5 && is_last_frame) g = gm.inv_hw_neg;
    if (c == 3 || c == 4) {
      const float xr = __fmul_rn(xs0, k.rescaler[c]);
      g = __fmul_rn(gm.energy_coef, __fmul_rn(2.0f, xr));
    }
  } else {
    g = g_user;
  }
  const float grad_final = __fadd_rn(__fmul_rn(k.guidance_coef, g), __fmul_rn(k.prior_coef, ew));
  pred_noise = __fadd_rn(ej, grad_final);
  x_start = __fsub_rn(srx, __fmul_rn(k.sqrt_recipm1_alphas_cumprod, pred_noise));
  if (CLIP) {
    x_start = clamp1(x_start);
    pred_noise = __fdiv_rn(__fsub_rn(srx, x_start), k.sqrt_recipm1_alphas_cumprod);  // rederive_pred_noise
  }
}

template <bool DDIM, int VEC>
__global__ void __launch_bounds__(256)
guided_step_kernel(const float* x, const float* __restrict__ eps_joint, const float* __restrict__ eps_w,
                   const float* __restrict__ noise, const float* __restrict__ init, const float* __restrict__ g_user,
                   int stock, const dpc_step_coefs k_host, const dpc_step_coefs* __restrict__ k_dev, const StepGeom gm, float* x_out,
                   float* __restrict__ x_start_out, int64_t nvec) {
  // k_dev != NULL: the step's coefficients live in device memory (written by sampler_prepare_kernel inside a captured CUDA
  // graph, so one graph serves every step of the schedule); the noise operand is then gated by the coefficients themselves
  const dpc_step_coefs k = k_dev ? *k_dev : k_host;
  if (k_dev && (DDIM ? k.last != 0 : k.add_noise == 0)) noise = nullptr;
  const int hwv = gm.HW / VEC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int pv = (int)(i % hwv);
    const int64_t r = i / hwv;       // (b*F + f)*6 + c
    const int c = (int)(r % 6);
    const int64_t bf = r / 6;
    const int f = (int)(bf % gm.F);
    const int64_t b = bf / gm.F;
    const size_t off = (size_t)i * VEC;
    float xv[VEC], ej[VEC], ew[VEC], nz[VEC], gu[VEC], o[VEC], xs[VEC];
    const bool wch = (c == 3 || c == 4);
    const size_t woff = ((size_t)bf * 2 + (c - 3)) * gm.HW + (size_t)pv * VEC;
    if (VEC == 4) {
      float4 a = *reinterpret_cast<const float4*>(x + off);
      float4 e = __ldcs(reinterpret_cast<const float4*>(eps_joint + off));
      xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
      ej[0] = e.x; ej[1] = e.y; ej[2] = e.z; ej[3] = e.w;
      if (wch) {
        float4 w4 = __ldcs(reinterpret_cast<const float4*>(eps_w + woff));
        ew[0] = w4.x; ew[1] = w4.y; ew[2] = w4.z; ew[3] = w4.w;
      }
      if (noise) {
        float4 n4 = __ldcs(reinterpret_cast<const float4*>(noise + off));
        nz[0] = n4.x; nz[1] = n4.y; nz[2] = n4.z; nz[3] = n4.w;
      }
      if (!stock) {
        float4 g4 = __ldcs(reinterpret_cast<const float4*>(g_user + off));
        gu[0] = g4.x; gu[1] = g4.y; gu[2] = g4.z; gu[3] = g4.w;
      }
    } else {
      xv[0] = x[off];
      ej[0] = eps_joint[off];
      if (wch) ew[0] = eps_w[woff];
      if (noise) nz[0] = noise[off];
      if (!stock) gu[0] = g_user[off];
    }
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      if (!wch) ew[v] = 0.0f;
      if (!noise) nz[v] = 0.0f;
      if (stock) gu[v] = 0.0f;
      float pn, x0;
      if (DDIM) {
        guided_prediction<true>(xv[v], ej[v], ew[v], gu[v], stock != 0, c, f == gm.F - 1, k, gm, pn, x0);
        if (k.last) {
          o[v] = x0;
        } else {
          const float t1 = __fadd_rn(__fmul_rn(x0, k.sqrt_alpha_next), __fmul_rn(k.c, pn));
          o[v] = __fadd_rn(t1, __fmul_rn(k.ddim_sigma, nz[v]));
        }
      } else {
        guided_prediction<false>(xv[v], ej[v], ew[v], gu[v], stock != 0, c, f == gm.F - 1, k, gm, pn, x0);
        x0 = clamp1(x0);
        const float mean = __fadd_rn(__fmul_rn(k.posterior_mean_coef1, x0), __fmul_rn(k.posterior_mean_coef2, xv[v]));
        o[v] = k.add_noise ? __fadd_rn(mean, __fmul_rn(k.sigma, nz[v])) : mean;
      }
      xs[v] = x0;
    }
    // re-impose the initial condition x[:, 0, 0] = init (smoke.py:720, :775); not applied on the final DDIM return
    if (init != nullptr && f == 0 && c == 0 && !(DDIM && k.last)) {
#pragma unroll
      for (int v = 0; v < VEC; ++v) o[v] = init[(size_t)b * gm.HW + (size_t)pv * VEC + v];
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(x_out + off) = make_float4(o[0], o[1], o[2], o[3]);
      if (x_start_out) *reinterpret_cast<float4*>(x_start_out + off) = make_float4(xs[0], xs[1], xs[2], xs[3]);
    } else {
      x_out[off] = o[0];
      if (x_start_out) x_start_out[off] = xs[0];
    }
  }
}

__global__ void __launch_bounds__(256)
predict_x_start_kernel(const float* __restrict__ x, const float* __restrict__ eps, float sr, float srm1, int clip,
                       float* __restrict__ out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __fsub_rn(__fmul_rn(sr, x[i]), __fmul_rn(srm1, eps[i]));
    if (clip) v = clamp1(v);
    out[i] = v;
  }
}

// sidx is a device-side step counter: tt[0..B) <- t_table[*sidx], *cur <- c_table[*sidx], then *sidx += 1 (one thread block)
__global__ void sampler_prepare_kernel(const int64_t* __restrict__ t_table, const dpc_step_coefs* __restrict__ c_table,
                                       int32_t* sidx, int32_t nsteps, int64_t* __restrict__ tt, int B, dpc_step_coefs* cur) {
  int i = *sidx;
  if (i >= nsteps) i = nsteps - 1;             // replaying past the schedule repeats its last step instead of reading out of bounds
  const int64_t t = t_table[i];
  for (int b = threadIdx.x; b < B; b += blockDim.x) tt[b] = t;
  __syncthreads();                              // every thread has read *sidx before it moves
  if (threadIdx.x == 0) {
    *cur = c_table[i];
    *sidx = i + 1;
  }
}

template <bool DDIM>
static int launch_step(const float* x, const float* eps_joint, const float* eps_w, const float* noise, const float* init,
                       const float* g, int stock, const dpc_step_coefs* coefs, float* x_out, float* x_start_out, int B,
                       int F, int H, int W, void* stream, const dpc_step_coefs* coefs_dev = nullptr) {
  DPC_CHECK_ARG(x && eps_joint && eps_w && coefs && x_out && B > 0 && F > 0 && H > 0 && W > 0);
  DPC_CHECK_ARG(stock || g != nullptr);
  if (coefs_dev) DPC_CHECK_ARG(noise != nullptr);
  else if (!DDIM) DPC_CHECK_ARG(!coefs->add_noise || noise != nullptr);
  else DPC_CHECK_ARG(coefs->last || noise != nullptr);
  StepGeom gm;
  gm.F = F;
  gm.HW = H * W;
  gm.inv_hw_neg = -(1.0f / (float)(H * W));
  gm.energy_coef = coefs->w_energy / (float)((int64_t)F * 2 * H * W);
  const int64_t n = (int64_t)B * F * 6 * H * W;
  const bool vec = (gm.HW % 4 == 0);
  const int64_t nvec = vec ? n / 4 : n;
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = 148LL * 8 * 4;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const float* nz = noise;
  if (!coefs_dev) {
    if (!DDIM && !coefs->add_noise) nz = nullptr;
    if (DDIM && coefs->last) nz = nullptr;
  }
  if (vec)
    guided_step_kernel<DDIM, 4><<<(unsigned)blocks, 256, 0, st>>>(x, eps_joint, eps_w, nz, init, g, stock, *coefs, coefs_dev, gm,
                                                                  x_out, x_start_out, nvec);
  else
    guided_step_kernel<DDIM, 1><<<(unsigned)blocks, 256, 0, st>>>(x, eps_joint, eps_w, nz, init, g, stock, *coefs, coefs_dev, gm,
                                                                  x_out, x_start_out, nvec);
  DPC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpc

extern "C" int dpc_ddpm_guided_step(const float* x, const float* eps_joint, const float* eps_w, const float* noise,
                                    const float* init, const float* g, int32_t use_stock_guidance,
                                    const dpc_step_coefs* coefs, float* x_out, float* x_start_out, int32_t B, int32_t F,
                                    int32_t H, int32_t W, void* stream) {
  return dpc::launch_step<false>(x, eps_joint, eps_w, noise, init, g, use_stock_guidance, coefs, x_out, x_start_out, B, F, H,
                                 W, stream);
}

extern "C" int dpc_ddim_guided_step(const float* x, const float* eps_joint, const float* eps_w, const float* noise,
                                    const float* init, const float* g, int32_t use_stock_guidance,
                                    const dpc_step_coefs* coefs, float* x_out, float* x_start_out, int32_t B, int32_t F,
                                    int32_t H, int32_t W, void* stream) {
  return dpc::launch_step<true>(x, eps_joint, eps_w, noise, init, g, use_stock_guidance, coefs, x_out, x_start_out, B, F, H,
                                W, stream);
}

extern "C" int dpc_sampler_prepare(const int64_t* t_table, const dpc_step_coefs* c_table, int32_t* step_index, int32_t nsteps,
                                   int64_t* tt, int32_t B, dpc_step_coefs* cur, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(t_table && c_table && step_index && tt && cur && nsteps > 0 && B > 0);
  sampler_prepare_kernel<<<1, 128, 0, (cudaStream_t)stream>>>(t_table, c_table, step_index, nsteps, tt, B, cur);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_guided_step_dev(int32_t ddim, const float* x, const float* eps_joint, const float* eps_w, const float* noise,
                                   const float* init, const dpc_step_coefs* coefs_host, const dpc_step_coefs* coefs_dev,
                                   float* x_out, float* x_start_out, int32_t B, int32_t F, int32_t H, int32_t W, void* stream) {
  DPC_CHECK_ARG(coefs_dev != nullptr);
  if (ddim)
    return dpc::launch_step<true>(x, eps_joint, eps_w, noise, init, nullptr, 1, coefs_host, x_out, x_start_out, B, F, H, W, stream,
                                  coefs_dev);
  return dpc::launch_step<false>(x, eps_joint, eps_w, noise, init, nullptr, 1, coefs_host, x_out, x_start_out, B, F, H, W, stream,
                                 coefs_dev);
}

namespace dpc {
// x_t = a * x + (b * z): the reference's evaluation order (two rounded products, one rounded sum); z may be null (t == 0: + 0)
__global__ void __launch_bounds__(256) renoise_kernel(const float* x, const float* __restrict__ z, float a, float b, float* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(__fmul_rn(a, x[i]), z ? __fmul_rn(b, z[i]) : 0.0f);
}
}  // namespace dpc

extern "C" int dpc_renoise(const float* x, const float* z, float a, float b, float* out, int64_t n, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(x && out && n > 0);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = 148LL * 8 * 4;
  if (blocks > cap) blocks = cap;
  renoise_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, z, a, b, out, n);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_predict_x_start(const float* x, const float* eps, float sqrt_recip, float sqrt_recipm1, int32_t clip,
                                   float* x_start, int64_t n, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(x && eps && x_start && n > 0);
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = 148LL * 8 * 4;
  if (blocks > cap) blocks = cap;
  predict_x_start_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, eps, sqrt_recip, sqrt_recipm1, clip, x_start, n);
  DPC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// Burgers sampler (diffusion/diffusion_1d_burgers.py:396-470): two-model prior re-weighting + x0 prediction, and the
// guided posterior step.  Same conventions as above: explicit round-to-nearest ops in the reference's evaluation order.
// ---------------------------------------------------------------------------------------------------------------------
namespace dpc {

// mode 0: out = eps1 - coef*eps2' ; mode 1: out = (eps1 - coef*eps2') / beta ; mode 2: out = (beta*eps1)'
// (' = channel 0 zeroed, diffusion_1d_burgers.py:403, :414);  x_start = sr*x - srm1*out
__global__ void __launch_bounds__(256)
burgers_model_output_kernel(const float* __restrict__ x, const float* __restrict__ eps1, const float* __restrict__ eps2,
                            float* __restrict__ out, float* __restrict__ x_start, int mode, float coef, float beta, float sr,
                            float srm1, int C, int64_t plane, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / plane) % C);
    float o;
    if (mode == 2) {
      o = (c == 0) ? 0.0f : __fmul_rn(beta, eps1[i]);
    } else {
      const float e2 = (c == 0) ? 0.0f : eps2[i];
      o = __fsub_rn(eps1[i], __fmul_rn(coef, e2));
      if (mode == 1) o = __fdiv_rn(o, beta);
    }
    out[i] = o;
    if (x_start) x_start[i] = __fsub_rn(__fmul_rn(sr, x[i]), __fmul_rn(srm1, o));
  }
}

// pred_noise = eps (+ g*gscale); x_start = [clamp](sr*x - srm1*pred_noise); x_{t-1} = c1*x_start + c2*x (+ sigma*noise)
__global__ void __launch_bounds__(256)
ddpm_posterior_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ g,
                           const float* __restrict__ noise, float* __restrict__ x_out, float* __restrict__ x_start_out,
                           float* __restrict__ pred_noise_out, float gscale, float sr, float srm1, int clip, float c1,
                           float c2, float sigma, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pn = eps[i];
    if (g) pn = __fadd_rn(pn, __fmul_rn(g[i], gscale));
    float xs = __fsub_rn(__fmul_rn(sr, x[i]), __fmul_rn(srm1, pn));
    if (clip) xs = fminf(fmaxf(xs, -1.0f), 1.0f);
    const float mean = __fadd_rn(__fmul_rn(c1, xs), __fmul_rn(c2, x[i]));
    x_out[i] = noise ? __fadd_rn(mean, __fmul_rn(sigma, noise[i])) : mean;
    if (x_start_out) x_start_out[i] = xs;
    if (pred_noise_out) pred_noise_out[i] = pn;
  }
}

}  // namespace dpc

extern "C" int dpc_burgers_model_output(const float* x, const float* eps1, const float* eps2, float* out, float* x_start,
                                        int32_t mode, float coef, float beta, float sqrt_recip, float sqrt_recipm1, int32_t C,
                                        int64_t plane, int64_t n, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(x && eps1 && out && n > 0 && C > 0 && plane > 0 && mode >= 0 && mode <= 2 && (mode == 2 || eps2));
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  burgers_model_output_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, eps1, eps2, out, x_start, mode, coef, beta,
                                                                                  sqrt_recip, sqrt_recipm1, C, plane, n);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_ddpm_posterior_step(const float* x, const float* eps, const float* g, const float* noise, float* x_out,
                                       float* x_start_out, float* pred_noise_out, float gscale, float sqrt_recip,
                                       float sqrt_recipm1, int32_t clip, float coef1, float coef2, float sigma, int64_t n,
                                       void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(x && eps && x_out && n > 0);
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  ddpm_posterior_step_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, eps, g, noise, x_out, x_start_out,
                                                                                 pred_noise_out, gscale, sqrt_recip, sqrt_recipm1,
                                                                                 clip, coef1, coef2, sigma, n);
  DPC_LAUNCH_CHECK();
  return 0;
}
