// Jellyfish sampler step: the elementwise / reduction work of
// diffusion/diffusion_2d_jellyfish.py p_sample (:776-806), p_mean_variance (:759-771), the in-model guidance of the DDIM path
// (:716-742), update_bd's theta mean (:810) and the re-imposed conditions of p_sample_loop (:858-864) in three kernels.
//
// State layout is the reference's: x [B, F, 7, H, W] = [state(3), boundary(3), theta(1)], fp32, plane P = H*W.
// The 4 diffused channels are {0,1,2,6}; eps / x_start / pred / g are [B, F, 4, P]; eps_w is [B, F, 1, P].
// Every product and sum is a separately rounded fp32 operation in the reference's evaluation order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace dpc {

__device__ __forceinline__ int jelly_src_channel(int c) { return c < 3 ? c : 6; }

// x_start = clamp?(sr * x4 - srm1 * eps)                                   (jellyfish.py:714, :744, :764)
__global__ void __launch_bounds__(256)
jelly_x_start_kernel(const float* __restrict__ x, const float* __restrict__ eps, float* __restrict__ x_start, float sr,
                     float srm1, int clip, int64_t P, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % P, c = (i / P) & 3, bf = i / (4 * P);
    const float xv = x[(bf * 7 + jelly_src_channel((int)c)) * P + p];
    float xs = __fsub_rn(__fmul_rn(sr, xv), __fmul_rn(srm1, eps[i]));
    if (clip) xs = fminf(fmaxf(xs, -1.0f), 1.0f);
    x_start[i] = xs;
  }
}

struct JellyStepArgs {
  const float* x;        // [B,F,7,P] current state
  const float* x_start;  // [B,F,4,P]
  const float* eps;      // [B,F,4,P] joint-model noise (DDIM only)
  const float* eps_w;    // [B,F,1,P]
  const float* g;        // [B,F,4,P] or null
  const float* noise;    // [B,F,4,P] or null
  const float* state_0;  // [B,3,P]
  const float* thetas_0; // [B]
  float* x_next;         // [B,F,7,P]: channels 0..2 and 6 written
  float* x_w;            // [B,F,7,P]: channel 6 written (input of the prior model)
  float* dtheta;         // [B,F] mean(theta) before conditioning - thetas_0   (update_bd, jellyfish.py:810-814)
  float* theta_mean;     // [B,F] mean(theta) after conditioning               (jellyfish.py:876)
  float ga, gb;          // guidance = ga * g - gb * eps_w
  float c1, c2, sigma;   // DDPM: mean = c1*x_start + c2*x ; DDIM: x = c1*x_start + c2*pred_noise
  int ddim, F, cond_steps;
  int64_t P;
};

// One CTA per (b, f): update the 4 diffused channels, reduce theta, impose the conditions.
__global__ void __launch_bounds__(256) jelly_step_kernel(JellyStepArgs a) {
  __shared__ float red[8];
  const int64_t bf = blockIdx.x, P = a.P;
  const int b = (int)(bf / a.F), f = (int)(bf % a.F);
  const bool head = f < a.cond_steps, tail = f >= a.F - a.cond_steps;
  const float th0 = a.thetas_0[b];
  float tsum = 0.0f;
  for (int64_t p = threadIdx.x; p < P; p += blockDim.x) {
    const float ew = a.eps_w[bf * P + p];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int64_t i4 = (bf * 4 + c) * P + p, i7 = (bf * 7 + jelly_src_channel(c)) * P + p;
      const float xs = a.x_start[i4];
      float v;
      if (a.ddim) {           // jellyfish.py:738-742 (guidance inside model_predictions, eps_w padded to channel 3), :925-927
        float pn = a.eps[i4];
        if (a.g) pn = __fadd_rn(pn, __fsub_rn(__fmul_rn(a.ga, a.g[i4]), __fmul_rn(a.gb, c == 3 ? ew : 0.0f)));
        v = __fadd_rn(__fmul_rn(xs, a.c1), __fmul_rn(a.c2, pn));
        if (a.noise) v = __fadd_rn(v, __fmul_rn(a.sigma, a.noise[i4]));
      } else {                // jellyfish.py:601-604, :789-804 (eps_w broadcast over the 4 channels)
        v = __fadd_rn(__fmul_rn(a.c1, xs), __fmul_rn(a.c2, a.x[i7]));
        if (a.noise) v = __fadd_rn(v, __fmul_rn(a.sigma, a.noise[i4]));
        if (a.g) v = __fsub_rn(v, __fsub_rn(__fmul_rn(a.ga, a.g[i4]), __fmul_rn(a.gb, ew)));
      }
      if (c < 3) {
        a.x_next[i7] = head ? a.state_0[((int64_t)b * 3 + c) * P + p] : v;
      } else {
        tsum += v;
        const float t = (head || tail) ? th0 : v;
        a.x_next[i7] = t;
        a.x_w[i7] = t;
      }
    }
  }
  tsum = warp_sum(tsum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = tsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.0f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    const float m = s / (float)P;
    a.dtheta[bf] = __fsub_rn(m, th0);
    a.theta_mean[bf] = (head || tail) ? th0 : m;
  }
}

// x_next[:, :, 3:6] = x_w[:, :, 3:6] = bd_updater output, frames < cond and >= F - cond <- bd_0   (jellyfish.py:860-861)
__global__ void __launch_bounds__(256)
jelly_write_bd_kernel(const float* __restrict__ pred_bd, const float* __restrict__ bd_0, float* __restrict__ x_next,
                      float* __restrict__ x_w, int F, int cond_steps, int64_t P, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i % P, c = (i / P) % 3, bf = i / (3 * P);
    const int64_t b = bf / F, f = bf % F;
    const bool fixed = f < cond_steps || f >= F - cond_steps;
    const float v = fixed ? bd_0[(b * 3 + c) * P + p] : pred_bd[i];
    const int64_t o = (bf * 7 + 3 + c) * P + p;
    x_next[o] = v;
    x_w[o] = v;
  }
}

}  // namespace dpc

extern "C" int dpc_jelly_x_start(const float* x, const float* eps, float* x_start, float sqrt_recip, float sqrt_recipm1,
                                 int32_t clip, int64_t BF, int64_t P, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(x && eps && x_start && BF > 0 && P > 0);
  const int64_t n = BF * 4 * P;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  jelly_x_start_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, eps, x_start, sqrt_recip, sqrt_recipm1, clip, P, n);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_jelly_step(const float* x, const float* x_start, const float* eps, const float* eps_w, const float* g,
                              const float* noise, const float* state_0, const float* thetas_0, float* x_next, float* x_w,
                              float* dtheta, float* theta_mean, float ga, float gb, float c1, float c2, float sigma,
                              int32_t ddim, int32_t B, int32_t F, int32_t cond_steps, int64_t P, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(x && x_start && eps_w && state_0 && thetas_0 && x_next && x_w && dtheta && theta_mean);
  DPC_CHECK_ARG(B > 0 && F > 0 && P > 0 && cond_steps >= 0 && 2 * cond_steps <= F && (!ddim || eps));
  JellyStepArgs a{x, x_start, eps, eps_w, g, noise, state_0, thetas_0, x_next, x_w, dtheta, theta_mean,
                  ga, gb, c1, c2, sigma, ddim, F, cond_steps, P};
  jelly_step_kernel<<<(unsigned)((int64_t)B * F), 256, 0, (cudaStream_t)stream>>>(a);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_jelly_write_bd(const float* pred_bd, const float* bd_0, float* x_next, float* x_w, int32_t B, int32_t F,
                                  int32_t cond_steps, int64_t P, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(pred_bd && bd_0 && x_next && x_w && B > 0 && F > 0 && P > 0 && cond_steps >= 0);
  const int64_t n = (int64_t)B * F * 3 * P;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  jelly_write_bd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred_bd, bd_0, x_next, x_w, F, cond_steps, P, n);
  DPC_LAUNCH_CHECK();
  return 0;
}
