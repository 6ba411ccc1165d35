// Time conditioning: SinusoidalPosEmb -> Linear -> GELU -> Linear (conv3d.py:139-151, :404-409) and the per-ResnetBlock
// SiLU -> Linear projections (conv3d.py:211-214, :222-224), batched over all blocks as one row-concatenated matrix.
// Tiny, latency-bound: one warp per output element, fp32 FMA dot products with a shuffle reduction.
#include "common.cuh"

namespace dpc {

__global__ void sinusoidal_linear_gelu_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                              const float* __restrict__ w1, const float* __restrict__ b1,
                                              float* __restrict__ hidden, int B, int dim, int tdim) {
  // hidden[b][j] = gelu(b1[j] + sum_k w1[j][k] * emb[b][k]),  emb = [sin(t f_i) | cos(t f_i)]
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= (int64_t)B * tdim) return;
  const int b = (int)(wg / tdim), j = (int)(wg % tdim);
  const int half = dim / 2;
  const float tv = (float)t[b];
  float acc = 0.f;
  for (int k = lane; k < dim; k += 32) {
    const float arg = __fmul_rn(tv, freqs[k < half ? k : k - half]);
    const float e = (k < half) ? sinf(arg) : cosf(arg);
    acc = fmaf(w1[(size_t)j * dim + k], e, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    const float v = acc + b1[j];
    hidden[(size_t)b * tdim + j] = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  }
}

template <bool SILU_IN>
__global__ void linear_rows_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                   const float* __restrict__ bias, float* __restrict__ out, int B, int K, int N) {
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= (int64_t)B * N) return;
  const int b = (int)(wg / N), j = (int)(wg % N);
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    float v = x[(size_t)b * K + k];
    if (SILU_IN) v = silu_f(v);
    acc = fmaf(W[(size_t)j * K + k], v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)b * N + j] = acc + (bias ? bias[j] : 0.f);
}

}  // namespace dpc

extern "C" int dpc_time_embed(const int64_t* t, const float* freqs, const float* w1, const float* b1, const float* w2,
                              const float* b2, float* hidden_ws, float* t_emb, int32_t B, int32_t dim, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(t && freqs && w1 && b1 && w2 && b2 && hidden_ws && t_emb && B > 0 && dim > 0 && dim % 2 == 0);
  const int tdim = dim * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t warps = (int64_t)B * tdim;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  sinusoidal_linear_gelu_kernel<<<blocks, 256, 0, st>>>(t, freqs, w1, b1, hidden_ws, B, dim, tdim);
  DPC_LAUNCH_CHECK();
  linear_rows_kernel<false><<<blocks, 256, 0, st>>>(hidden_ws, w2, b2, t_emb, B, tdim, tdim);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_time_proj(const float* t_emb, const float* W, const float* bias, float* out, int32_t B, int32_t tdim,
                             int32_t total, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(t_emb && W && out && B > 0 && tdim > 0 && total > 0);
  const int64_t warps = (int64_t)B * total;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  linear_rows_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(t_emb, W, bias, out, B, tdim, total);
  DPC_LAUNCH_CHECK();
  return 0;
}
