// Evaluation-stage reductions of the smoke task: the sums behind InferencePipeline.multi_evaluate
// (inference/inference_2d_smoke.py:388-427) taken straight from the rollout outputs, without materialising solver_out
// [B,256,6,128,128] or its strided views.  The reference compares pred [B,F,6,S,S] with data_current = solver_out[:, ::T/F, :,
// ::128/S, ::128/S] (:384-386) after zeroing frame 0 of both (:398-401); channels of data_current are density, velocity x/y,
// the (masked, tiled) sampled controls, and the smoke portion broadcast over the frame (:366-371).
// sums[b][k], double:  k = 0..5   sum (pred_c - data_c)^2 over frames 1..F-1 and the frame, c = 0..5
//                      k = 6..10  sum data_c^2, c = 0..4
//                      k = 11     sum over the last frame of pred[:, F-1, 5]
#include "common.cuh"

namespace dpc {

constexpr int NSUM = 12;

__global__ void __launch_bounds__(256)
smoke_eval_sums_kernel(const float* __restrict__ pred, const float* __restrict__ dens, const double* __restrict__ vel,
                       const double* __restrict__ smoke_out, double* __restrict__ sums, int F, int S, int T, int lo, int hi) {
  const int b = blockIdx.y;
  const int tstep = T / F, sstep = 128 / S;
  const int64_t per = (int64_t)(F - 1) * S * S;          // frames 1..F-1
  double acc[NSUM];
#pragma unroll
  for (int k = 0; k < NSUM; ++k) acc[k] = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % S);
    const int y = (int)((i / S) % S);
    const int f = (int)(i / ((int64_t)S * S)) + 1;
    const int t = f * tstep;
    const size_t pbase = (((size_t)b * F + f) * 6) * S * S + (size_t)y * S + x;
    const size_t gbase = (((size_t)b * T + t) * 128 + (size_t)y * sstep) * 128 + (size_t)x * sstep;
    float p[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) p[c] = pred[pbase + (size_t)c * S * S];
    const bool inner = (y >= lo && y < hi && x >= lo && x < hi);   // indirect control: the sampled force is dropped there (:322)
    double d[6];
    d[0] = (double)dens[gbase];
    d[1] = vel[gbase * 2];
    d[2] = vel[gbase * 2 + 1];
    d[3] = inner ? 0.0 : (double)p[3];
    d[4] = inner ? 0.0 : (double)p[4];
    d[5] = smoke_out[(size_t)b * T + t];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const double e = (double)p[c] - d[c];
      acc[c] += e * e;
    }
#pragma unroll
    for (int c = 0; c < 5; ++c) acc[6 + c] += d[c] * d[c];
    if (f == F - 1) acc[11] += (double)p[5];
  }
  __shared__ double s_red[8][NSUM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NSUM; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_double(v, o);
    if (lane == 0) s_red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < NSUM) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += s_red[w][threadIdx.x];
    atomicAdd(sums + (size_t)b * NSUM + threadIdx.x, v);
  }
}

}  // namespace dpc

extern "C" int dpc_smoke_eval_sums(const float* pred, const float* densitys, const double* velocitys, const double* smoke_out,
                                   double* sums, int32_t B, int32_t F, int32_t S, int32_t T, int32_t mask_lo, int32_t mask_hi,
                                   void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(pred && densitys && velocitys && smoke_out && sums && B > 0 && B <= 65535 && F > 1 && S > 0 && 128 % S == 0 &&
                T % F == 0);
  cudaStream_t st = (cudaStream_t)stream;
  DPC_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * NSUM * sizeof(double), st));
  const int64_t per = (int64_t)(F - 1) * S * S;
  int blocks = (int)((per + 255) / 256);
  if (blocks > 64) blocks = 64;
  smoke_eval_sums_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, st>>>(pred, densitys, velocitys, smoke_out, sums, F, S, T,
                                                                             mask_lo, mask_hi);
  DPC_LAUNCH_CHECK();
  return 0;
}
