// Burgers finite-difference rollout: dataset/apps/generate_burgers.py:207-299
// (burgers_numeric_solve_free) with the stencils of Diff_mat_1D (:95-110).  Explicit Euler, 10 000 steps per trajectory,
// homogeneous Dirichlet ends, force piecewise constant per record window.  One CTA per trajectory, one thread per grid
// point, the state ping-pongs between two shared-memory rows: the whole time loop runs in one launch (the reference
// issues ~8 tensor ops per step).  float32 with explicit round-to-nearest ops in the reference's evaluation order, so
// the result is bit-identical to the PyTorch CPU run.
#include "common.cuh"

namespace dpc {

__global__ void __launch_bounds__(1024)
burgers_rollout_kernel(const float* __restrict__ u0, const float* __restrict__ f, float* __restrict__ traj, int s, int Nt,
                       int steps, int rec, float t0, float t1, float d0, float d1, float d2, float dt) {
  extern __shared__ float sm[];                 // [2][s + 2]
  const int n = blockIdx.x, i = threadIdx.x;    // i = interior point index
  float* ua = sm;
  float* ub = sm + (s + 2);
  if (i < s) {
    const float v = u0[(size_t)n * s + i];
    ua[i + 1] = v;
    traj[((size_t)n * (Nt + 1)) * s + i] = v;
  }
  if (i == 0) { ua[0] = 0.f; ua[s + 1] = 0.f; ub[0] = 0.f; ub[s + 1] = 0.f; }
  __syncthreads();
  int c = 0, fi = -1;
  float fv = 0.f;
  for (int j = 0; j < steps; ++j) {
    if (j % rec == 0) {
      ++fi;
      if (i < s) fv = f[((size_t)n * Nt + fi) * s + i];
    }
    if (i < s) {
      const float um = ua[i], uc = ua[i + 1], up = ua[i + 2];
      const float transport = __fadd_rn(__fmul_rn(__fmul_rn(um, um), t0), __fmul_rn(__fmul_rn(up, up), t1));
      const float diffusion = __fadd_rn(__fadd_rn(__fmul_rn(um, d0), __fmul_rn(uc, d1)), __fmul_rn(up, d2));
      const float rhs = __fadd_rn(__fadd_rn(__fmul_rn(-0.5f, transport), diffusion), fv);
      const float un = __fadd_rn(uc, __fmul_rn(dt, rhs));
      ub[i + 1] = un;
      if ((j + 1) % rec == 0 && c < Nt) traj[((size_t)n * (Nt + 1) + c + 1) * s + i] = un;
    }
    if ((j + 1) % rec == 0) ++c;
    __syncthreads();
    float* tmp = ua; ua = ub; ub = tmp;
  }
}

}  // namespace dpc

extern "C" int dpc_burgers_rollout(const float* u0, const float* f, float* traj, int32_t N, int32_t s, int32_t Nt,
                                   int32_t steps, float t0, float t1, float d0, float d1, float d2, float dt,
                                   void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(u0 && f && traj && N > 0 && s > 0 && s <= 1024 && Nt > 0 && steps >= Nt);
  const int rec = steps / Nt;
  const int threads = ((s + 31) / 32) * 32;
  burgers_rollout_kernel<<<(unsigned)N, threads, (size_t)2 * (s + 2) * sizeof(float), (cudaStream_t)stream>>>(
      u0, f, traj, s, Nt, steps, rec, t0, t1, d0, d1, d2, dt);
  DPC_LAUNCH_CHECK();
  return 0;
}
