// Post-sampling smoke rollout: one persistent CTA per trajectory runs all simulation steps of
// dataset/apps/evaluate_solver.py::solver (es.py:205-310) on chip — control injection (es.py:128-142), masking,
// divergence (phi/math/nd.py:367-377), the reference's plain conjugate-gradient pressure solve (phi/solver/base.py:56-103,
// <= 500 iterations, max|r| >= 1e-8 stop test, including its first-iteration aliasing of `momentum` and `residual`),
// pressure-gradient subtraction (nd.py:602-614), two semi-Lagrangian advections with the reference's clamp/zero-fill
// rule (nd.py:422-427, scipy_backend.py:58-77, :181-185) and the smoke-in-bucket accounting (es.py:279-305).
//
// Layout: 512 threads; cell i = tid + 512*j (j < 32) of the 127x127 pressure grid is owned by one thread, which keeps
// the residual r of its cells in registers (fp64, like the reference's NumPy run).  The search direction p of the whole
// grid lives in shared memory (129 KB) so the 5-point stencil never touches HBM, the solution x is accumulated in an
// L2-resident workspace (one coalesced read-modify-write per iteration); three fp64 vectors do not fit the register
// file + shared memory of one SM.  The Laplacian is applied from 4 neighbour bits + a small integer diagonal per cell
// (no matrix is materialised — the reference rebuilds a scipy sparse matrix every step).  Dot products are warp-shuffle + one shared-memory exchange with a fixed reduction tree (deterministic).
// The kernel is latency/fp64-issue bound (127 500 strictly sequential CG iterations per trajectory), not HBM bound.
#include "common.cuh"

namespace dpc {
namespace rollout {

constexpr int N = 127;
constexpr int NC = N * N;          // 16129 pressure cells
constexpr int NS = 128;
constexpr int NV = NS * NS;        // staggered samples per component
constexpr int THREADS = 512;
constexpr int PER = 32;            // cells per thread (512 x 32 >= 16129)

struct Args {
  const int8_t* fluid;             // [127][127]
  const float* vmask;              // [128][128][2]
  const float* init_velocity;      // [B][128][128][2]
  const float* init_density;       // [B][nx][nx]
  const float* c1;                 // [B][nt][nx][nx]
  const float* c2;
  double* vel_ws;                  // [B][2][128][128][2] ping-pong
  double* x_ws;                    // [B][127*127] CG solution accumulator
  float* dens_ws;                  // [B][2][2][127][127]: (buffer, field) ping-pong for density / zeroed density
  float* densitys;                 // [B][T][128][128]
  float* zero_densitys;            // [B][T][128][128]
  double* velocitys;               // [B][T][128][128][2]
  double* smoke_out;               // [B][T]
  int32_t* iterations;             // [B][T] CG iterations per step (0 for frame 0)
  int nt, nx, T;
  double dt, accuracy;
  int max_iterations;
};

// block-wide sum/max of NV values; all threads get the result.  `slot` alternates so that back-to-back calls never
// overwrite partials that slower warps are still reading (one __syncthreads per call).
template <int NVAL>
__device__ __forceinline__ void block_reduce(double (&v)[NVAL], bool last_is_max, double* s_part, int slot) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NVAL; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = shfl_xor_double(v[k], o);
      v[k] = (last_is_max && k == NVAL - 1) ? fmax(v[k], other) : v[k] + other;
    }
  }
  double* part = s_part + slot * (32 * 9);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NVAL; ++k) part[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NVAL; ++k) {
    double t = (lane < THREADS / 32) ? part[k * 32 + lane] : 0.0;   // identities: 0 for sums and for max|.|
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = shfl_xor_double(t, o);
      t = (last_is_max && k == NVAL - 1) ? fmax(t, other) : t + other;
    }
    v[k] = t;
  }
}

__device__ __forceinline__ int bucket_of(int y, int x) {
  // es.py:150-171: three bottom buckets, four side buckets; -1 = not in a bucket
  if (y >= 112 && y < 127) {
    if (x >= 22 && x < 42) return 0;
    if (x >= 54 && x < 74) return 1;
    if (x >= 86 && x < 106) return 2;
  }
  if (x < 16) {
    if (y >= 22 && y < 42) return 3;
    if (y >= 54 && y < 74) return 4;
  }
  if (x >= 112 && x < 127) {
    if (y >= 22 && y < 42) return 5;
    if (y >= 54 && y < 74) return 6;
  }
  return -1;
}

__global__ void __launch_bounds__(THREADS, 1) smoke_rollout_kernel(const Args a) {
  extern __shared__ __align__(16) double sm[];
  double* p_s = sm;                       // [NC] search direction / pressure
  double* xg = a.x_ws + (size_t)blockIdx.x * NC;
  double* s_part = sm + NC + 1;           // reduction partials: 4 slots x 32 x up to 9 values
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const int nx = a.nx, si = 128 / nx, ti = a.T / a.nt;
  double* vel0 = a.vel_ws + (size_t)b * 2 * NV * 2;
  float* dws = a.dens_ws + (size_t)b * 4 * NC;

  // ---- per-thread constants: neighbour bits and diagonal of the masked Laplacian (phi/solver/sparse.py:27-78) ----
  int info[PER];         // bit0 lower y, bit1 upper y, bit2 lower x, bit3 upper x neighbour present; bits 4-6: -diagonal
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int i = tid + THREADS * j;
    info[j] = 1 << 4;
    if (i < NC) {
      const int y = i / N, x = i - y * N;
      const int c = a.fluid[i];
      auto act = [&](int yy, int xx) { return (yy >= 0 && yy < N && xx >= 0 && xx < N) ? (int)a.fluid[yy * N + xx] : 0; };
      auto flu = [&](int yy, int xx) { return (yy >= 0 && yy < N && xx >= 0 && xx < N) ? (int)a.fluid[yy * N + xx] : 1; };
      const int nbits = (act(y - 1, x) * c) | ((act(y + 1, x) * c) << 1) | ((act(y, x - 1) * c) << 2) | ((act(y, x + 1) * c) << 3);
      const int center = (flu(y + 1, x) + flu(y - 1, x)) + (flu(y, x + 1) + flu(y, x - 1));   // = -stencil_center
      info[j] = nbits | ((center > 1 ? center : 1) << 4);                                      // diag = min(-center, -1)
    }
  }
  auto apply_A = [&](int j, int i, double pc) {
    double s = -(double)(info[j] >> 4) * pc;
    if (info[j] & 1) s += p_s[i - N];
    if (info[j] & 2) s += p_s[i + N];
    if (info[j] & 4) s += p_s[i - 1];
    if (info[j] & 8) s += p_s[i + 1];
    return s;
  };

  // ---- frame 0 (es.py:236-270) ----
  double smoke_outs[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) smoke_outs[k] = 0.0;
  {
    const float* iv = a.init_velocity + (size_t)b * NV * 2;
    double* vout = a.velocitys + ((size_t)b * a.T) * NV * 2;
    for (int s = tid; s < NV * 2; s += THREADS) {
      const double v = (double)iv[s];
      vel0[s] = v;
      vout[s] = v;
    }
    const float* d0 = a.init_density + (size_t)b * nx * nx;
    float* dout = a.densitys + ((size_t)b * a.T) * NV;
    float* zout = a.zero_densitys + ((size_t)b * a.T) * NV;
    double red[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) red[k] = 0.0;
    for (int s = tid; s < NV; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      float v = 0.f;
      if (y < N && x < N) {
        v = d0[(y / si) * nx + (x / si)];
        const int bk = bucket_of(y, x);
        if (bk >= 0) red[bk] += (double)v; else red[7] += (double)v;
      }
      dout[s] = v;
    }
    block_reduce<8>(red, false, s_part, 0);
    double bsum = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) bsum += red[k];
    const bool zero_it = bsum > 0.0;
    if (zero_it) {
#pragma unroll
      for (int k = 0; k < 7; ++k) smoke_outs[k] += red[k];
    }
    for (int s = tid; s < NV; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      float v = 0.f;
      if (y < N && x < N) {
        v = d0[(y / si) * nx + (x / si)];
        dws[0 * NC + y * N + x] = v;                       // buffer 0, field 0: density
        if (zero_it && bucket_of(y, x) >= 0) v = 0.f;
        dws[1 * NC + y * N + x] = v;                       // buffer 0, field 1: zeroed density
      }
      zout[s] = v;
    }
    double so = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) so += smoke_outs[k];
    if (tid == 0) {
      a.smoke_out[(size_t)b * a.T] = smoke_outs[1] / (so + red[7]);
      a.iterations[(size_t)b * a.T] = 0;
    }
  }
  __syncthreads();

  int slot = 1;
  for (int frame = 0; frame < a.T - 1; ++frame) {
    const double* vprev = vel0 + (size_t)(frame & 1) * NV * 2;
    double* vcur = vel0 + (size_t)((frame + 1) & 1) * NV * 2;
    const float* c1f = a.c1 + ((size_t)b * a.nt + frame / ti) * nx * nx;
    const float* c2f = a.c2 + ((size_t)b * a.nt + frame / ti) * nx * nx;
    // ---- A. control injection + boundary mask (es.py:128-142, flow.py:294-298) ----
    for (int s = tid; s < NV; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      double vx, vy;
      if (y >= 16 && y < 112 && x >= 16 && x < 112) {
        vx = vprev[2 * s];
        vy = vprev[2 * s + 1];
      } else {
        vx = (double)c1f[(y / si) * nx + (x / si)];
        vy = (double)c2f[(y / si) * nx + (x / si)];
      }
      vcur[2 * s] = vx * (double)a.vmask[2 * s];
      vcur[2 * s + 1] = vy * (double)a.vmask[2 * s + 1];
    }
    __syncthreads();
    // ---- B. divergence -> r (= p: aliased in the reference), x = 0 ----
    double rr[PER];
    double mx = 0.0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + THREADS * j;
      rr[j] = 0.0;
      if (i < NC) {
        const int y = i / N, x = i - y * N;
        const double d = (vcur[2 * ((y + 1) * NS + x) + 1] - vcur[2 * (y * NS + x) + 1]) +
                         (vcur[2 * (y * NS + x + 1)] - vcur[2 * (y * NS + x)]);
        rr[j] = d;
        p_s[i] = d;
        xg[i] = 0.0;
        mx = fmax(mx, fabs(d));
      }
    }
    {
      double red[1] = {mx};
      block_reduce<1>(red, true, s_part, slot);   // its barrier also orders the p_s writes before the stencil reads
      slot = (slot + 1) & 3;
      mx = red[0];
    }
    // ---- C. conjugate gradient (phi/solver/base.py:56-103) ----
    int it = 0;
    while (mx >= a.accuracy && it < a.max_iterations) {
      double red2[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int i = tid + THREADS * j;
        if (i < NC) {
          const double pc = p_s[i];
          const double ap = apply_A(j, i, pc);
          red2[0] += pc * ap;         // tmp = sum(p * Ap)
          red2[1] += pc * rr[j];      // sum(p * r)
        }
      }
      block_reduce<2>(red2, false, s_part, slot);
      slot = (slot + 1) & 3;
      const double tmp = red2[0];
      const double alpha = red2[1] / tmp;
      double red3[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int i = tid + THREADS * j;
        if (i < NC) {
          const double pc = p_s[i];
          const double ap = apply_A(j, i, pc);     // recomputed: three fp64 vectors do not fit on chip
          xg[i] += alpha * pc;
          rr[j] -= alpha * ap;
          red3[0] += rr[j] * ap;
          red3[1] = fmax(red3[1], fabs(rr[j]));
        }
      }
      block_reduce<2>(red3, true, s_part, slot);    // barrier: every stencil read of p_s is done before p is rewritten
      slot = (slot + 1) & 3;
      const double beta = -red3[0] / tmp;
      mx = red3[1];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int i = tid + THREADS * j;
        if (i < NC) {
          // first iteration: the reference's momentum still aliases the (already updated) residual
          const double pold = (it == 0) ? rr[j] : p_s[i];
          p_s[i] = rr[j] + beta * pold;
        }
      }
      __syncthreads();
      ++it;
    }
    // ---- D. pressure to shared memory ----
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + THREADS * j;
      if (i < NC) p_s[i] = xg[i];
    }
    __syncthreads();
    // ---- E. v <- (v - mask * grad p) * mask  (nd.py:602-614 with symmetric padding; flow.py:318-327; es.py:145) ----
    double* vout = a.velocitys + ((size_t)b * a.T + frame + 1) * NV * 2;
    for (int s = tid; s < NV; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      const int yc = min(y, N - 1), xc = min(x, N - 1), ym = max(y - 1, 0), xm = max(x - 1, 0);
      const double pc = p_s[yc * N + xc];
      const double gx = pc - p_s[yc * N + min(xm, N - 1)];
      const double gy = pc - p_s[min(ym, N - 1) * N + xc];
      const double mxk = (double)a.vmask[2 * s], myk = (double)a.vmask[2 * s + 1];
      const double nvx = (vcur[2 * s] - gx * mxk) * mxk;
      const double nvy = (vcur[2 * s + 1] - gy * myk) * myk;
      vcur[2 * s] = nvx;
      vcur[2 * s + 1] = nvy;
      vout[2 * s] = nvx;
      vout[2 * s + 1] = nvy;
    }
    __syncthreads();
    // ---- F. advect density and zeroed density (nd.py:422-427, scipy_backend.py:58-77, :181-185) ----
    const float* din = dws + (size_t)(frame & 1) * 2 * NC;
    float* dnew = dws + (size_t)((frame + 1) & 1) * 2 * NC;
    float zval[PER];
    double red8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) red8[k] = 0.0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + THREADS * j;
      zval[j] = 0.f;
      if (i < NC) {
        const int y = i / N, x = i - y * N;
        const double vyc = (vcur[2 * ((y + 1) * NS + x) + 1] + vcur[2 * (y * NS + x) + 1]) / 2;
        const double vxc = (vcur[2 * (y * NS + x + 1)] + vcur[2 * (y * NS + x)]) / 2;
        const double sy = fmin(fmax((double)y - vyc * a.dt, 0.0), (double)N);
        const double sx = fmin(fmax((double)x - vxc * a.dt, 0.0), (double)N);
        float dv = 0.f, zv = 0.f;
        if (sy <= (double)(N - 1) && sx <= (double)(N - 1)) {
          int y0 = (int)floor(sy), x0 = (int)floor(sx);
          y0 = min(max(y0, 0), N - 2);
          x0 = min(max(x0, 0), N - 2);
          const double fy = sy - y0, fx = sx - x0;
          const double w00 = (1 - fy) * (1 - fx), w10 = fy * (1 - fx), w01 = (1 - fy) * fx, w11 = fy * fx;
          const int o = y0 * N + x0;
          dv = (float)((double)din[o] * w00 + (double)din[o + N] * w10 + (double)din[o + 1] * w01 + (double)din[o + N + 1] * w11);
          zv = (float)((double)din[NC + o] * w00 + (double)din[NC + o + N] * w10 + (double)din[NC + o + 1] * w01 +
                       (double)din[NC + o + N + 1] * w11);
        }
        dnew[i] = dv;
        zval[j] = zv;
        const int bk = bucket_of(y, x);
        if (bk >= 0) red8[bk] += (double)zv; else red8[7] += (double)zv;
      }
    }
    // ---- G. smoke accounting (es.py:279-305) ----
    block_reduce<8>(red8, false, s_part, slot);
    slot = (slot + 1) & 3;
    double bsum = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) bsum += red8[k];
    const bool zero_it = bsum > 0.0;
    if (zero_it) {
#pragma unroll
      for (int k = 0; k < 7; ++k) smoke_outs[k] += red8[k];
    }
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int i = tid + THREADS * j;
      if (i < NC) {
        float zv = zval[j];
        if (zero_it && bucket_of(i / N, i % N) >= 0) zv = 0.f;
        dnew[NC + i] = zv;
      }
    }
    __syncthreads();
    // ---- H. frame outputs ----
    float* dout = a.densitys + ((size_t)b * a.T + frame + 1) * NV;
    float* zout = a.zero_densitys + ((size_t)b * a.T + frame + 1) * NV;
    for (int s = tid; s < NV; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      const bool in = (y < N && x < N);
      dout[s] = in ? dnew[y * N + x] : 0.f;
      zout[s] = in ? dnew[NC + y * N + x] : 0.f;
    }
    double so = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) so += smoke_outs[k];
    if (tid == 0) {
      a.smoke_out[(size_t)b * a.T + frame + 1] = smoke_outs[1] / (so + red8[7]);
      a.iterations[(size_t)b * a.T + frame + 1] = it;
    }
    __syncthreads();
  }
}

}  // namespace rollout
}  // namespace dpc

extern "C" int dpc_smoke_rollout(const int8_t* fluid_mask, const float* velocity_mask, const float* init_velocity,
                                 const float* init_density, const float* c1, const float* c2, double* vel_ws,
                                 double* x_ws, float* dens_ws, float* densitys, float* zero_densitys, double* velocitys,
                                 double* smoke_out, int32_t* iterations, int32_t B, int32_t nt, int32_t nx, int32_t T,
                                 double dt, double accuracy, int32_t max_iterations, void* stream) {
  using namespace dpc;
  using namespace dpc::rollout;
  DPC_CHECK_ARG(fluid_mask && velocity_mask && init_velocity && init_density && c1 && c2 && vel_ws && x_ws && dens_ws);
  DPC_CHECK_ARG(densitys && zero_densitys && velocitys && smoke_out && iterations);
  DPC_CHECK_ARG(B > 0 && nt > 0 && nx > 0 && 128 % nx == 0 && T >= 1 && T % nt == 0 && max_iterations >= 0);
  Args a;
  a.fluid = fluid_mask; a.vmask = velocity_mask; a.init_velocity = init_velocity; a.init_density = init_density;
  a.c1 = c1; a.c2 = c2; a.vel_ws = vel_ws; a.x_ws = x_ws; a.dens_ws = dens_ws; a.densitys = densitys; a.zero_densitys = zero_densitys;
  a.velocitys = velocitys; a.smoke_out = smoke_out; a.iterations = iterations;
  a.nt = nt; a.nx = nx; a.T = T; a.dt = dt; a.accuracy = accuracy; a.max_iterations = max_iterations;
  const size_t smem = (size_t)(NC + 1 + 4 * 32 * 9) * sizeof(double);
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[device_ordinal()];
  if (!configured) {
    DPC_CUDA(cudaFuncSetAttribute(smoke_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  smoke_rollout_kernel<<<(unsigned)B, THREADS, smem, (cudaStream_t)stream>>>(a);
  DPC_LAUNCH_CHECK();
  return 0;
}
