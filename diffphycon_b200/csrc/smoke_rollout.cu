// Post-sampling smoke rollout: one persistent CTA per trajectory runs all simulation steps of
// dataset/apps/evaluate_solver.py::solver (es.py:205-310) on chip — control injection (es.py:128-142), masking,
// divergence (phi/math/nd.py:367-377), the reference's plain conjugate-gradient pressure solve (phi/solver/base.py:56-103,
// <= 500 iterations, max|r| >= 1e-8 stop test, including its first-iteration aliasing of `momentum` and `residual`),
// pressure-gradient subtraction (nd.py:602-614), two semi-Lagrangian advections with the reference's clamp/zero-fill
// rule (nd.py:422-427, scipy_backend.py:58-77, :181-185) and the smoke-in-bucket accounting (es.py:279-305).
//
// Layout: a thread-block CLUSTER of CS = 2, 4 or 8 CTAs per trajectory (512 threads each); CTA `rk` owns a band of
// RPC = ceil(127 / CS) rows of the 127x127 pressure grid (and the same rows of the staggered 128x128 fields), cell
// l = tid + 512*j (j < 16, 8, 4) of the band is owned by one thread.  Everything the conjugate-gradient iteration touches is
// on chip: the residual r and A.p of a thread's cells live in registers (fp64, like the reference's NumPy run), the search
// direction p of the band plus one halo row above and below and the solution x of the band live in shared memory.  After
// every p update the first / last band row is pushed into the neighbour CTA's halo through distributed shared memory; the
// three dot products and the max|r| stop test are reduced per CTA (warp shuffles + one shared-memory exchange, fixed tree),
// the CTA totals are pushed to every CTA of the cluster and summed there in rank order (deterministic; the summation order
// depends on CS only).  One stencil pass, two cluster reductions and one halo barrier per iteration.  The Laplacian is
// applied from 4 neighbour bits + a small integer diagonal per cell (no matrix is materialised — the reference rebuilds a
// scipy sparse matrix every step).  Round 1 ran one CTA per trajectory with x in an L2 workspace and two stencil passes
// (three fp64 vectors do not fit one SM): 18 us per iteration; this kernel: 6.8 us (CS = 2, 64 trajectories), 5.4 us (CS = 4,
// 32 trajectories), of which ~2.4 us are the three cluster barriers (tools/cluster_sync_probe.cu) and the rest is bounded by the
// shared-memory passes over fp64 vectors (two wavefronts per access) — measured with DPC_ROLLOUT_PROF=1.  The velocity / density fields of the other phases stay in
// global workspaces (L2-resident); a cluster barrier (release / acquire) orders the phases.
// The kernel is latency bound (127 500 strictly sequential CG iterations per trajectory), not HBM bound.
#include "common.cuh"

#include <cooperative_groups.h>

namespace dpc {
namespace rollout {

namespace cg = cooperative_groups;

constexpr int N = 127;
constexpr int NC = N * N;          // 16129 pressure cells
constexpr int NS = 128;
constexpr int NV = NS * NS;        // staggered samples per component

struct Args {
  const int8_t* fluid;             // [127][127]
  const float* vmask;              // [128][128][2]
  const float* init_velocity;      // [B][128][128][2]
  const float* init_density;       // [B][nx][nx]
  const float* c1;                 // [B][nt][nx][nx]
  const float* c2;
  double* vel_ws;                  // [B][2][128][128][2] ping-pong
  double* x_ws;                    // unused since round 2 (the CG solution lives in shared memory); kept in the ABI
  float* dens_ws;                  // [B][2][2][127][127]: (buffer, field) ping-pong for density / zeroed density
  float* densitys;                 // [B][T][128][128]
  float* zero_densitys;            // [B][T][128][128]
  double* velocitys;               // [B][T][128][128][2]
  double* smoke_out;               // [B][T]
  int32_t* iterations;             // [B][T] CG iterations per step (0 for frame 0)
  int nt, nx, T;
  double dt, accuracy;
  int max_iterations;
  int prof;                        // development: per-phase clock64 totals of the CG loop printed by trajectory 0 (DPC_ROLLOUT_PROF)
};

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

constexpr int RED_SLOTS = 4, RED_MAXV = 9;

struct NoPost {
  __device__ __forceinline__ void operator()(double*) const {}
};

// cluster-wide sum / max of NVAL values; every thread of every CTA gets the same result.  `slot` cycles so that back-to-back
// calls never overwrite partials that a slower warp / CTA is still reading (a CTA is at most one barrier ahead).
// The vector FP64 pipe of this part issues one warp instruction per 2 clk per SM sub-partition (measured: 16 lanes/clk/SM), so
// everything that is the same for the whole CTA runs in ONE thread: the rank-order sum of the CTA totals and `post` (the
// alpha / beta divisions of the CG iteration) are done by thread 0 and broadcast through shared memory.
template <int CS, int THREADS, int NVAL, class Post = NoPost>
__device__ __forceinline__ void cluster_reduce(double (&v)[NVAL], bool last_is_max, double* s_part, double* s_cl, int slot, int rk,
                                               Post post = Post()) {
  cg::cluster_group cluster = cg::this_cluster();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NVAL; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double other = shfl_xor_double(v[k], o);
      v[k] = (last_is_max && k == NVAL - 1) ? fmax(v[k], other) : v[k] + other;
    }
  }
  double* part = s_part + slot * (32 * RED_MAXV);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NVAL; ++k) part[k * 32 + warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NVAL; ++k) {
      double t = (lane < THREADS / 32) ? part[k * 32 + lane] : 0.0;   // identities: 0 for sums and for max|.|
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double other = shfl_xor_double(t, o);
        t = (last_is_max && k == NVAL - 1) ? fmax(t, other) : t + other;
      }
      // lane c pushes this CTA's total into CTA c's table (distributed shared memory)
      if (lane < CS) cluster.map_shared_rank(s_cl, lane)[(slot * RED_MAXV + k) * 8 + rk] = t;
    }
  }
  cluster.sync();
  double* bc = s_cl + RED_SLOTS * RED_MAXV * 8 + slot * RED_MAXV;   // broadcast slot of this call
  if (threadIdx.x == 0) {
    double tot[NVAL];
#pragma unroll
    for (int k = 0; k < NVAL; ++k) {
      const double* row = s_cl + (slot * RED_MAXV + k) * 8;
      double t = row[0];
#pragma unroll
      for (int c = 1; c < CS; ++c) t = (last_is_max && k == NVAL - 1) ? fmax(t, row[c]) : t + row[c];
      tot[k] = t;
    }
    post(tot);
#pragma unroll
    for (int k = 0; k < NVAL; ++k) bc[k] = tot[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NVAL; ++k) v[k] = bc[k];
}

template <int CS, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) smoke_rollout_kernel(const Args a) {
  constexpr int RPC = (N + CS - 1) / CS;     // pressure rows per CTA (the last CTA may own fewer)
  constexpr int PER = (RPC * N + THREADS - 1) / THREADS;   // cells per thread
  static_assert(PER * THREADS >= RPC * N, "band does not fit");
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double sm[];
  double* p_s = sm;                                   // [(RPC + 2) * N] search direction / pressure: halo row, band, halo row
  double* x_s = p_s + (RPC + 2) * N + 1;              // [RPC * N] CG solution of the band
  double* s_part = x_s + RPC * N + 1;                 // block-level partials: RED_SLOTS x 32 x RED_MAXV
  double* s_cl = s_part + RED_SLOTS * 32 * RED_MAXV;  // cluster table: RED_SLOTS x RED_MAXV x 8 ranks, then RED_SLOTS x RED_MAXV broadcast slots
  double* diag_s = s_cl + RED_SLOTS * RED_MAXV * 8 + RED_SLOTS * RED_MAXV;   // [8]: -diagonal as a double (an int -> double conversion is an FP64-pipe instruction)
  if (threadIdx.x < 8) diag_s[threadIdx.x] = -(double)threadIdx.x;
  const int tid = threadIdx.x;
  const int rk = (int)cluster.block_rank();
  const int b = blockIdx.x / CS;
  const int r0 = rk * RPC, r1 = min(N, r0 + RPC), nrows = r1 - r0, ncell = nrows * N;
  const int s_lo = r0 * NS, s_hi = (rk == CS - 1 ? NS : r1) * NS;   // this CTA's rows of the staggered 128 x 128 fields
  const int nx = a.nx, si = 128 / nx, ti = a.T / a.nt;
  double* vel0 = a.vel_ws + (size_t)b * 2 * NV * 2;
  float* dws = a.dens_ws + (size_t)b * 4 * NC;
  double* p_up = rk > 0 ? cluster.map_shared_rank(p_s, rk - 1) : nullptr;        // neighbour above: its lower halo row = my first row
  double* p_dn = rk < CS - 1 ? cluster.map_shared_rank(p_s, rk + 1) : nullptr;   // neighbour below: its upper halo row = my last row

  // ---- per-thread constants: neighbour bits and diagonal of the masked Laplacian (phi/solver/sparse.py:27-78) ----
  int info[PER];         // bit0 lower y, bit1 upper y, bit2 lower x, bit3 upper x neighbour present; bits 4-6: -diagonal
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int l = tid + THREADS * j;
    info[j] = 1 << 4;
    if (l < ncell) {
      const int i = r0 * N + l;
      const int y = i / N, x = i - y * N;
      const int c = a.fluid[i];
      auto act = [&](int yy, int xx) { return (yy >= 0 && yy < N && xx >= 0 && xx < N) ? (int)a.fluid[yy * N + xx] : 0; };
      auto flu = [&](int yy, int xx) { return (yy >= 0 && yy < N && xx >= 0 && xx < N) ? (int)a.fluid[yy * N + xx] : 1; };
      const int nbits = (act(y - 1, x) * c) | ((act(y + 1, x) * c) << 1) | ((act(y, x - 1) * c) << 2) | ((act(y, x + 1) * c) << 3);
      const int center = (flu(y + 1, x) + flu(y - 1, x)) + (flu(y, x + 1) + flu(y, x - 1));   // = -stencil_center
      info[j] = nbits | ((center > 1 ? center : 1) << 4);                                      // diag = min(-center, -1)
    }
  }
  // p of band cell l sits at p_s[l + N] (one halo row in front)
  // branch-free: the four neighbour loads issue back to back (always inside p_s: halo rows / row wrap-around), absent
  // neighbours are SELECTED out (never multiplied: a halo row outside the domain is never written and may hold anything);
  // 3 adds + 1 fma per cell on the FP64 pipe
  auto apply_A = [&](int j, int l, double pc) {
    const double up = p_s[l], dn = p_s[l + 2 * N], lf = p_s[l + N - 1], rt = p_s[l + N + 1];
    const double a0 = ((info[j] & 1) ? up : 0.0) + ((info[j] & 2) ? dn : 0.0);
    const double a1 = ((info[j] & 4) ? lf : 0.0) + ((info[j] & 8) ? rt : 0.0);
    return fma(diag_s[info[j] >> 4], pc, a0 + a1);
  };
  // store p of band cell l; the first / last band row also goes into the neighbour's halo row
  auto put_p = [&](int l, double v) {
    p_s[l + N] = v;
    if (p_up && l < N) p_up[(RPC + 1) * N + l] = v;                      // neighbours above always own RPC rows
    if (p_dn && l >= ncell - N) p_dn[l - (ncell - N)] = v;
  };
  long long prof_t[7] = {0, 0, 0, 0, 0, 0, 0}, prof_f[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tf0 = 0;
  auto flap = [&](int i) { if (a.prof) { const long long t = clock64(); prof_f[i] += t - tf0; tf0 = t; } };
  int slot = 0;
  auto next_slot = [&]() { const int s0 = slot; slot = (slot + 1) & (RED_SLOTS - 1); return s0; };

  // ---- frame 0 (es.py:236-270) ----
  double smoke_outs[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) smoke_outs[k] = 0.0;
  {
    const float* iv = a.init_velocity + (size_t)b * NV * 2;
    double* vout = a.velocitys + ((size_t)b * a.T) * NV * 2;
    for (int s = 2 * s_lo + tid; s < 2 * s_hi; s += THREADS) {
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
    for (int s = s_lo + tid; s < s_hi; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      float v = 0.f;
      if (y < N && x < N) {
        v = d0[(y / si) * nx + (x / si)];
        const int bk = bucket_of(y, x);
        if (bk >= 0) red[bk] += (double)v; else red[7] += (double)v;
      }
      dout[s] = v;
    }
    cluster_reduce<CS, THREADS, 8>(red, false, s_part, s_cl, next_slot(), rk);
    double bsum = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) bsum += red[k];
    const bool zero_it = bsum > 0.0;
    if (zero_it) {
#pragma unroll
      for (int k = 0; k < 7; ++k) smoke_outs[k] += red[k];
    }
    for (int s = s_lo + tid; s < s_hi; s += THREADS) {
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
    if (tid == 0 && rk == 0) {
      a.smoke_out[(size_t)b * a.T] = smoke_outs[1] / (so + red[7]);
      a.iterations[(size_t)b * a.T] = 0;
    }
  }
  cluster.sync();

  for (int frame = 0; frame < a.T - 1; ++frame) {
    const double* vprev = vel0 + (size_t)(frame & 1) * NV * 2;
    double* vcur = vel0 + (size_t)((frame + 1) & 1) * NV * 2;
    const float* c1f = a.c1 + ((size_t)b * a.nt + frame / ti) * nx * nx;
    const float* c2f = a.c2 + ((size_t)b * a.nt + frame / ti) * nx * nx;
    if (a.prof) tf0 = clock64();
    // ---- A. control injection + boundary mask (es.py:128-142, flow.py:294-298) ----
    for (int s = s_lo + tid; s < s_hi; s += THREADS) {
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
    flap(0);
    cluster.sync();                                       // the divergence reads the row below (next CTA's band)
    flap(1);
    // ---- B. divergence -> r (= p: aliased in the reference), x = 0 ----
    double rr[PER];
    double mx = 0.0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int l = tid + THREADS * j;
      rr[j] = 0.0;
      if (l < ncell) {
        const int i = r0 * N + l;
        const int y = i / N, x = i - y * N;
        const double d = (vcur[2 * ((y + 1) * NS + x) + 1] - vcur[2 * (y * NS + x) + 1]) +
                         (vcur[2 * (y * NS + x + 1)] - vcur[2 * (y * NS + x)]);
        rr[j] = d;
        put_p(l, d);
        x_s[l] = 0.0;
        mx = fmax(mx, fabs(d));
      }
    }
    {
      double red[1] = {mx};
      cluster_reduce<CS, THREADS, 1>(red, true, s_part, s_cl, next_slot(), rk);   // its barrier also orders the p writes (band + halos) before the stencil reads
      mx = red[0];
    }
    flap(2);
    // ---- C. conjugate gradient (phi/solver/base.py:56-103) ----
    int it = 0;
    long long tk0 = 0;
    while (mx >= a.accuracy && it < a.max_iterations) {
      if (a.prof) tk0 = clock64();
      double ap[PER];
      double red2[2] = {0.0, 0.0};
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int l = tid + THREADS * j;
        ap[j] = 0.0;
        if (l < ncell) {
          const double pc = p_s[l + N];
          ap[j] = apply_A(j, l, pc);
          red2[0] = fma(pc, ap[j], red2[0]);      // tmp = sum(p * Ap)
          red2[1] = fma(pc, rr[j], red2[1]);      // sum(p * r)
        }
      }
      if (a.prof) { const long long t = clock64(); prof_t[0] += t - tk0; tk0 = t; }
      cluster_reduce<CS, THREADS, 2>(red2, false, s_part, s_cl, next_slot(), rk, [](double* t) { t[1] = t[1] / t[0]; });
      if (a.prof) { const long long t = clock64(); prof_t[1] += t - tk0; tk0 = t; }
      const double tmp = red2[0];
      const double alpha = red2[1];           // sum(p * r) / tmp
      double red3[2] = {0.0, 0.0};
      unsigned long long amax = 0ull;         // max|r| on the integer pipe: non-negative doubles order like their bit patterns
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int l = tid + THREADS * j;
        if (l < ncell) {
          x_s[l] = fma(alpha, p_s[l + N], x_s[l]);
          rr[j] = fma(-alpha, ap[j], rr[j]);
          red3[0] = fma(rr[j], ap[j], red3[0]);
          const unsigned long long ar = (unsigned long long)__double_as_longlong(rr[j]) & 0x7fffffffffffffffull;
          amax = ar > amax ? ar : amax;
        }
      }
      red3[1] = __longlong_as_double((long long)amax);
      if (a.prof) { const long long t = clock64(); prof_t[2] += t - tk0; tk0 = t; }
      cluster_reduce<CS, THREADS, 2>(red3, true, s_part, s_cl, next_slot(), rk,
                                     [tmp](double* t) { t[0] = -t[0] / tmp; });   // barrier: every stencil read of p is done before p is rewritten
      if (a.prof) { const long long t = clock64(); prof_t[3] += t - tk0; tk0 = t; }
      const double beta = red3[0];            // -sum(r * Ap) / tmp
      mx = red3[1];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int l = tid + THREADS * j;
        if (l < ncell) {
          // first iteration: the reference's momentum still aliases the (already updated) residual
          const double pold = (it == 0) ? rr[j] : p_s[l + N];
          put_p(l, fma(beta, pold, rr[j]));
        }
      }
      if (a.prof) { const long long t = clock64(); prof_t[4] += t - tk0; tk0 = t; }
      cluster.sync();                                     // band and halo rows of the new p are in place
      if (a.prof) { const long long t = clock64(); prof_t[5] += t - tk0; prof_t[6] += 1; }
      ++it;
    }
    flap(3);
    // ---- D. pressure to shared memory (band + halos: the gradient reads the row above) ----
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int l = tid + THREADS * j;
      if (l < ncell) put_p(l, x_s[l]);
    }
    cluster.sync();
    // ---- E. v <- (v - mask * grad p) * mask  (nd.py:602-614 with symmetric padding; flow.py:318-327; es.py:145) ----
    double* vout = a.velocitys + ((size_t)b * a.T + frame + 1) * NV * 2;
    auto P = [&](int yy, int xx) { return p_s[(yy - r0 + 1) * N + xx]; };   // yy in [r0 - 1, r1]
    for (int s = s_lo + tid; s < s_hi; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      const int yc = min(y, N - 1), xc = min(x, N - 1), ym = max(y - 1, 0), xm = max(x - 1, 0);
      const double pc = P(yc, xc);
      const double gx = pc - P(yc, min(xm, N - 1));
      const double gy = pc - P(min(ym, N - 1), xc);
      const double mxk = (double)a.vmask[2 * s], myk = (double)a.vmask[2 * s + 1];
      const double nvx = (vcur[2 * s] - gx * mxk) * mxk;
      const double nvy = (vcur[2 * s + 1] - gy * myk) * myk;
      vcur[2 * s] = nvx;
      vcur[2 * s + 1] = nvy;
      vout[2 * s] = nvx;
      vout[2 * s + 1] = nvy;
    }
    flap(4);
    cluster.sync();                                       // the advection reads the row below
    // ---- F. advect density and zeroed density (nd.py:422-427, scipy_backend.py:58-77, :181-185) ----
    const float* din = dws + (size_t)(frame & 1) * 2 * NC;
    float* dnew = dws + (size_t)((frame + 1) & 1) * 2 * NC;
    float zval[PER];
    double red8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) red8[k] = 0.0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
      const int l = tid + THREADS * j;
      zval[j] = 0.f;
      if (l < ncell) {
        const int i = r0 * N + l;
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
    flap(5);
    // ---- G. smoke accounting (es.py:279-305) ----
    cluster_reduce<CS, THREADS, 8>(red8, false, s_part, s_cl, next_slot(), rk);
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
      const int l = tid + THREADS * j;
      if (l < ncell) {
        const int i = r0 * N + l;
        float zv = zval[j];
        if (zero_it && bucket_of(i / N, i % N) >= 0) zv = 0.f;
        dnew[NC + i] = zv;
      }
    }
    flap(6);
    __syncthreads();                                      // the frame outputs below read this CTA's own band only
    // ---- H. frame outputs ----
    float* dout = a.densitys + ((size_t)b * a.T + frame + 1) * NV;
    float* zout = a.zero_densitys + ((size_t)b * a.T + frame + 1) * NV;
    for (int s = s_lo + tid; s < s_hi; s += THREADS) {
      const int y = s >> 7, x = s & 127;
      const bool in = (y < N && x < N);
      dout[s] = in ? dnew[y * N + x] : 0.f;
      zout[s] = in ? dnew[NC + y * N + x] : 0.f;
    }
    double so = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) so += smoke_outs[k];
    if (tid == 0 && rk == 0) {
      a.smoke_out[(size_t)b * a.T + frame + 1] = smoke_outs[1] / (so + red8[7]);
      a.iterations[(size_t)b * a.T + frame + 1] = it;
    }
    cluster.sync();                                       // next frame: the advection gathers from every band of dnew
    flap(7);
  }
  if (a.prof && b == 0 && tid == 0 && rk == 0)
    printf("rollout frame phases (total clk over %d frames): A %lld, sync %lld, B div+reduce %lld, C cg %lld, D+E %lld, sync+F %lld, G %lld, H+sync %lld\n", a.T - 1,
           prof_f[0], prof_f[1], prof_f[2], prof_f[3], prof_f[4], prof_f[5], prof_f[6], prof_f[7]);
  if (a.prof && b == 0 && tid == 0 && prof_t[6] > 0)
    printf("rollout CG rank %d of %d: per iteration clk: stencil+dots %lld, reduce1 %lld, update %lld, reduce2 %lld, p update %lld, halo sync %lld (%lld iterations)\n",
           rk, CS, prof_t[0] / prof_t[6], prof_t[1] / prof_t[6], prof_t[2] / prof_t[6], prof_t[3] / prof_t[6], prof_t[4] / prof_t[6], prof_t[5] / prof_t[6], prof_t[6]);
}

template <int CS, int THREADS>
static int launch(const Args& a, int B, cudaStream_t st) {
  constexpr int RPC = (N + CS - 1) / CS;
  const size_t smem = (size_t)((RPC + 2) * N + 1 + RPC * N + 1 + RED_SLOTS * 32 * RED_MAXV + RED_SLOTS * RED_MAXV * 8 + RED_SLOTS * RED_MAXV + 8) * sizeof(double);
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[device_ordinal()];
  if (!configured) {
    DPC_CUDA(cudaFuncSetAttribute(smoke_rollout_kernel<CS, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(B * CS));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  DPC_CUDA(cudaLaunchKernelEx(&cfg, smoke_rollout_kernel<CS, THREADS>, a));
  DPC_LAUNCH_CHECK();
  return 0;
}

// co-resident clusters of CS CTAs on this device (cached per device): a cluster must fit one GPC, so this is well below
// num_SMs / CS for CS = 8
template <int CS, int THREADS>
static int max_clusters() {
  static int cached[kMaxDevices] = {};
  int& n = cached[device_ordinal()];
  if (n) return n;
  constexpr int RPC = (N + CS - 1) / CS;
  const size_t smem = (size_t)((RPC + 2) * N + 1 + RPC * N + 1 + RED_SLOTS * 32 * RED_MAXV + RED_SLOTS * RED_MAXV * 8 + RED_SLOTS * RED_MAXV + 8) * sizeof(double);
  if (cudaFuncSetAttribute(smoke_rollout_kernel<CS, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return n = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CS * 64));
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int v = 0;
  if (cudaOccupancyMaxActiveClusters(&v, smoke_rollout_kernel<CS, THREADS>, &cfg) != cudaSuccess || v < 1) v = 1;
  return n = v;
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
  { static const int prof = getenv("DPC_ROLLOUT_PROF") ? atoi(getenv("DPC_ROLLOUT_PROF")) : 0; a.prof = prof; }
  // CTAs per trajectory: 4 while every trajectory still runs in one wave (cudaOccupancyMaxActiveClusters), else 2
  // (DPC_ROLLOUT_CLUSTER = 2 / 4 / 8 overrides; eight-CTA clusters must fit one GPC: ~8 co-resident on a B200 = two waves at 16).
  // Measured wall time for 32 frames (tools/time_rollout.py): 16 trajectories 161 / 111 / 149 ms for 2 / 4 / 8 CTAs, 32 trajectories
  // 121 / 86 / 182 ms, 64 trajectories 109 ms with 2: eight-CTA clusters pay more in the three barriers per iteration than they
  // gain.  The reductions sum the CTA totals in rank order, so results depend on the cluster size at the 1e-16 level only.
  int cs = (B <= max_clusters<4, 512>()) ? 4 : 2;
  if (const char* e = getenv("DPC_ROLLOUT_CLUSTER")) { const int v = atoi(e); if (v == 2 || v == 4 || v == 8) cs = v; }
  if (cs == 8) return launch<8, 512>(a, B, (cudaStream_t)stream);
  if (cs == 4) return launch<4, 512>(a, B, (cudaStream_t)stream);
  // two CTAs per trajectory: 512 threads x 16 cells measured 5 % faster than 1024 x 8 (shared-memory bandwidth bound either way)
  static const int t2 = getenv("DPC_ROLLOUT_T2") ? atoi(getenv("DPC_ROLLOUT_T2")) : 512;
  if (t2 == 1024) return launch<2, 1024>(a, B, (cudaStream_t)stream);
  return launch<2, 512>(a, B, (cudaStream_t)stream);
}
