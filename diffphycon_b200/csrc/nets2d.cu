// Backward (gradient w.r.t. activations / conditioning — no weight gradients) and helper kernels of the jellyfish surrogate
// networks, the 2-D `Unet` boundary updater and `ForceUnet` of diffusion/diffusion_2d_jellyfish.py:276-481, which the guidance
// `force_fn` (inference/inference_2d_jellyfish.py:85-114) differentiates through every denoising step.
// All tensors are channels-last fp32 [N images][HW pixels][C].  The contractions of the backward pass (3x3 / 1x1 / 7x7 dgrad
// convolutions) reuse the implicit-GEMM conv kernels with transposed, spatially flipped weights; this file holds the rest:
//   dpc_gn_silu_bwd            GroupNorm -> (scale+1, shift) -> SiLU backward (jf.py:189-204) + d(scale), d(shift)
//   dpc_layernorm_channels_bwd channel LayerNorm (gain only) backward (jf.py:122-132)
//   dpc_linattn2d_bwd          LinearAttention core backward (jf.py:206-225)
//   dpc_attention2d_bwd        softmax Attention core backward (jf.py:241-255)
//   dpc_add / dpc_sumpool2x2 / dpc_mean_head / dpc_mean_head_bwd / dpc_time_embed_f32 / dpc_time_mlp_bwd
// fp32 SIMT arithmetic (these are < 6 % of the networks' FLOPs); reductions over pixels accumulate in double.
#include "common.cuh"

namespace dpc {
namespace n2d {

constexpr int DH = 32;

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm + scale/shift + SiLU backward.
//   forward:  xh = (y - mean_g) * rstd_g ; z = xh*gamma + beta ; u = z*(scale+1) + shift ; out = silu(u)
//   backward: dU = dout * silu'(u) ; S1_c = sum_p dU ; S3_c = sum_p dU*xh
//             dshift_c = S1_c ; dscale_c = gamma_c*S3_c + beta_c*S1_c
//             dxh = dU*(scale_c+1)*gamma_c ; dy = rstd_g * (dxh - mean_g(dxh) - xh * mean_g(dxh*xh))
// pass 1 (reduce) fills sums[n][c][2] (double); pass 2 (apply) forms the group means from them and writes dy (and dss).
// ------------------------------------------------------------------------------------------------------------------
struct GnArgs {
  const float* y; const double* stats; const float* gamma; const float* beta; const float* ss; int64_t ss_stride, ss_off;
  const float* dout; float* dy; double* sums; float* dss; int64_t rows; int C, groups; float eps; int64_t rows_per_cta;
};

__device__ __forceinline__ void gn_channel_setup(const GnArgs& a, int n, int c, float& mean, float& rstd, float& g, float& b,
                                                 float& sc, float& sh) {
  const int cpg = a.C / a.groups;
  const double inv_n = 1.0 / ((double)a.rows * (double)cpg);
  const int grp = c / cpg;
  const double s = a.stats[((size_t)n * a.groups + grp) * 2 + 0];
  const double q = a.stats[((size_t)n * a.groups + grp) * 2 + 1];
  const double m = s * inv_n;
  double var = q * inv_n - m * m;
  if (var < 0.0) var = 0.0;
  mean = (float)m;
  rstd = (float)(1.0 / sqrt(var + (double)a.eps));
  g = a.gamma[c];
  b = a.beta[c];
  if (a.ss) {
    sc = a.ss[(size_t)n * a.ss_stride + a.ss_off + c] + 1.0f;
    sh = a.ss[(size_t)n * a.ss_stride + a.ss_off + a.C + c];
  } else {
    sc = 1.0f;
    sh = 0.0f;
  }
}

__global__ void __launch_bounds__(256) gn_silu_bwd_reduce_kernel(const GnArgs a) {
  __shared__ double s_red[256][2];
  const int n = blockIdx.y;
  const int c4n = a.C >> 2;                 // float4 lanes per row; 256 % c4n == 0 (checked on the host)
  const int cq = (threadIdx.x % c4n) * 4;
  const int rsub = threadIdx.x / c4n, rstep = 256 / c4n;
  float mean[4], rstd[4], g[4], b[4], sc[4], sh[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) gn_channel_setup(a, n, cq + k, mean[k], rstd[k], g[k], b[k], sc[k], sh[k]);
  const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
  int64_t r1 = r0 + a.rows_per_cta;
  if (r1 > a.rows) r1 = a.rows;
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s3[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t r = r0 + rsub; r < r1; r += rstep) {
    const size_t off = ((size_t)n * a.rows + r) * a.C + cq;
    const float4 yv = __ldg(reinterpret_cast<const float4*>(a.y + off));
    const float4 dv = __ldg(reinterpret_cast<const float4*>(a.dout + off));
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (yy[k] - mean[k]) * rstd[k];
      const float u = fmaf(fmaf(xh, g[k], b[k]), sc[k], sh[k]);
      const float sg = sigmoid_f(u);
      const float dU = dd[k] * sg * (1.0f + u * (1.0f - sg));
      s1[k] += dU;
      s3[k] = fmaf(dU, xh, s3[k]);
    }
  }
  // combine the rstep partial sums of every channel inside the CTA, then one double atomic per (channel, sum)
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __syncthreads();
    s_red[threadIdx.x][0] = (double)s1[k];
    s_red[threadIdx.x][1] = (double)s3[k];
    __syncthreads();
    if (rsub == 0) {
      double t0 = 0.0, t1 = 0.0;
      for (int j = 0; j < rstep; ++j) {
        t0 += s_red[threadIdx.x + j * c4n][0];
        t1 += s_red[threadIdx.x + j * c4n][1];
      }
      atomicAdd(a.sums + ((size_t)n * a.C + cq + k) * 2 + 0, t0);
      atomicAdd(a.sums + ((size_t)n * a.C + cq + k) * 2 + 1, t1);
    }
  }
}

__global__ void __launch_bounds__(256) gn_silu_bwd_apply_kernel(const GnArgs a) {
  extern __shared__ float sm[];             // [C] k*S1, [C] k*S3, [groups] A, [groups] B
  float* s_a = sm;
  float* s_b = sm + a.C;
  float* s_ga = sm + 2 * a.C;
  float* s_gb = s_ga + a.groups;
  const int n = blockIdx.y;
  const int cpg = a.C / a.groups;
  for (int c = threadIdx.x; c < a.C; c += 256) {
    float mean, rstd, g, b, sc, sh;
    gn_channel_setup(a, n, c, mean, rstd, g, b, sc, sh);
    const double S1 = a.sums[((size_t)n * a.C + c) * 2 + 0], S3 = a.sums[((size_t)n * a.C + c) * 2 + 1];
    const float k = sc * g;
    s_a[c] = (float)(S1 * (double)k);
    s_b[c] = (float)(S3 * (double)k);
    if (a.dss && blockIdx.x == 0) {
      a.dss[(size_t)n * a.ss_stride + a.ss_off + c] = (float)((double)g * S3 + (double)b * S1);     // d scale
      a.dss[(size_t)n * a.ss_stride + a.ss_off + a.C + c] = (float)S1;                             // d shift
    }
  }
  __syncthreads();
  if (threadIdx.x < a.groups) {
    double ta = 0.0, tb = 0.0;
    for (int j = 0; j < cpg; ++j) {
      ta += (double)s_a[threadIdx.x * cpg + j];
      tb += (double)s_b[threadIdx.x * cpg + j];
    }
    const double inv_m = 1.0 / ((double)a.rows * (double)cpg);
    s_ga[threadIdx.x] = (float)(ta * inv_m);
    s_gb[threadIdx.x] = (float)(tb * inv_m);
  }
  __syncthreads();
  const int c4n = a.C >> 2;
  const int cq = (threadIdx.x % c4n) * 4;
  const int rsub = threadIdx.x / c4n, rstep = 256 / c4n;
  float mean[4], rstd[4], g[4], b[4], sc[4], sh[4], ga[4], gb[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    gn_channel_setup(a, n, cq + k, mean[k], rstd[k], g[k], b[k], sc[k], sh[k]);
    ga[k] = s_ga[(cq + k) / cpg];
    gb[k] = s_gb[(cq + k) / cpg];
  }
  const int64_t r0 = (int64_t)blockIdx.x * a.rows_per_cta;
  int64_t r1 = r0 + a.rows_per_cta;
  if (r1 > a.rows) r1 = a.rows;
  for (int64_t r = r0 + rsub; r < r1; r += rstep) {
    const size_t off = ((size_t)n * a.rows + r) * a.C + cq;
    const float4 yv = __ldg(reinterpret_cast<const float4*>(a.y + off));
    const float4 dv = __ldg(reinterpret_cast<const float4*>(a.dout + off));
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w}, dd[4] = {dv.x, dv.y, dv.z, dv.w};
    float o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float xh = (yy[k] - mean[k]) * rstd[k];
      const float u = fmaf(fmaf(xh, g[k], b[k]), sc[k], sh[k]);
      const float sg = sigmoid_f(u);
      const float dU = dd[k] * sg * (1.0f + u * (1.0f - sg));
      const float dxh = dU * sc[k] * g[k];
      o[k] = rstd[k] * (dxh - ga[k] - xh * gb[k]);
    }
    *reinterpret_cast<float4*>(a.dy + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Channel LayerNorm backward, one warp per row (C <= 1024, C % 32 == 0):
//   y = (x-mean)*rstd*g ;  a = g*dy ;  dx = rstd*(a - mean_c(a) - xh*mean_c(a*xh))  (+ add)
// ------------------------------------------------------------------------------------------------------------------
template <int PER>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ dy,
                     const float* add, float* dx, int64_t rows, int C, float eps, int use_rsqrt) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  const float* dr = dy + (size_t)row * C;
  float xv[PER], av[PER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    xv[i] = c < C ? xr[c] : 0.f;
    s += xv[i];
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    const float d = c < C ? xv[i] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float var = warp_sum(q) / (float)C;
  const float rstd = use_rsqrt ? rsqrtf(var + eps) : 1.0f / sqrtf(var + eps);
  float sa = 0.f, sax = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      xv[i] = (xv[i] - mean) * rstd;                  // xh
      av[i] = gamma[c] * dr[c];
      sa += av[i];
      sax = fmaf(av[i], xv[i], sax);
    } else {
      xv[i] = av[i] = 0.f;
    }
  }
  const float ma = warp_sum(sa) / (float)C, max_ = warp_sum(sax) / (float)C;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    if (c < C) {
      float v = rstd * (av[i] - ma - xv[i] * max_);
      if (add) v += add[(size_t)row * C + c];
      dx[(size_t)row * C + c] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// LinearAttention core backward (jf.py:206-225).  Per (image, head), d = e = 32:
//   qs = softmax_d(q) * scale ; ks[n][d] = exp(k[n][d] - kmax[d]) / ksum[d] ; vs = v * vscale
//   ctx[d][e] = sum_n ks[n][d] vs[n][e] ; out[n][e] = sum_d ctx[d][e] qs[n][d]
// backward:
//   dctx[d][e] = sum_n qs[n][d] dout[n][e]                                    (kernel 1, reduction over pixels)
//   sk[d]      = sum_n dks[n][d] ks[n][d] = sum_e dctx[d][e] ctx[d][e]         (no pass over pixels needed)
//   dqs[n][d]  = sum_e ctx[d][e] dout[n][e] ;  dq = scale * s * (dqs - sum_d dqs s),  s = softmax_d(q)
//   dk[n][d]   = ks[n][d] * (sum_e dctx[d][e] vs[n][e] - sk[d])
//   dv[n][e]   = vscale * sum_d ks[n][d] dctx[d][e]                           (kernel 2, per pixel)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
linattn_bwd_dctx_kernel(const float* __restrict__ qkv, const float* __restrict__ dout, float* __restrict__ dctx, int HW,
                        int heads, float scale) {
  constexpr int TP = 64;
  __shared__ float s_q[TP][DH + 1];
  __shared__ __align__(16) float s_do[TP][DH];
  const int head = blockIdx.x % heads;
  const int64_t img = blockIdx.x / heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const float* qbase = qkv + (size_t)img * HW * C3 + head * DH;
  const float* dbase = dout + (size_t)img * HW * hid + head * DH;
  const int prow = tid >> 3, pc = (tid & 7) * 4;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  for (int t0 = 0; t0 < HW; t0 += TP) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = prow + 16 * i, n = t0 + r;
      float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < HW) {
        q4 = __ldg(reinterpret_cast<const float4*>(qbase + (size_t)n * C3 + pc));
        d4 = __ldg(reinterpret_cast<const float4*>(dbase + (size_t)n * hid + pc));
      }
      s_q[r][pc] = q4.x; s_q[r][pc + 1] = q4.y; s_q[r][pc + 2] = q4.z; s_q[r][pc + 3] = q4.w;
      *reinterpret_cast<float4*>(&s_do[r][pc]) = d4;     // zero rows beyond HW contribute nothing
    }
    __syncthreads();
    if (tid < TP) {                                       // softmax over d of row tid, times scale, in place
      float m = -INFINITY;
#pragma unroll
      for (int d = 0; d < DH; ++d) m = fmaxf(m, s_q[tid][d]);
      float sum = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) {
        const float e = expf(s_q[tid][d] - m);
        s_q[tid][d] = e;
        sum += e;
      }
      const float inv = scale / sum;
#pragma unroll
      for (int d = 0; d < DH; ++d) s_q[tid][d] *= inv;
    }
    __syncthreads();
#pragma unroll 8
    for (int n = 0; n < TP; ++n) {
      const float qv = s_q[n][lane];
      const float4 v0 = *reinterpret_cast<const float4*>(&s_do[n][warp * 8]);
      const float4 v1 = *reinterpret_cast<const float4*>(&s_do[n][warp * 8 + 4]);
      acc[0] = fmaf(qv, v0.x, acc[0]); acc[1] = fmaf(qv, v0.y, acc[1]);
      acc[2] = fmaf(qv, v0.z, acc[2]); acc[3] = fmaf(qv, v0.w, acc[3]);
      acc[4] = fmaf(qv, v1.x, acc[4]); acc[5] = fmaf(qv, v1.y, acc[5]);
      acc[6] = fmaf(qv, v1.z, acc[6]); acc[7] = fmaf(qv, v1.w, acc[7]);
    }
  }
  float* dst = dctx + (size_t)blockIdx.x * DH * DH + lane * DH + warp * 8;   // [d][e]
#pragma unroll
  for (int e = 0; e < 8; ++e) dst[e] = acc[e];
}

// kernel 2: 128 pixels of one (image, head) per CTA, thread per pixel; rows staged through shared memory (coalesced).
constexpr int LB_PIX = 128;
constexpr int LB_PITCH = DH + 1;
constexpr size_t LB_SMEM = (size_t)(2 * DH * DH + DH + 3 * LB_PIX * LB_PITCH) * sizeof(float);

__global__ void __launch_bounds__(LB_PIX)
linattn_bwd_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ ctx, const float* __restrict__ kstat,
                         const float* __restrict__ dctx, const float* __restrict__ dout, float* __restrict__ dqkv, int HW,
                         int heads, float scale, float vscale) {
  extern __shared__ __align__(16) float smf[];
  float* s_ctx = smf;                       // [d][e]
  float* s_dctx = smf + DH * DH;            // [d][e]
  float* s_sk = smf + 2 * DH * DH;          // [d]
  float* bufA = s_sk + DH;                  // [128][33] staging, later dqs
  float* bufB = bufA + LB_PIX * LB_PITCH;   // softmax_d(q)
  float* bufC = bufB + LB_PIX * LB_PITCH;   // ks
  const int head = blockIdx.y;
  const int64_t img = blockIdx.z;
  const int tid = threadIdx.x;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const int n0 = blockIdx.x * LB_PIX;
  const size_t ch = ((size_t)img * heads + head) * DH * DH;
  for (int i = tid; i < DH * DH; i += LB_PIX) {
    s_ctx[i] = ctx[ch + i];
    s_dctx[i] = dctx[ch + i];
  }
  __syncthreads();
  if (tid < DH) {
    float t = 0.f;
    for (int e = 0; e < DH; ++e) t = fmaf(s_dctx[tid * DH + e], s_ctx[tid * DH + e], t);
    s_sk[tid] = t;
  }
  const float* kst = kstat + ((size_t)img * heads + head) * DH * 2;   // [d][2] = (max, sum)
  // coalesced staging of a [128][32] slab: 8 lanes cover one pixel's 128-byte head slice
  auto stage = [&](const float* base, int pitch_floats, float* dst) {
    __syncthreads();
    for (int i = tid; i < LB_PIX * 8; i += LB_PIX) {
      const int r = i >> 3, c = (i & 7) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < HW) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)(n0 + r) * pitch_floats + c));
      float* d = dst + r * LB_PITCH + c;
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    __syncthreads();
  };
  const float* qbase = qkv + (size_t)img * HW * C3 + head * DH;
  float dov[DH], vv[DH];
  stage(dout + (size_t)img * HW * hid + head * DH, hid, bufA);
#pragma unroll
  for (int e = 0; e < DH; ++e) dov[e] = bufA[tid * LB_PITCH + e];
  stage(qbase + 2 * hid, C3, bufA);
#pragma unroll
  for (int e = 0; e < DH; ++e) vv[e] = bufA[tid * LB_PITCH + e] * vscale;
  stage(qbase, C3, bufB);
  stage(qbase + hid, C3, bufC);
  {  // own row: s = softmax_d(q) in bufB, ks in bufC
    float* qr = bufB + tid * LB_PITCH;
    float m = -INFINITY;
#pragma unroll
    for (int d = 0; d < DH; ++d) m = fmaxf(m, qr[d]);
    float sum = 0.f;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      const float e = expf(qr[d] - m);
      qr[d] = e;
      sum += e;
    }
    const float inv = 1.0f / sum;
    float* kr = bufC + tid * LB_PITCH;
#pragma unroll
    for (int d = 0; d < DH; ++d) {
      qr[d] *= inv;
      kr[d] = expf(kr[d] - kst[2 * d]) / kst[2 * d + 1];
    }
  }
  __syncthreads();   // s_sk visible; bufA free (every thread has copied its v row)
  float dv[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) dv[e] = 0.f;
  float dot = 0.f;
  float* dqs = bufA + tid * LB_PITCH;
  float* sr = bufB + tid * LB_PITCH;
  float* kr = bufC + tid * LB_PITCH;
#pragma unroll 1
  for (int d = 0; d < DH; ++d) {
    float a = 0.f, b = 0.f;
    const float ksd = kr[d];
#pragma unroll
    for (int e = 0; e < DH; e += 4) {      // broadcast LDS.128: 2 shared loads per 12 FMAs
      const float4 c = *reinterpret_cast<const float4*>(&s_ctx[d * DH + e]);
      const float4 dc = *reinterpret_cast<const float4*>(&s_dctx[d * DH + e]);
      a = fmaf(c.x, dov[e], fmaf(c.y, dov[e + 1], fmaf(c.z, dov[e + 2], fmaf(c.w, dov[e + 3], a))));       // dqs[d]
      b = fmaf(dc.x, vv[e], fmaf(dc.y, vv[e + 1], fmaf(dc.z, vv[e + 2], fmaf(dc.w, vv[e + 3], b))));       // dks[d]
      dv[e] = fmaf(ksd, dc.x, dv[e]);
      dv[e + 1] = fmaf(ksd, dc.y, dv[e + 1]);
      dv[e + 2] = fmaf(ksd, dc.z, dv[e + 2]);
      dv[e + 3] = fmaf(ksd, dc.w, dv[e + 3]);
    }
    dqs[d] = a;
    dot = fmaf(a, sr[d], dot);
    kr[d] = ksd * (b - s_sk[d]);         // dk[d]
  }
#pragma unroll
  for (int d = 0; d < DH; ++d) sr[d] = scale * sr[d] * (dqs[d] - dot);   // dq[d]
#pragma unroll
  for (int e = 0; e < DH; ++e) dqs[e] = dv[e] * vscale;                   // dv[e]
  __syncthreads();
  float* obase = dqkv + (size_t)img * HW * C3 + head * DH;
  for (int i = tid; i < LB_PIX * 8; i += LB_PIX) {
    const int r = i >> 3, c = (i & 7) * 4;
    if (n0 + r < HW) {
      float* o = obase + (size_t)(n0 + r) * C3 + c;
      const float* b0 = bufB + r * LB_PITCH + c;
      const float* b1 = bufC + r * LB_PITCH + c;
      const float* b2 = bufA + r * LB_PITCH + c;
      *reinterpret_cast<float4*>(o) = make_float4(b0[0], b0[1], b0[2], b0[3]);
      *reinterpret_cast<float4*>(o + hid) = make_float4(b1[0], b1[1], b1[2], b1[3]);
      *reinterpret_cast<float4*>(o + 2 * hid) = make_float4(b2[0], b2[1], b2[2], b2[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// softmax Attention core backward (jf.py:241-255), T = HW tokens per image, one CTA per (image, head):
//   S = scale q k^T ; P = softmax_j(S) ; O = P v
//   D_i = dO_i . O_i ; dS = P (dO v^T - D) ; dq = scale dS k ; dk = scale dS^T q ; dv = P^T dO
// pass 1: thread per query (K, V in shared memory); pass 2: thread per key (scaled Q, dO in shared memory).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ out, const float* __restrict__ dout,
                     float* __restrict__ dqkv, int T, int heads, float scale) {
  extern __shared__ __align__(16) float sma[];
  float* s_x = sma;                 // [T][32]: K, then scaled Q
  float* s_y = sma + (size_t)T * DH;  // [T][32]: V, then dO
  float* s_m = s_y + (size_t)T * DH;  // [T] row max
  float* s_l = s_m + T;             // [T] row sum
  float* s_D = s_l + T;             // [T]
  const int head = blockIdx.x % heads;
  const int64_t img = blockIdx.x / heads;
  const int tid = threadIdx.x;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const float* qb = qkv + (size_t)img * T * C3 + head * DH;
  const float* ob = out + (size_t)img * T * hid + head * DH;
  const float* db = dout + (size_t)img * T * hid + head * DH;
  float* gq = dqkv + (size_t)img * T * C3 + head * DH;
  for (int i = tid; i < T * 8; i += 128) {
    const int r = i >> 3, c = (i & 7) * 4;
    *reinterpret_cast<float4*>(s_x + r * DH + c) = __ldg(reinterpret_cast<const float4*>(qb + (size_t)r * C3 + hid + c));
    *reinterpret_cast<float4*>(s_y + r * DH + c) = __ldg(reinterpret_cast<const float4*>(qb + (size_t)r * C3 + 2 * hid + c));
  }
  __syncthreads();
  for (int i = tid; i < T; i += 128) {
    float q[DH], dO[DH], dq[DH];
    float D = 0.f;
#pragma unroll
    for (int e = 0; e < DH; e += 4) {
      const float4 q4 = __ldg(reinterpret_cast<const float4*>(qb + (size_t)i * C3 + e));
      const float4 d4 = __ldg(reinterpret_cast<const float4*>(db + (size_t)i * hid + e));
      const float4 o4 = __ldg(reinterpret_cast<const float4*>(ob + (size_t)i * hid + e));
      q[e] = q4.x * scale; q[e + 1] = q4.y * scale; q[e + 2] = q4.z * scale; q[e + 3] = q4.w * scale;
      dO[e] = d4.x; dO[e + 1] = d4.y; dO[e + 2] = d4.z; dO[e + 3] = d4.w;
      D = fmaf(d4.x, o4.x, fmaf(d4.y, o4.y, fmaf(d4.z, o4.z, fmaf(d4.w, o4.w, D))));
      dq[e] = dq[e + 1] = dq[e + 2] = dq[e + 3] = 0.f;
    }
    float m = -INFINITY;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < DH; ++e) s = fmaf(q[e], s_x[j * DH + e], s);
      m = fmaxf(m, s);
    }
    float l = 0.f;
    for (int j = 0; j < T; ++j) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < DH; ++e) s = fmaf(q[e], s_x[j * DH + e], s);
      l += expf(s - m);
    }
    const float inv_l = 1.0f / l;
    for (int j = 0; j < T; ++j) {
      float s = 0.f, dP = 0.f;
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        s = fmaf(q[e], s_x[j * DH + e], s);
        dP = fmaf(dO[e], s_y[j * DH + e], dP);
      }
      const float ds = expf(s - m) * inv_l * (dP - D);
#pragma unroll
      for (int e = 0; e < DH; ++e) dq[e] = fmaf(ds, s_x[j * DH + e], dq[e]);
    }
#pragma unroll
    for (int e = 0; e < DH; e += 4)
      *reinterpret_cast<float4*>(gq + (size_t)i * C3 + e) =
          make_float4(dq[e] * scale, dq[e + 1] * scale, dq[e + 2] * scale, dq[e + 3] * scale);
    s_m[i] = m;
    s_l[i] = inv_l;
    s_D[i] = D;
  }
  __syncthreads();
  for (int i = tid; i < T * 8; i += 128) {
    const int r = i >> 3, c = (i & 7) * 4;
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(qb + (size_t)r * C3 + c));
    *reinterpret_cast<float4*>(s_x + r * DH + c) = make_float4(q4.x * scale, q4.y * scale, q4.z * scale, q4.w * scale);
    *reinterpret_cast<float4*>(s_y + r * DH + c) = __ldg(reinterpret_cast<const float4*>(db + (size_t)r * hid + c));
  }
  __syncthreads();
  for (int j = tid; j < T; j += 128) {
    float k[DH], v[DH], dk[DH], dv[DH];
#pragma unroll
    for (int e = 0; e < DH; e += 4) {
      const float4 k4 = __ldg(reinterpret_cast<const float4*>(qb + (size_t)j * C3 + hid + e));
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(qb + (size_t)j * C3 + 2 * hid + e));
      k[e] = k4.x; k[e + 1] = k4.y; k[e + 2] = k4.z; k[e + 3] = k4.w;
      v[e] = v4.x; v[e + 1] = v4.y; v[e + 2] = v4.z; v[e + 3] = v4.w;
      dk[e] = dk[e + 1] = dk[e + 2] = dk[e + 3] = 0.f;
      dv[e] = dv[e + 1] = dv[e + 2] = dv[e + 3] = 0.f;
    }
    for (int i = 0; i < T; ++i) {
      float s = 0.f, dP = 0.f;
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        s = fmaf(s_x[i * DH + e], k[e], s);
        dP = fmaf(s_y[i * DH + e], v[e], dP);
      }
      const float pij = expf(s - s_m[i]) * s_l[i];
      const float ds = pij * (dP - s_D[i]);
#pragma unroll
      for (int e = 0; e < DH; ++e) {
        dk[e] = fmaf(ds, s_x[i * DH + e], dk[e]);
        dv[e] = fmaf(pij, s_y[i * DH + e], dv[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < DH; e += 4) {
      *reinterpret_cast<float4*>(gq + (size_t)j * C3 + hid + e) = make_float4(dk[e], dk[e + 1], dk[e + 2], dk[e + 3]);
      *reinterpret_cast<float4*>(gq + (size_t)j * C3 + 2 * hid + e) = make_float4(dv[e], dv[e + 1], dv[e + 2], dv[e + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// small elementwise / reduction helpers
// ------------------------------------------------------------------------------------------------------------------
__global__ void add_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* out, int64_t n4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
  }
}

// backward of nn.Upsample(scale 2, nearest): dx[n,h,w,:] = sum of the 2x2 block of dy [N,2H,2W,C]
__global__ void sumpool2x2_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, int64_t N, int H, int W, int C4) {
  const int64_t total = N * H * W * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    int64_t t = i / C4;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int64_t n = t / H;
    const size_t base = (((size_t)n * 2 * H + 2 * h) * 2 * W + 2 * w) * C4 + c;
    const float4 a = dy[base], b = dy[base + C4], d = dy[base + (size_t)2 * W * C4], e = dy[base + (size_t)2 * W * C4 + C4];
    dx[i] = make_float4(a.x + b.x + d.x + e.x, a.y + b.y + d.y + e.y, a.z + b.z + d.z + e.z, a.w + b.w + d.w + e.w);
  }
}

// ForceUnet head (jf.py:478-479): out[n][o] = bias[o] + sum_c W[o][c] * mean_p x[n][p][c]; one CTA per image
__global__ void __launch_bounds__(256)
mean_head_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ out,
                 int HW, int C, int O) {
  extern __shared__ float s_mean[];   // [C]
  const int64_t n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += 256) {
    double acc = 0.0;
    for (int p = 0; p < HW; ++p) acc += (double)x[((size_t)n * HW + p) * C + c];
    s_mean[c] = (float)(acc / (double)HW);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int o = warp; o < O; o += 8) {
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) acc = fmaf(W[(size_t)o * C + c], s_mean[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[n * O + o] = acc + (bias ? bias[o] : 0.f);
  }
}

// dx[n][p][c] = sum_o dout[n][o] W[o][c] / HW
__global__ void mean_head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ W, float* __restrict__ dx,
                                     int64_t N, int HW, int C, int O) {
  const int64_t total = N * HW * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t n = i / ((int64_t)HW * C);
    float acc = 0.f;
    for (int o = 0; o < O; ++o) acc = fmaf(dout[n * O + o], W[(size_t)o * C + c], acc);
    dx[i] = acc / (float)HW;
  }
}

// SinusoidalPosEmb(float time) -> Linear -> GELU (jf.py:137-149, :313-318); hidden[b][j]
__global__ void sinusoidal_linear_gelu_f32_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                                                  const float* __restrict__ w1, const float* __restrict__ b1,
                                                  float* __restrict__ hidden, int B, int dim, int tdim) {
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= (int64_t)B * tdim) return;
  const int b = (int)(wg / tdim), j = (int)(wg % tdim);
  const int half = dim / 2;
  const float tv = t[b];
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

__global__ void linear_rows2_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                    float* __restrict__ out, int B, int K, int N) {
  const int64_t wg = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wg >= (int64_t)B * N) return;
  const int b = (int)(wg / N), j = (int)(wg % N);
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(W[(size_t)j * K + k], x[(size_t)b * K + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[(size_t)b * N + j] = acc + (bias ? bias[j] : 0.f);
}

// Backward of the whole time path for one sample per CTA (tdim threads, tdim <= 1024):
//   emb = [sin(t f) | cos(t f)] ; hp = W1 emb + b1 ; h = gelu(hp) ; te = W2 h + b2 ; ss = Wp silu(te) + bp
//   given dss [total]: dt = d ss / d t
__global__ void time_mlp_bwd_kernel(const float* __restrict__ t, const float* __restrict__ freqs, const float* __restrict__ w1,
                                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ wp,
                                    const float* __restrict__ t_emb, const float* __restrict__ dss, float* __restrict__ dt,
                                    int dim, int tdim, int total) {
  extern __shared__ float smt[];     // [dim] emb, [tdim] a, [tdim] b
  float* s_emb = smt;
  float* s_a = smt + dim;
  float* s_b = s_a + tdim;
  const int n = blockIdx.x, k = threadIdx.x;
  const int half = dim / 2;
  const float tv = t[n];
  if (k < dim) {
    const float arg = __fmul_rn(tv, freqs[k < half ? k : k - half]);
    s_emb[k] = (k < half) ? sinf(arg) : cosf(arg);
  }
  // d silu(te)[k] = sum_r Wp[r][k] dss[r]   (coalesced over k)
  float acc = 0.f;
  const float* dr = dss + (size_t)n * total;
  for (int r = 0; r < total; ++r) acc = fmaf(wp[(size_t)r * tdim + k], dr[r], acc);
  const float te = t_emb[(size_t)n * tdim + k];
  const float sg = sigmoid_f(te);
  s_a[k] = acc * sg * (1.0f + te * (1.0f - sg));            // d te[k]
  __syncthreads();
  // d h[j] = sum_k W2[k][j] d te[k] ; hp[j] recomputed
  float dh = 0.f;
  for (int kk = 0; kk < tdim; ++kk) dh = fmaf(w2[(size_t)kk * tdim + k], s_a[kk], dh);
  float hp = b1[k];
  for (int i = 0; i < dim; ++i) hp = fmaf(w1[(size_t)k * dim + i], s_emb[i], hp);
  const float cdf = 0.5f * (1.0f + erff(hp * 0.70710678118654752440f));
  const float pdf = 0.39894228040143267794f * expf(-0.5f * hp * hp);
  s_b[k] = dh * (cdf + hp * pdf);                           // d hp[j]
  __syncthreads();
  float contrib = 0.f;
  if (k < dim) {
    float de = 0.f;
    for (int j = 0; j < tdim; ++j) de = fmaf(w1[(size_t)j * dim + k], s_b[j], de);
    const float f = freqs[k < half ? k : k - half];
    // d sin(t f)/dt = f cos(t f) ; d cos(t f)/dt = -f sin(t f)
    const float other = (k < half) ? s_emb[k + half] : s_emb[k - half];
    contrib = (k < half) ? de * f * other : -de * f * other;
  }
  __syncthreads();
  s_a[k] = contrib;
  __syncthreads();
  if (k == 0) {
    double tot = 0.0;
    for (int i = 0; i < dim; ++i) tot += (double)s_a[i];
    dt[n] = (float)tot;
  }
}

}  // namespace n2d
}  // namespace dpc

using namespace dpc;
using namespace dpc::n2d;

extern "C" int dpc_gn_silu_bwd(const float* y, const double* stats, const float* gamma, const float* beta, const float* scale_shift,
                               int64_t ss_stride, int64_t ss_off, const float* dout, float* dy, double* sums_ws, float* dss,
                               int32_t B, int64_t rows_per_sample, int32_t C, int32_t groups, float eps, void* stream) {
  DPC_CHECK_ARG(y && stats && gamma && beta && dout && dy && sums_ws && B > 0 && B <= 65535 && rows_per_sample > 0);
  DPC_CHECK_ARG(C % 4 == 0 && C >= 4 && C <= 1024 && 256 % (C / 4) == 0 && groups > 0 && groups <= 32 && C % groups == 0);
  DPC_CHECK_ARG(dss == nullptr || scale_shift != nullptr);
  cudaStream_t st = (cudaStream_t)stream;
  DPC_CUDA(cudaMemsetAsync(sums_ws, 0, (size_t)B * C * 2 * sizeof(double), st));
  GnArgs a{y, stats, gamma, beta, scale_shift, ss_stride, ss_off, dout, dy, sums_ws, dss, rows_per_sample, C, groups, eps, 0};
  const int rstep = 256 / (C / 4);
  int64_t ctas = (148 * 8 + B - 1) / B;                    // ~8 CTAs per SM over the whole batch
  int64_t rpc = (rows_per_sample + ctas - 1) / ctas;
  if (rpc < 4 * rstep) rpc = 4 * rstep;
  rpc = (rpc + rstep - 1) / rstep * rstep;
  a.rows_per_cta = rpc;
  dim3 grid((unsigned)((rows_per_sample + rpc - 1) / rpc), (unsigned)B);
  gn_silu_bwd_reduce_kernel<<<grid, 256, 0, st>>>(a);
  DPC_LAUNCH_CHECK();
  gn_silu_bwd_apply_kernel<<<grid, 256, (size_t)(2 * C + 2 * groups) * sizeof(float), st>>>(a);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_layernorm_channels_bwd(const float* x, const float* gamma, const float* dy, const float* add, float* dx,
                                          int64_t rows, int32_t C, float eps, int32_t use_rsqrt, void* stream) {
  DPC_CHECK_ARG(x && gamma && dy && dx && rows > 0 && C > 0 && C <= 1024);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((rows * 32 + 255) / 256);
  if (C <= 64) layernorm_bwd_kernel<2><<<blocks, 256, 0, st>>>(x, gamma, dy, add, dx, rows, C, eps, use_rsqrt);
  else if (C <= 128) layernorm_bwd_kernel<4><<<blocks, 256, 0, st>>>(x, gamma, dy, add, dx, rows, C, eps, use_rsqrt);
  else if (C <= 256) layernorm_bwd_kernel<8><<<blocks, 256, 0, st>>>(x, gamma, dy, add, dx, rows, C, eps, use_rsqrt);
  else if (C <= 512) layernorm_bwd_kernel<16><<<blocks, 256, 0, st>>>(x, gamma, dy, add, dx, rows, C, eps, use_rsqrt);
  else layernorm_bwd_kernel<32><<<blocks, 256, 0, st>>>(x, gamma, dy, add, dx, rows, C, eps, use_rsqrt);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_linattn2d_bwd(const float* qkv, const float* ctx, const float* kstat, const float* dout, float* dctx_ws,
                                 float* dqkv, int32_t BF, int32_t HW, int32_t heads, float scale, float v_scale, void* stream) {
  DPC_CHECK_ARG(qkv && ctx && kstat && dout && dctx_ws && dqkv && BF > 0 && BF <= 65535 && HW > 0 && heads > 0 && heads <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[device_ordinal()];
  if (!configured) {
    DPC_CUDA(cudaFuncSetAttribute(linattn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LB_SMEM));
    configured = true;
  }
  linattn_bwd_dctx_kernel<<<(unsigned)((int64_t)BF * heads), 128, 0, st>>>(qkv, dout, dctx_ws, HW, heads, scale);
  DPC_LAUNCH_CHECK();
  dim3 grid((unsigned)((HW + LB_PIX - 1) / LB_PIX), (unsigned)heads, (unsigned)BF);
  linattn_bwd_apply_kernel<<<grid, LB_PIX, LB_SMEM, st>>>(qkv, ctx, kstat, dctx_ws, dout, dqkv, HW, heads, scale, v_scale);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_attention2d_bwd(const float* qkv, const float* out, const float* dout, float* dqkv, int32_t BF, int32_t HW,
                                   int32_t heads, float scale, void* stream) {
  DPC_CHECK_ARG(qkv && out && dout && dqkv && BF > 0 && HW > 0 && heads > 0);
  const size_t smem = ((size_t)2 * HW * DH + 3 * (size_t)HW) * sizeof(float);
  if (smem > 200 * 1024) return -2;                         // > ~780 tokens per image: not served
  static size_t configured_[kMaxDevices] = {};
  size_t& configured = configured_[device_ordinal()];
  if (smem > configured) {
    DPC_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  attention_bwd_kernel<<<(unsigned)((int64_t)BF * heads), 128, smem, (cudaStream_t)stream>>>(qkv, out, dout, dqkv, HW, heads, scale);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  DPC_CHECK_ARG(a && b && out && n > 0 && n % 4 == 0);
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  add_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                                reinterpret_cast<float4*>(out), n / 4);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_sumpool2x2(const float* dy, float* dx, int64_t N, int32_t H, int32_t W, int32_t C, void* stream) {
  DPC_CHECK_ARG(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0);
  int64_t blocks = (N * H * W * (C / 4) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  sumpool2x2_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(dy),
                                                                       reinterpret_cast<float4*>(dx), N, H, W, C / 4);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_mean_head(const float* x, const float* W, const float* bias, float* out, int64_t N, int32_t HW, int32_t C,
                             int32_t O, void* stream) {
  DPC_CHECK_ARG(x && W && out && N > 0 && HW > 0 && C > 0 && C <= 4096 && O > 0);
  mean_head_kernel<<<(unsigned)N, 256, (size_t)C * sizeof(float), (cudaStream_t)stream>>>(x, W, bias, out, HW, C, O);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_mean_head_bwd(const float* dout, const float* W, float* dx, int64_t N, int32_t HW, int32_t C, int32_t O,
                                 void* stream) {
  DPC_CHECK_ARG(dout && W && dx && N > 0 && HW > 0 && C > 0 && O > 0);
  int64_t blocks = (N * HW * C + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mean_head_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dout, W, dx, N, HW, C, O);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_time_embed_f32(const float* t, const float* freqs, const float* w1, const float* b1, const float* w2,
                                  const float* b2, float* hidden_ws, float* t_emb, int32_t B, int32_t dim, void* stream) {
  DPC_CHECK_ARG(t && freqs && w1 && b1 && w2 && b2 && hidden_ws && t_emb && B > 0 && dim > 0 && dim % 2 == 0);
  const int tdim = dim * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t warps = (int64_t)B * tdim;
  const unsigned blocks = (unsigned)((warps * 32 + 255) / 256);
  sinusoidal_linear_gelu_f32_kernel<<<blocks, 256, 0, st>>>(t, freqs, w1, b1, hidden_ws, B, dim, tdim);
  DPC_LAUNCH_CHECK();
  linear_rows2_kernel<<<blocks, 256, 0, st>>>(hidden_ws, w2, b2, t_emb, B, tdim, tdim);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_time_mlp_bwd(const float* t, const float* freqs, const float* w1, const float* b1, const float* w2,
                                const float* w_proj, const float* t_emb, const float* dss, float* dt, int32_t B, int32_t dim,
                                int32_t total, void* stream) {
  DPC_CHECK_ARG(t && freqs && w1 && b1 && w2 && w_proj && t_emb && dss && dt && B > 0 && dim > 0 && dim % 2 == 0 && total > 0);
  const int tdim = dim * 4;
  DPC_CHECK_ARG(tdim <= 1024 && tdim % 32 == 0);
  time_mlp_bwd_kernel<<<(unsigned)B, tdim, (size_t)(dim + 2 * tdim) * sizeof(float), (cudaStream_t)stream>>>(
      t, freqs, w1, b1, w2, w_proj, t_emb, dss, dt, dim, tdim, total);
  DPC_LAUNCH_CHECK();
  return 0;
}
