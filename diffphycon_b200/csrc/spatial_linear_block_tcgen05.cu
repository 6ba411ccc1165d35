// Fused SpatialLinearAttention block of Unet3D_with_Conv3D at dim 64 (conv3d.py:165-174 LayerNorm via PreNorm, :232-257
// SpatialLinearAttention, :153-157 Residual), per frame of HW pixels:
//     q, k, v = to_qkv(LN x);  q = softmax_d(q) * scale;  k = softmax_n(k);  ctx = k v^T;  y = x + to_out(ctx^T q) + b
// The 384-wide qkv tensor (12.9 GB at the metric shape) and the 128-wide attention output never reach HBM.  Three launches:
//
//  1. linattn_context_tc_kernel   one frame at a time per persistent CTA, 128-pixel tiles:
//       TMA x tile -> LayerNorm in place (thread = pixel; the gain is folded into the weights on the host)
//       MMA  kT[128 (h,d) x 128 px] = Wk . xhat^T,  vT = Wv . xhat^T   (the TRANSPOSED projections: a TMEM lane is one
//            (head, d) row, so the softmax over the pixels is a per-thread loop, no shuffles)
//       rows online softmax over n (running max m, rescale alpha = exp(m - m'), partial sums Z): p = exp(k - m') written
//            back to TMEM in place; the v rows go to smem as a K-major B operand
//       MMA  ctx_tile[128 (h,d) x 128 (h',e)] = P . V^T  with the A operand read straight from TMEM (only the h == h'
//            blocks are used; the tile GEMM is small)
//       rows ctx = ctx * alpha + ctx_tile (own head block, 16 columns per thread, in registers)
//       frame end: ctx * scale / Z -> global [BF][4][32][32]
//  2. linattn_fold_out_kernel     MT[f][c][(h,d)] = sum_e Wout[c][h*32+e] ctx[f][h][d][e]: the out-projection is folded
//       into the context, so the apply pass is a single K = 128 GEMM per tile
//  3. linattn_apply_tc_kernel     128-pixel tiles (persistent):
//       TMA x tile + MT of its frame -> LayerNorm in place, raw x parked in TMEM
//       MMA  q[128 px x 128 (h,d)] = xhat . Wq^T
//       rows softmax over d per head (thread = pixel) -> q~ written back to TMEM in place
//       MMA  y[128 px x 64] = q~ (TMEM A operand) . MT
//       rows y + bias + raw x -> swizzled smem -> TMA store
// HBM traffic: x twice, y once (6.4 GB at the metric shape instead of ~45 GB for the unfused sequence).
// Warp roles (384 threads) as in temporal_block_tcgen05.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator,
// warps 4-11 = 256 row threads in two groups (warp%4 = TMEM lane quarter, group = channel chunk / column half).
#include "tc_common.cuh"

namespace dpc {
namespace sl {

using namespace dpc::tc;

constexpr int C = 64, HID = 128, HEADS = 4, DH = 32, TP = 128;
constexpr int THREADS = 384;
constexpr float ATT_SCALE = 0.17677669529663687f;       // 32^-0.5 (conv3d.py:236)
constexpr uint32_t XA_BYTES = 32768;                     // 2 channel chunks x [128 rows x 128 B]
constexpr uint32_t IDESC_BASE = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);   // f32 accum, tf32 x tf32, M = 128
constexpr uint32_t IDESC_N128 = IDESC_BASE | ((uint32_t)(128 >> 3) << 17);
constexpr uint32_t IDESC_N64 = IDESC_BASE | ((uint32_t)(64 >> 3) << 17);

// LayerNorm of tile row r over its 64 channels, in place in the 128B-swizzled x tile (this thread: channel chunk gi; the two
// groups exchange their partial moments through `exch`).  PARK: keep the raw row in TMEM for the residual.
template <bool PARK>
__device__ __forceinline__ void layernorm_row(uint32_t xa, int gi, int r, uint32_t sw, float2* exch, float eps, uint32_t tpark) {
  float x[32];
  const uint32_t xrow = xa + gi * 16384 + r * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = lds128(xrow + ((j ^ sw) << 4));
    x[j * 4 + 0] = v.x; x[j * 4 + 1] = v.y; x[j * 4 + 2] = v.z; x[j * 4 + 3] = v.w;
  }
  const float x0 = lds32(xa + r * 128 + (sw << 4));
  if (PARK) {
    uint32_t u[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(x[j]);
    tmem_st32(tpark, u);
  }
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) { x[j] -= x0; s1 += x[j]; s2 = fmaf(x[j], x[j], s2); }
  exch[gi * 128 + r] = make_float2(s1, s2);
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const float2 o = exch[(gi ^ 1) * 128 + r];
  const float t1 = gi ? o.x + s1 : s1 + o.x, t2 = gi ? o.y + s2 : s2 + o.y;
  const float dm = t1 * (1.0f / 64.0f);
  const float var = fmaxf(t2 * (1.0f / 64.0f) - dm * dm, 0.f);
  const float rstd = 1.0f / sqrtf(var + eps);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    sts128(xrow + ((j ^ sw) << 4), (x[4 * j] - dm) * rstd, (x[4 * j + 1] - dm) * rstd, (x[4 * j + 2] - dm) * rstd,
           (x[4 * j + 3] - dm) * rstd);
  if (PARK) tmem_wait_st();
}

// ------------------------------------------------------------------------------------------------------------------------
// 1. context pass
// ------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t A_OFF_W = 0;                           // 2 chunks x [256 rows x 128 B]: rows 0..127 to_k, 128..255 to_v
constexpr uint32_t A_OFF_XA = 65536;                      // 2 buffers x XA_BYTES
constexpr uint32_t A_OFF_VT = A_OFF_XA + 2 * XA_BYTES;    // 4 pixel chunks x [128 (h,e) rows x 128 B]
constexpr uint32_t A_OFF_EXLN = A_OFF_VT + 65536;         // float2 [2][128]
constexpr uint32_t A_OFF_EXMAX = A_OFF_EXLN + 2048;       // float [2][128]
constexpr uint32_t A_OFF_EXZ = A_OFF_EXMAX + 1024;        // float [2][128]
constexpr uint32_t A_OFF_BAR = A_OFF_EXZ + 1024;
constexpr uint32_t A_SMEM = A_OFF_BAR + 128 + 1024;

struct CtxParams {
  float* ctx;       // [BF][4][32][32], scale / Z applied
  float eps;
  int BF, HW;
};

__global__ void __launch_bounds__(THREADS, 1)
linattn_context_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const CtxParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + A_OFF_BAR;
  const uint32_t w_full = bars, x_full = bars + 8, x_empty = bars + 24, a_ready = bars + 40, kv_full = bars + 48;
  const uint32_t p_ready = bars + 56, ctx_full = bars + 64, tmem_slot = bars + 72;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpf = p.HW / TP;
  const int nfr = ((int)blockIdx.x < p.BF) ? (p.BF - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int ntl = nfr * tpf;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(x_full + 8 * i, 1); mbar_init(x_empty + 8 * i, 1); }
    mbar_init(a_ready, 256);
    mbar_init(kv_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(ctx_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0 && lane == 0) {
    // ------------------------------------------- TMA producer -------------------------------------------
    if (ntl > 0) {
      mbar_expect_tx(w_full, 65536);
      for (int c = 0; c < 2; ++c) tma_load_2d(base + A_OFF_W + c * 32768, &tmW, w_full, c * 32, HID);
    }
    for (int it = 0; it < ntl; ++it) {
      const int fi = it / tpf, t = it - fi * tpf;
      const int row0 = ((int)blockIdx.x + fi * (int)gridDim.x) * p.HW + t * TP;
      const int buf = it & 1;
      mbar_wait(x_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
      mbar_expect_tx(x_full + 8 * buf, XA_BYTES);
      const uint32_t dst = base + A_OFF_XA + buf * XA_BYTES;
      for (int c = 0; c < 2; ++c) tma_load_2d(dst + c * 16384, &tmX, x_full + 8 * buf, c * 32, row0);
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    const uint64_t w_desc = umma_desc(base + A_OFF_W), vt_desc = umma_desc(base + A_OFF_VT);
    if (ntl > 0) mbar_wait(w_full, 0);
    for (int it = 0; it < ntl; ++it) {
      const uint64_t xa_desc = umma_desc(base + A_OFF_XA + (it & 1) * XA_BYTES);
      mbar_wait(a_ready, it & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a = w_desc + (uint64_t)(c * (32768 >> 4) + 2 * k);
            const uint64_t b = xa_desc + (uint64_t)(c * (16384 >> 4) + 2 * k);
            umma_tf32(tmem_base + 0, a, b, IDESC_N128, (uint32_t)(c | k));
            umma_tf32(tmem_base + 128, a + (uint64_t)(16384 >> 4), b, IDESC_N128, (uint32_t)(c | k));
          }
        umma_commit(kv_full);
        umma_commit(x_empty + 8 * (it & 1));
      }
      __syncwarp();
      mbar_wait(p_ready, it & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32_ts(tmem_base + 256, tmem_base + (uint32_t)(kc * 32 + k * 8), vt_desc + (uint64_t)(kc * (16384 >> 4) + 2 * k),
                         IDESC_N128, (uint32_t)(kc | k));
        umma_commit(ctx_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------- row threads --------------------------------------------
    const int q = warp & 3;                              // TMEM lane quarter; in the transposed tiles also the head
    const int gi = (warp - 4) >> 2;                      // LayerNorm: channel chunk; softmax: pixel columns 64gi..64gi+63
    const int r = q * 32 + lane;
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float2* exln = reinterpret_cast<float2*>(gbase + A_OFF_EXLN);
    float* exmax = reinterpret_cast<float*>(gbase + A_OFF_EXMAX);
    float* exz = reinterpret_cast<float*>(gbase + A_OFF_EXZ);
    float m = -INFINITY, Z = 0.f;
    float ctx[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) ctx[j] = 0.f;
    if (ntl > 0) {
      mbar_wait(x_full, 0);
      layernorm_row<false>(base + A_OFF_XA, gi, r, sw, exln, p.eps, 0);
      fence_async_proxy();
      mbar_arrive(a_ready);
    }
    int t_in = 0, fi = 0;
    for (int it = 0; it < ntl; ++it) {
      // ---- online softmax over the pixels of row (h,d) = r: this thread's 64 columns of kT ----
      mbar_wait(kv_full, it & 1);
      tc_fence_after();
      uint32_t a[32], b[32];
      tmem_ld32(tlane + 64 * gi, a);
      tmem_ld32(tlane + 64 * gi + 32, b);
      tmem_wait_ld();
      float tmax = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) tmax = fmaxf(tmax, fmaxf(__uint_as_float(a[j]), __uint_as_float(b[j])));
      exmax[gi * 128 + r] = tmax;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      tmax = fmaxf(tmax, exmax[(gi ^ 1) * 128 + r]);
      const float m_new = fmaxf(m, tmax);
      const float alpha = __expf(m - m_new);             // 0 on the first tile of a frame (m = -inf)
      m = m_new;
      float zs0 = 0.f, zs1 = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        a[j] = to_tf32(__expf(__uint_as_float(a[j]) - m_new));   // rounded once: Z sums exactly what the MMA multiplies
        b[j] = to_tf32(__expf(__uint_as_float(b[j]) - m_new));
        zs0 += __uint_as_float(a[j]);
        zs1 += __uint_as_float(b[j]);
      }
      Z = fmaf(Z, alpha, zs0 + zs1);
      tmem_st32(tlane + 64 * gi, a);
      tmem_st32(tlane + 64 * gi + 32, b);
      // ---- row (h,e) = r of vT -> K-major B operand (K = pixel): chunks 2gi, 2gi+1 ----
      tmem_ld32(tlane + 128 + 64 * gi, a);
      tmem_ld32(tlane + 128 + 64 * gi + 32, b);
      tmem_wait_ld();
      {
        const uint32_t v0 = base + A_OFF_VT + (uint32_t)((2 * gi) * 16384 + r * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sts128(v0 + ((j ^ sw) << 4), __uint_as_float(to_tf32(__uint_as_float(a[4 * j]))),
                 __uint_as_float(to_tf32(__uint_as_float(a[4 * j + 1]))), __uint_as_float(to_tf32(__uint_as_float(a[4 * j + 2]))),
                 __uint_as_float(to_tf32(__uint_as_float(a[4 * j + 3]))));
          sts128(v0 + 16384 + ((j ^ sw) << 4), __uint_as_float(to_tf32(__uint_as_float(b[4 * j]))),
                 __uint_as_float(to_tf32(__uint_as_float(b[4 * j + 1]))), __uint_as_float(to_tf32(__uint_as_float(b[4 * j + 2]))),
                 __uint_as_float(to_tf32(__uint_as_float(b[4 * j + 3]))));
        }
      }
      tmem_wait_st();
      fence_async_proxy();
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- LayerNorm of the next tile while the tensor core multiplies this one ----
      if (it + 1 < ntl) {
        const int nb = (it + 1) & 1;
        mbar_wait(x_full + 8 * nb, ((it + 1) >> 1) & 1);
        layernorm_row<false>(base + A_OFF_XA + nb * XA_BYTES, gi, r, sw, exln, p.eps, 0);
        fence_async_proxy();
        mbar_arrive(a_ready);
      }
      // ---- ctx = ctx * alpha + ctx_tile (own head block, columns 16gi..16gi+15) ----
      mbar_wait(ctx_full, it & 1);
      tc_fence_after();
      {
        uint32_t c16[16];
        tmem_ld16(tlane + 256 + 32 * q + 16 * gi, c16);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) ctx[j] = fmaf(ctx[j], alpha, __uint_as_float(c16[j]));
      }
      if (++t_in == tpf) {
        // ---- frame end: scale / Z, write the context, reset the running statistics ----
        exz[gi * 128 + r] = Z;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float inv = ATT_SCALE / (Z + exz[(gi ^ 1) * 128 + r]);
        const int frame = (int)blockIdx.x + fi * (int)gridDim.x;
        float4* dst = reinterpret_cast<float4*>(p.ctx + ((size_t)frame * 128 + r) * 32 + 16 * gi);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(ctx[4 * j] * inv, ctx[4 * j + 1] * inv, ctx[4 * j + 2] * inv, ctx[4 * j + 3] * inv);
        m = -INFINITY;
        Z = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) ctx[j] = 0.f;
        t_in = 0;
        ++fi;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// 2. MT[f][c][h*32+d] = sum_e Wout[c][h*32+e] ctx[f][h][d][e]   (rounded to TF32 once, it is an MMA operand of pass 3)
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
linattn_fold_out_kernel(const float* __restrict__ ctx, const float* __restrict__ w_out, float* __restrict__ mt) {
  __shared__ __align__(16) float s_ctx[HEADS * DH * DH];
  const float4* src = reinterpret_cast<const float4*>(ctx + (size_t)blockIdx.x * HEADS * DH * DH);
  for (int i = threadIdx.x; i < HEADS * DH * DH / 4; i += 256) reinterpret_cast<float4*>(s_ctx)[i] = __ldg(src + i);
  const int c = threadIdx.x & 63, h = threadIdx.x >> 6;
  float w[DH];
#pragma unroll
  for (int j = 0; j < DH / 4; ++j) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(w_out + c * HID + h * DH) + j);
    w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
  }
  __syncthreads();
  float* dst = mt + ((size_t)blockIdx.x * C + c) * HID + h * DH;
#pragma unroll 4
  for (int d4 = 0; d4 < DH / 4; ++d4) {
    float o[4];
#pragma unroll
    for (int dd = 0; dd < 4; ++dd) {
      const float* row = s_ctx + (h * DH + d4 * 4 + dd) * DH;
      float acc = 0.f;
#pragma unroll
      for (int e = 0; e < DH; ++e) acc = fmaf(w[e], row[e], acc);
      o[dd] = __uint_as_float(to_tf32(acc));
    }
    reinterpret_cast<float4*>(dst)[d4] = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------------------------------------------
// 3. apply pass
// ------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t B_OFF_W = 0;                           // 2 chunks x [128 rows x 128 B] (to_q)
constexpr uint32_t B_OFF_XA = 32768;                      // 2 buffers x XA_BYTES
constexpr uint32_t B_OFF_MT = B_OFF_XA + 2 * XA_BYTES;    // 2 buffers x 4 (h,d) chunks x [64 rows x 128 B]
constexpr uint32_t MT_BYTES = 32768;
constexpr uint32_t B_OFF_BIAS = B_OFF_MT + 2 * MT_BYTES;  // float [64]
constexpr uint32_t B_OFF_EXLN = B_OFF_BIAS + 256;
constexpr uint32_t B_OFF_BAR = B_OFF_EXLN + 2048;
constexpr uint32_t B_SMEM = B_OFF_BAR + 128 + 1024;

struct ApplyParams {
  const float* bias;   // [64] or null
  float eps;
  int ntiles, tpf;
};

__global__ void __launch_bounds__(THREADS, 1)
linattn_apply_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                        const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmM, const ApplyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(gbase + B_OFF_BIAS);
  const uint32_t bars = base + B_OFF_BAR;
  const uint32_t w_full = bars, x_full = bars + 8, x_empty = bars + 24, a_ready = bars + 40, q_full = bars + 48;
  const uint32_t qs_ready = bars + 56, y_full = bars + 64, tmem_slot = bars + 72;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(x_full + 8 * i, 1); mbar_init(x_empty + 8 * i, 8); }
    mbar_init(a_ready, 256);
    mbar_init(q_full, 1);
    mbar_init(qs_ready, 256);
    mbar_init(y_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmM) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < C) bias_s[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0 && lane == 0) {
    // ------------------------------------------- TMA producer -------------------------------------------
    mbar_expect_tx(w_full, 32768);
    for (int c = 0; c < 2; ++c) tma_load_2d(base + B_OFF_W + c * 16384, &tmW, w_full, c * 32, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(x_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
      mbar_expect_tx(x_full + 8 * buf, XA_BYTES + MT_BYTES);
      const uint32_t dst = base + B_OFF_XA + buf * XA_BYTES, dm = base + B_OFF_MT + buf * MT_BYTES;
      for (int c = 0; c < 2; ++c) tma_load_2d(dst + c * 16384, &tmX, x_full + 8 * buf, c * 32, tile * TP);
      const int frame = tile / p.tpf;
      for (int kc = 0; kc < 4; ++kc) tma_load_2d(dm + kc * 8192, &tmM, x_full + 8 * buf, kc * 32, frame * C);
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    const uint64_t w_desc = umma_desc(base + B_OFF_W);
    mbar_wait(w_full, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const uint64_t xa_desc = umma_desc(base + B_OFF_XA + (it & 1) * XA_BYTES);
      const uint64_t mt_desc = umma_desc(base + B_OFF_MT + (it & 1) * MT_BYTES);
      mbar_wait(a_ready, it & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + 0, xa_desc + (uint64_t)(c * (16384 >> 4) + 2 * k), w_desc + (uint64_t)(c * (16384 >> 4) + 2 * k),
                      IDESC_N128, (uint32_t)(c | k));
        umma_commit(q_full);
      }
      __syncwarp();
      mbar_wait(qs_ready, it & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32_ts(tmem_base + 128, tmem_base + (uint32_t)(kc * 32 + k * 8), mt_desc + (uint64_t)(kc * (8192 >> 4) + 2 * k),
                         IDESC_N64, (uint32_t)(kc | k));
        umma_commit(y_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------- row threads --------------------------------------------
    const int q = warp & 3;
    const int gi = (warp - 4) >> 2;                      // channel chunk gi (LayerNorm, output), heads 2gi and 2gi+1 (softmax)
    const int r = q * 32 + lane;                         // pixel of the tile == TMEM lane
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float2* exln = reinterpret_cast<float2*>(gbase + B_OFF_EXLN);
    int it = 0, pending_buf = -1;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t xa = base + B_OFF_XA + buf * XA_BYTES;
      const uint32_t mine = xa + (uint32_t)((gi * 4 + q) * 4096);   // this warp's rows of chunk gi; later its output box
      mbar_wait(x_full + 8 * buf, (it >> 1) & 1);
      layernorm_row<true>(xa, gi, r, sw, exln, p.eps, tlane + 192 + gi * 32);
      fence_async_proxy();
      tc_fence_before();
      mbar_arrive(a_ready);
      if (pending_buf >= 0 && lane == 0) {               // the previous tile's TMA store has finished reading its buffer
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(x_empty + 8 * pending_buf);
      }
      __syncwarp();
      // ---- q~ = softmax over d per head, in place in TMEM (the A operand of the second MMA) ----
      mbar_wait(q_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        const uint32_t col = tlane + (uint32_t)((2 * gi + hh) * DH);
        tmem_ld32(col, v);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = __expf(__uint_as_float(v[j]) - mx), e1 = __expf(__uint_as_float(v[j + 1]) - mx);
          v[j] = __float_as_uint(e0);
          v[j + 1] = __float_as_uint(e1);
          s0 += e0;
          s1 += e1;
        }
        const float inv = 1.0f / (s0 + s1);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = to_tf32(__uint_as_float(v[j]) * inv);
        tmem_st32(col, v);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(qs_ready);
      // ---- y + bias + raw x -> swizzled box -> TMA store ----
      mbar_wait(y_full, it & 1);
      tc_fence_after();
      {
        uint32_t yv[32], xv[32];
        tmem_ld32(tlane + 128 + gi * 32, yv);
        tmem_ld32(tlane + 192 + gi * 32, xv);
        tmem_wait_ld();
        const float* bs = bias_s + gi * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(mine + lane * 128 + ((j ^ sw) << 4), __uint_as_float(yv[4 * j]) + bs[4 * j] + __uint_as_float(xv[4 * j]),
                 __uint_as_float(yv[4 * j + 1]) + bs[4 * j + 1] + __uint_as_float(xv[4 * j + 1]),
                 __uint_as_float(yv[4 * j + 2]) + bs[4 * j + 2] + __uint_as_float(xv[4 * j + 2]),
                 __uint_as_float(yv[4 * j + 3]) + bs[4 * j + 3] + __uint_as_float(xv[4 * j + 3]));
      }
      fence_async_proxy();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmY, mine, gi * 32, tile * TP + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending_buf = buf;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// 3b. apply pass, software-pipelined (default): same arithmetic as linattn_apply_tc_kernel, but the two tensor-core steps of a
// tile no longer leave the row threads idle.  The accumulators are double-buffered in TMEM (2 x [q 128 | y 64 | raw x 64] = 512
// columns), the x tile ring has three slots (a slot doubles as the store staging of its tile, so the load of tile i+3 waits for
// the store of tile i) and the frame's folded out-projection MT has its own two-slot ring.  Row threads per tile:
//   softmax(i)  |  LayerNorm(i+1) while the y MMA(i) runs  |  epilogue(i) while the q MMA(i+1) runs
// ------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t P_OFF_W = 0;                           // 2 chunks x [128 rows x 128 B] (to_q)
constexpr uint32_t P_OFF_XA = 32768;                      // 3 slots x XA_BYTES
constexpr uint32_t P_OFF_MT = P_OFF_XA + 3 * XA_BYTES;    // 2 slots x 4 (h,d) chunks x [64 rows x 128 B]
constexpr uint32_t P_OFF_BIAS = P_OFF_MT + 2 * MT_BYTES;  // float [64]
constexpr uint32_t P_OFF_EXLN = P_OFF_BIAS + 256;
constexpr uint32_t P_OFF_BAR = P_OFF_EXLN + 2048;
constexpr uint32_t P_SMEM = P_OFF_BAR + 128 + 1024;

__global__ void __launch_bounds__(THREADS, 1)
linattn_apply_pipe_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                          const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmM, const ApplyParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* bias_s = reinterpret_cast<float*>(gbase + P_OFF_BIAS);
  const uint32_t bars = base + P_OFF_BAR;
  const uint32_t w_full = bars, x_full = bars + 8, x_empty = bars + 32, mt_full = bars + 56, mt_empty = bars + 72;
  const uint32_t a_ready = bars + 88, q_full = bars + 96, qs_ready = bars + 104, y_full = bars + 112, tmem_slot = bars + 120;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nloc = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(x_full + 8 * i, 1); mbar_init(x_empty + 8 * i, 8); }
    for (int i = 0; i < 2; ++i) { mbar_init(mt_full + 8 * i, 1); mbar_init(mt_empty + 8 * i, 1); }
    mbar_init(a_ready, 256);
    mbar_init(q_full, 1);
    mbar_init(qs_ready, 256);
    mbar_init(y_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmM) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x < C) bias_s[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0 && lane == 0) {
    // ------------------------------------------- TMA producer -------------------------------------------
    if (nloc > 0) {
      mbar_expect_tx(w_full, 32768);
      for (int c = 0; c < 2; ++c) tma_load_2d(base + P_OFF_W + c * 16384, &tmW, w_full, c * 32, 0);
    }
    for (int i = 0; i < nloc; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int xs = i % 3, ms = i & 1;
      mbar_wait(x_empty + 8 * xs, ((i / 3) & 1) ^ 1);
      mbar_expect_tx(x_full + 8 * xs, XA_BYTES);
      const uint32_t dst = base + P_OFF_XA + xs * XA_BYTES;
      for (int c = 0; c < 2; ++c) tma_load_2d(dst + c * 16384, &tmX, x_full + 8 * xs, c * 32, tile * TP);
      mbar_wait(mt_empty + 8 * ms, ((i >> 1) & 1) ^ 1);
      mbar_expect_tx(mt_full + 8 * ms, MT_BYTES);
      const uint32_t dm = base + P_OFF_MT + ms * MT_BYTES;
      const int frame = tile / p.tpf;
      for (int kc = 0; kc < 4; ++kc) tma_load_2d(dm + kc * 8192, &tmM, mt_full + 8 * ms, kc * 32, frame * C);
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    const uint64_t w_desc = umma_desc(base + P_OFF_W);
    auto issue_q = [&](int i) {
      const uint64_t xa_desc = umma_desc(base + P_OFF_XA + (i % 3) * XA_BYTES);
      const uint32_t tq = tmem_base + 256u * (uint32_t)(i & 1);
      mbar_wait(a_ready, i & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tq, xa_desc + (uint64_t)(c * (16384 >> 4) + 2 * k), w_desc + (uint64_t)(c * (16384 >> 4) + 2 * k), IDESC_N128,
                      (uint32_t)(c | k));
        umma_commit(q_full);
      }
      __syncwarp();
    };
    if (nloc > 0) {
      mbar_wait(w_full, 0);
      issue_q(0);
    }
    for (int i = 0; i < nloc; ++i) {
      const uint32_t ts = tmem_base + 256u * (uint32_t)(i & 1);
      const uint64_t mt_desc = umma_desc(base + P_OFF_MT + (i & 1) * MT_BYTES);
      mbar_wait(qs_ready, i & 1);
      mbar_wait(mt_full + 8 * (i & 1), (i >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kc = 0; kc < 4; ++kc)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32_ts(ts + 128, ts + (uint32_t)(kc * 32 + k * 8), mt_desc + (uint64_t)(kc * (8192 >> 4) + 2 * k), IDESC_N64,
                         (uint32_t)(kc | k));
        umma_commit(y_full);
        umma_commit(mt_empty + 8 * (i & 1));
      }
      __syncwarp();
      if (i + 1 < nloc) issue_q(i + 1);
    }
  } else if (warp >= 4 && nloc > 0) {
    // ------------------------------------------- row threads --------------------------------------------
    const int q = warp & 3;
    const int gi = (warp - 4) >> 2;                      // channel chunk gi (LayerNorm, output), heads 2gi and 2gi+1 (softmax)
    const int r = q * 32 + lane;                         // pixel of the tile == TMEM lane
    const uint32_t sw = (uint32_t)(lane & 7);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float2* exln = reinterpret_cast<float2*>(gbase + P_OFF_EXLN);
    auto layernorm = [&](int i) {                        // LayerNorm of local tile i in place, raw x parked in its TMEM set
      const int xs = i % 3;
      mbar_wait(x_full + 8 * xs, (i / 3) & 1);
      layernorm_row<true>(base + P_OFF_XA + xs * XA_BYTES, gi, r, sw, exln, p.eps, tlane + 256u * (uint32_t)(i & 1) + 192 + gi * 32);
      fence_async_proxy();
      tc_fence_before();
      mbar_arrive(a_ready);
    };
    layernorm(0);
    for (int i = 0; i < nloc; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const uint32_t ts = tlane + 256u * (uint32_t)(i & 1);
      const uint32_t mine = base + P_OFF_XA + (uint32_t)((i % 3) * XA_BYTES) + (uint32_t)((gi * 4 + q) * 4096);   // rows of chunk gi; later the output box
      // ---- q~ = softmax over d per head, in place in TMEM (the A operand of the second MMA) ----
      mbar_wait(q_full, i & 1);
      tc_fence_after();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        const uint32_t col = ts + (uint32_t)((2 * gi + hh) * DH);
        tmem_ld32(col, v);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = __expf(__uint_as_float(v[j]) - mx), e1 = __expf(__uint_as_float(v[j + 1]) - mx);
          v[j] = __float_as_uint(e0);
          v[j + 1] = __float_as_uint(e1);
          s0 += e0;
          s1 += e1;
        }
        const float inv = 1.0f / (s0 + s1);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = to_tf32(__uint_as_float(v[j]) * inv);
        tmem_st32(col, v);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(qs_ready);
      // ---- the previous tile's TMA store has finished reading its slot: hand it back to the producer ----
      if (i > 0 && lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(x_empty + 8 * ((i - 1) % 3));
      }
      __syncwarp();
      // ---- LayerNorm of the next tile while the tensor core multiplies this one ----
      if (i + 1 < nloc) layernorm(i + 1);
      // ---- y + bias + raw x -> swizzled box (in this tile's x slot) -> TMA store ----
      mbar_wait(y_full, i & 1);
      tc_fence_after();
      {
        uint32_t yv[32], xv[32];
        tmem_ld32(ts + 128 + gi * 32, yv);
        tmem_ld32(ts + 192 + gi * 32, xv);
        tmem_wait_ld();
        const float* bs = bias_s + gi * 32;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(mine + lane * 128 + ((j ^ sw) << 4), __uint_as_float(yv[4 * j]) + bs[4 * j] + __uint_as_float(xv[4 * j]),
                 __uint_as_float(yv[4 * j + 1]) + bs[4 * j + 1] + __uint_as_float(xv[4 * j + 1]),
                 __uint_as_float(yv[4 * j + 2]) + bs[4 * j + 2] + __uint_as_float(xv[4 * j + 2]),
                 __uint_as_float(yv[4 * j + 3]) + bs[4 * j + 3] + __uint_as_float(xv[4 * j + 3]));
      }
      fence_async_proxy();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(&tmY, mine, gi * 32, tile * TP + q * 32);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

static int make_map_2d(CUtensorMap* m, const float* ptr, int64_t cols, int64_t rows, int box_cols, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled failed", __FILE__, (int)r);
  return 0;
}

}  // namespace sl
}  // namespace dpc

extern "C" int dpc_spatial_linear_block_fused(const float* x, const float* w_qkv, const float* w_out, const float* b_out,
                                              float* ctx_ws, float* mt_ws, float* y, int32_t BF, int32_t HW, int32_t C,
                                              int32_t heads, float eps, void* stream) {
  using namespace dpc;
  using namespace dpc::sl;
  if (C != sl::C || heads != HEADS || HW % TP != 0) return -2;   // served by the unfused kernels
  DPC_CHECK_ARG(x && w_qkv && w_out && ctx_ws && mt_ws && y && BF > 0 && HW > 0);
  DPC_CHECK_ARG((int64_t)BF * HW < (int64_t)1 << 31);
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mx, my, mwkv, mwq, mm;
  int rc = make_map_2d(&mx, x, sl::C, (int64_t)BF * HW, 32, TP);
  if (rc) return rc;
  rc = make_map_2d(&my, y, sl::C, (int64_t)BF * HW, 32, 32);
  if (rc) return rc;
  rc = make_map_2d(&mwkv, w_qkv, sl::C, 3 * HID, 32, 256);
  if (rc) return rc;
  rc = make_map_2d(&mwq, w_qkv, sl::C, 3 * HID, 32, 128);
  if (rc) return rc;
  rc = make_map_2d(&mm, mt_ws, HID, (int64_t)BF * sl::C, 32, 64);
  if (rc) return rc;
  const int dev = device_ordinal();
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[dev];
  if (!configured) {
    DPC_CUDA(cudaFuncSetAttribute(linattn_context_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)A_SMEM));
    DPC_CUDA(cudaFuncSetAttribute(linattn_apply_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B_SMEM));
    DPC_CUDA(cudaFuncSetAttribute(linattn_apply_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P_SMEM));
    configured = true;
  }
  const int num_sms = sm_count(dev);
  CtxParams pa{ctx_ws, eps, BF, HW};
  linattn_context_tc_kernel<<<(unsigned)(BF < num_sms ? BF : num_sms), THREADS, A_SMEM, st>>>(mx, mwkv, pa);
  DPC_LAUNCH_CHECK();
  linattn_fold_out_kernel<<<(unsigned)BF, 256, 0, st>>>(ctx_ws, w_out, mt_ws);
  DPC_LAUNCH_CHECK();
  const int tpf = HW / TP, ntiles = BF * tpf;
  ApplyParams pb{b_out, eps, ntiles, tpf};
  static const int pipe = getenv("DPC_SL_PIPE") ? atoi(getenv("DPC_SL_PIPE")) : 1;   // 0: the unpipelined apply pass (A/B)
  if (pipe)
    linattn_apply_pipe_kernel<<<(unsigned)(ntiles < num_sms ? ntiles : num_sms), THREADS, P_SMEM, st>>>(mx, my, mwq, mm, pb);
  else
    linattn_apply_tc_kernel<<<(unsigned)(ntiles < num_sms ? ntiles : num_sms), THREADS, B_SMEM, st>>>(mx, my, mwq, mm, pb);
  DPC_LAUNCH_CHECK();
  return 0;
}
