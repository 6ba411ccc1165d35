// Generic implicit-GEMM convolution / linear layer: cp.async multi-stage pipeline, ldmatrix fragment loads,
// mma.sync m16n8k8 TF32 with fp32 accumulation.  Handles every conv shape of Unet3D_with_Conv3D that the
// TMA/tcgen05 kernel (conv3d_tcgen05.cu) does not: the 7x7x7 stem, 1x4x4 strided down-conv, the four parity classes
// of the transposed up-conv, 1x1x1 convs and the attention Linear layers (reference: conv3d.py:159-163, :189-204,
// :214, :240-241, :288-289, :403, :471).  Virtual concat of two channels-last sources replaces torch.cat
// (conv3d.py:538, :545).  Epilogue: bias, residual, GroupNorm partial statistics (sum / sum of squares in double).
#include "common.cuh"

namespace dpc {

constexpr int BM = 128;
constexpr int BK = 32;
constexpr int LDS_ = BK + 4;  // padded smem row (floats): 144 B rows keep ldmatrix conflict-free
constexpr int STAGES = 3;
constexpr int NTHREADS = 256;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// split an fp32 word into a TF32-exact "big" part and the (exact) remainder
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& big, uint32_t& small) {
  big = x & 0xffffe000u;
  small = __float_as_uint(__uint_as_float(x) - __uint_as_float(big));
}

template <int BN, bool PRECISE>
__global__ void __launch_bounds__(NTHREADS, PRECISE ? 1 : 2)
conv_igemm_kernel(const dpc_conv_params p, const int M, const int tilesN, const int K) {
  constexpr int WN = BN / 2;   // warp tile N (2 warps along N)
  constexpr int NT = WN / 8;   // n8 tiles per warp
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                              // [STAGES][BM][LDS_]
  float* Bs = smem + STAGES * BM * LDS_;         // [STAGES][BN][LDS_]
  __shared__ double s_stat[2][64];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int tile_n = blockIdx.x % tilesN;
  const int tile_m = blockIdx.x / tilesN;
  const int m0 = tile_m * BM;
  const int n0 = tile_n * BN;
  const int Cin = p.C1 + p.C2;

  // ---- loader row bookkeeping: this thread copies k-vector (tid&7) of rows (tid>>3) + 32*i ----
  const int kvec = tid & 7;
  int fi0[4], hi0[4], wi0[4], brow[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + (tid >> 3) + 32 * i;
    if (m < M) {
      int wo = m % p.Wo;
      int t1 = m / p.Wo;
      int ho = t1 % p.Ho;
      int t2 = t1 / p.Ho;
      int fo = t2 % p.Fo;
      int b = t2 / p.Fo;
      fi0[i] = fo * p.st - p.pt;
      hi0[i] = ho * p.sh - p.ph;
      wi0[i] = wo * p.sw - p.pw;
      brow[i] = ((b * p.Fi + fi0[i]) * p.Hi + hi0[i]) * p.Wi + wi0[i];
    } else {
      fi0[i] = -(1 << 28);
      hi0[i] = 0;
      wi0[i] = 0;
      brow[i] = 0;
    }
  }
  const int4* taps = reinterpret_cast<const int4*>(p.taps);
  const int nk = p.Kpad / BK;

  auto load_stage = [&](int stage, int kc) {
    // A operand (implicit im2col gather, zero fill for padding / K tail / M tail)
    const int kg = kc * BK + kvec * 4;
    const bool kvalid = kg < K;
    int tap = 0, ci = kg;
    if (p.ntaps > 1) {
      tap = kg / Cin;
      ci = kg - tap * Cin;
    }
    int4 tp = make_int4(0, 0, 0, 0);
    if (kvalid) tp = __ldg(&taps[tap]);
    const float* src_base;
    int cs, cc;
    if (ci < p.C1) {
      src_base = p.x1; cs = p.C1; cc = ci;
    } else {
      src_base = p.x2; cs = p.C2; cc = ci - p.C1;
    }
    float* a_dst = As + (size_t)stage * BM * LDS_ + (tid >> 3) * LDS_ + kvec * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      bool v = kvalid && (unsigned)(fi0[i] + tp.x) < (unsigned)p.Fi && (unsigned)(hi0[i] + tp.y) < (unsigned)p.Hi &&
               (unsigned)(wi0[i] + tp.z) < (unsigned)p.Wi;
      const float* src = v ? src_base + (size_t)(brow[i] + tp.w) * cs + cc : p.x1;
      cp_async16(smem_u32(a_dst + i * 32 * LDS_), src, v ? 16 : 0);
    }
    // B operand: packed weights [Npad][Kpad]
    float* b_dst = Bs + (size_t)stage * BN * LDS_ + (tid >> 3) * LDS_ + kvec * 4;
    const float* wsrc = p.w + (size_t)(n0 + (tid >> 3)) * p.Kpad + kc * BK + kvec * 4;
#pragma unroll
    for (int i = 0; i < BN / 32; ++i) cp_async16(smem_u32(b_dst + i * 32 * LDS_), wsrc + (size_t)i * 32 * p.Kpad, 16);
  };

  const int wm = warp & 3;   // 4 warps along M: rows wm*32
  const int wn = warp >> 2;  // 2 warps along N: cols wn*WN
  float acc[2][NT][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  // ldmatrix lane addressing ("fragment mapping")
  const int lmat = lane >> 3, lrow = lane & 7;
  const int a_row_off = wm * 32 + lrow + (lmat & 1) * 8;
  const int a_k_off = (lmat >> 1) * 4;
  const int b_row_off = wn * WN + lrow + (lmat >> 1) * 8;
  const int b_k_off = (lmat & 1) * 4;

  for (int kc = 0; kc < nk; ++kc) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      int nxt = kc + STAGES - 1;
      if (nxt < nk) load_stage(nxt % STAGES, nxt);
      cp_async_commit();
    }
    const int stage = kc % STAGES;
    const uint32_t a_base = smem_u32(As + (size_t)stage * BM * LDS_);
    const uint32_t b_base = smem_u32(Bs + (size_t)stage * BN * LDS_);
#pragma unroll
    for (int k8 = 0; k8 < BK / 8; ++k8) {
      uint32_t af[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
        ldmatrix_x4(af[i][0], af[i][1], af[i][2], af[i][3],
                    a_base + (uint32_t)(((a_row_off + i * 16) * LDS_ + k8 * 8 + a_k_off) * 4));
#pragma unroll
      for (int j2 = 0; j2 < NT / 2; ++j2) {
        uint32_t bf[4];
        ldmatrix_x4(bf[0], bf[1], bf[2], bf[3],
                    b_base + (uint32_t)(((b_row_off + j2 * 16) * LDS_ + k8 * 8 + b_k_off) * 4));
        if constexpr (!PRECISE) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            mma_tf32(acc[i][2 * j2], af[i], bf[0], bf[1]);
            mma_tf32(acc[i][2 * j2 + 1], af[i], bf[2], bf[3]);
          }
        } else {
          uint32_t bb[4], bs[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) split_tf32(bf[r], bb[r], bs[r]);
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            uint32_t ab[4], as[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) split_tf32(af[i][r], ab[r], as[r]);
            mma_tf32(acc[i][2 * j2], as, bb[0], bb[1]);
            mma_tf32(acc[i][2 * j2], ab, bs[0], bs[1]);
            mma_tf32(acc[i][2 * j2], ab, bb[0], bb[1]);
            mma_tf32(acc[i][2 * j2 + 1], as, bb[2], bb[3]);
            mma_tf32(acc[i][2 * j2 + 1], ab, bs[2], bs[3]);
            mma_tf32(acc[i][2 * j2 + 1], ab, bb[2], bb[3]);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  // ------------------------------------------------ epilogue ------------------------------------------------
  const int g = lane >> 2, t = lane & 3;
  const int rows_per_sample = p.Fo * p.Ho * p.Wo;
  const bool plain_rows = (p.oh_mul == 1 && p.ow_mul == 1 && p.Hfull == p.Ho && p.Wfull == p.Wo);
  const bool do_stats = p.gn_stats != nullptr;
  const int cpg = do_stats ? p.Cout / p.gn_groups : 1;
  const bool uniform_sample = do_stats && (m0 + BM <= M) && (m0 / rows_per_sample == (m0 + BM - 1) / rows_per_sample) &&
                              (cpg % 8 == 0 || cpg == 4 || cpg == 2) && (BN / cpg <= 64) && (BN % cpg == 0);
  if (do_stats && uniform_sample) {
    if (tid < 128) s_stat[tid >> 6][tid & 63] = 0.0;
    __syncthreads();
  }

  float bias_v[NT][2];
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    int n = n0 + wn * WN + j * 8 + 2 * t;
    bias_v[j][0] = (p.bias && n < p.Cout) ? __ldg(p.bias + n) : 0.f;
    bias_v[j][1] = (p.bias && n + 1 < p.Cout) ? __ldg(p.bias + n + 1) : 0.f;
  }
  float tsum[NT], tsq[NT];
#pragma unroll
  for (int j = 0; j < NT; ++j) tsum[j] = tsq[j] = 0.f;

#pragma unroll
  for (int i = 0; i < 2; ++i) {
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const int m = m0 + wm * 32 + i * 16 + hrow * 8 + g;
      if (m >= M) continue;
      int wo = 0, ho = 0, fo = 0, b = 0;
      size_t orow = (size_t)m;
      if (!plain_rows || p.out_layout == 1 || (do_stats && !uniform_sample)) {
        wo = m % p.Wo;
        int t1 = m / p.Wo;
        ho = t1 % p.Ho;
        int t2 = t1 / p.Ho;
        fo = t2 % p.Fo;
        b = t2 / p.Fo;
        orow = ((size_t)(b * p.Fo + fo) * p.Hfull + (ho * p.oh_mul + p.oh_off)) * p.Wfull + (wo * p.ow_mul + p.ow_off);
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = n0 + wn * WN + j * 8 + 2 * t;
        if (n >= p.Cout) continue;
        float v0 = acc[i][j][hrow * 2 + 0] + bias_v[j][0];
        float v1 = acc[i][j][hrow * 2 + 1] + bias_v[j][1];
        const bool has1 = (n + 1 < p.Cout);
        if (p.out_layout == 0) {
          float* dst = p.y + orow * p.Cout + n;
          if (p.residual) {
            const float* r = p.residual + orow * p.Cout + n;
            v0 += __ldg(r);
            if (has1) v1 += __ldg(r + 1);
          }
          if (has1 && ((p.Cout & 1) == 0)) {
            *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
          } else {
            dst[0] = v0;
            if (has1) dst[1] = v1;
          }
        } else {
          const size_t plane = (size_t)p.Hfull * p.Wfull;
          const size_t pix = (size_t)(ho * p.oh_mul + p.oh_off) * p.Wfull + (wo * p.ow_mul + p.ow_off);
          float* dst = p.y + ((size_t)(b * p.Fo + fo) * p.Cout + n) * plane + pix;
          dst[0] = v0;
          if (has1) dst[plane] = v1;
        }
        if (do_stats) {
          if (uniform_sample) {
            tsum[j] += v0 + (has1 ? v1 : 0.f);
            tsq[j] += v0 * v0 + (has1 ? v1 * v1 : 0.f);
          } else {
            // slow path (tile straddles samples / odd group width): per-element atomics
            double* st0 = p.gn_stats + ((size_t)b * p.gn_groups + n / cpg) * 2;
            atomicAdd(st0, (double)v0);
            atomicAdd(st0 + 1, (double)v0 * (double)v0);
            if (has1) {
              double* st1 = p.gn_stats + ((size_t)b * p.gn_groups + (n + 1) / cpg) * 2;
              atomicAdd(st1, (double)v1);
              atomicAdd(st1 + 1, (double)v1 * (double)v1);
            }
          }
        }
      }
    }
  }

  if (do_stats && uniform_sample) {
    // warp reduce in double over the lanes that share a group, then one shared atomic per (warp, group)
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      double s = (double)tsum[j], q = (double)tsq[j];
#pragma unroll
      for (int o = 16; o >= 4; o >>= 1) {
        s += shfl_xor_double(s, o);
        q += shfl_xor_double(q, o);
      }
      if (cpg >= 4) {
        s += shfl_xor_double(s, 1);
        q += shfl_xor_double(q, 1);
      }
      if (cpg >= 8) {
        s += shfl_xor_double(s, 2);
        q += shfl_xor_double(q, 2);
      }
      const int tmask = (cpg >= 8) ? 3 : (cpg == 4 ? 1 : 0);
      const int ncol = wn * WN + j * 8 + 2 * t;  // column inside the CTA tile
      if (g == 0 && (t & tmask) == 0 && (n0 + ncol) < p.Cout) {
        atomicAdd(&s_stat[0][ncol / cpg], s);
        atomicAdd(&s_stat[1][ncol / cpg], q);
      }
    }
    __syncthreads();
    const int ngl = BN / cpg;
    if (tid < 2 * ngl) {
      const int which = tid / ngl, gl = tid % ngl;
      const int gglob = n0 / cpg + gl;
      if (gglob < p.gn_groups) {
        const int b = m0 / rows_per_sample;
        atomicAdd(p.gn_stats + ((size_t)b * p.gn_groups + gglob) * 2 + which, s_stat[which][gl]);
      }
    }
  }
}

template <int BN, bool PRECISE>
static int launch(const dpc_conv_params& p, int M, int K, cudaStream_t st) {
  const int tilesN = p.Npad / BN;
  const int tilesM = (M + BM - 1) / BM;
  const size_t smem = (size_t)STAGES * (BM + BN) * LDS_ * sizeof(float);
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[device_ordinal()];
  if (!configured) {
    DPC_CUDA(cudaFuncSetAttribute(conv_igemm_kernel<BN, PRECISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  conv_igemm_kernel<BN, PRECISE><<<(unsigned)((size_t)tilesM * tilesN), NTHREADS, smem, st>>>(p, M, tilesN, K);
  DPC_LAUNCH_CHECK();
  return 0;
}

}  // namespace dpc

extern "C" int dpc_conv_igemm(const dpc_conv_params* pp, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(pp != nullptr);
  const dpc_conv_params& p = *pp;
  DPC_CHECK_ARG(p.x1 && p.w && p.y && p.taps);
  DPC_CHECK_ARG(p.C1 > 0 && p.C1 % 4 == 0 && p.C2 >= 0 && p.C2 % 4 == 0);
  DPC_CHECK_ARG(p.C2 == 0 || p.x2 != nullptr);
  DPC_CHECK_ARG(p.Npad % 64 == 0 && p.Kpad % BK == 0 && p.Cout <= p.Npad);
  DPC_CHECK_ARG(p.ntaps >= 1 && (int64_t)p.ntaps * (p.C1 + p.C2) <= p.Kpad);
  DPC_CHECK_ARG(p.out_layout == 0 || (p.residual == nullptr));
  DPC_CHECK_ARG(p.res_scale == nullptr && p.res_shift == nullptr);   // the folded GroupNorm residual is a dpc_conv3d_tcgen05 feature
  DPC_CHECK_ARG(p.gn_stats == nullptr || (p.gn_groups > 0 && p.Cout % p.gn_groups == 0));
  const int64_t M64 = (int64_t)p.B * p.Fo * p.Ho * p.Wo;
  const int64_t Min = (int64_t)p.B * p.Fi * p.Hi * p.Wi;
  DPC_CHECK_ARG(M64 > 0 && M64 < (1LL << 31) && Min < (1LL << 31));
  const int M = (int)M64;
  const int K = p.ntaps * (p.C1 + p.C2);
  cudaStream_t st = (cudaStream_t)stream;
  const bool wide = (p.Npad % 128 == 0);
  if (p.precise) return wide ? launch<128, true>(p, M, K, st) : launch<64, true>(p, M, K, st);
  return wide ? launch<128, false>(p, M, K, st) : launch<64, false>(p, M, K, st);
}
