// Stem convolution of Unet3D_with_Conv3D (init_conv, conv3d.py:392: Conv3d(channels, dim, 7x7x7, padding 3)) on tcgen05.
//
// Cin is tiny (2 or 6 channels), so a K block cannot be "32 channels of one tap" as in conv3d_tcgen05.cu.  Instead the input
// is kept channels-last in planes of 4 channels (16 B per pixel) and a K block is the 8 dw taps x 4 channels of one
// (dt, dh, plane) = 32 floats (kw = 7 real taps + one zero weight column).  In shared memory the raw image rows
// [R rows][PW pixels][4 floats] land unswizzled (one TMA box per (dt, plane), out-of-bounds = zero = the conv padding), and
// the A operand of every MMA is a K-major NO-SWIZZLE descriptor with OVERLAPPING core matrices: 8-pixel groups 128 B apart
// (SBO), the two 16-byte K core matrices of one K = 8 step 16 B apart (LBO), so  A[m][k] = raw[4 (m + shift) + k]  is the
// sliding window of the convolution, no im2col copy (layout verified on B200 with tools/umma_noswizzle_probe.cu).
// Outputs are enumerated in padded-flat order mu = h * PW + w' (PW = W + 6; columns w' >= W are computed and discarded), so tap (dh, dw) of a 128-row operand is the same buffer shifted by dh * PW + dw pixels.
// Weights stream through a TMA ring as 128B-swizzled [N][32] boxes; S = 4 accumulators of 128 x N share each weight box.
// Warp roles as conv3d_tcgen05.cu: warp 0 weight TMA producer, warp 3 input TMA producer, warp 1 MMA issuer, warp 2 TMEM
// allocator, warps 4-11 epilogue.
#include "tc_common.cuh"

namespace dpc {
namespace stem {

using namespace dpc::tc;

constexpr int NA = 3;              // A ring depth
constexpr int MAXS = 4;
constexpr int THREADS = 384;

struct Params {
  const float* bias;
  float* y;
  int B, F, H, W, N;
  int kt, kh, pt, ph;         // temporal / vertical taps and paddings
  int P;                      // 4-channel planes
  int PW, R, S, tiles_f;
  int NB, a_bytes, a_tx, b_bytes;
  int AB, tmem_cols;
};

__device__ __forceinline__ uint64_t desc_window(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(16 >> 4) << 16;                       // LBO: K-adjacent core matrix = next pixel
  d |= (uint64_t)(128 >> 4) << 32;                      // SBO: next 8 pixels
  d |= (uint64_t)1 << 46;
  return d;                                             // layout type 0 (no swizzle)
}

__global__ void __launch_bounds__(THREADS, 1)
stem_conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t b_buf = base;                            // swizzled weight boxes first (1024-aligned)
  const uint32_t a_buf = base + p.NB * p.b_bytes;
  const uint32_t bars = a_buf + NA * p.a_bytes;
  const uint32_t fullA = bars, emptyA = bars + 8 * NA;
  const uint32_t fullB = bars + 16 * NA, emptyB = fullB + 8 * p.NB;
  const uint32_t acc_full = emptyB + 8 * p.NB, acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = p.kt * p.P;                            // A boxes per tile
  const int ntiles = p.B * p.F * p.tiles_f;
  const int N = p.N;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    for (int i = 0; i < p.NB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 3 && lane == 0) {
    // ------------------------------------------- TMA producer: input rows ------------------------------
    // (own thread: the weight stream below blocks on its own ring)
    int sa = 0;
    uint32_t pha = 1;
    for (int a_tile = blockIdx.x; a_tile < ntiles; a_tile += gridDim.x) {
      const int tf = a_tile % p.tiles_f;
      const int f = (a_tile / p.tiles_f) % p.F;
      const int b = a_tile / (p.tiles_f * p.F);
      const int hq = (tf * p.S * 128) / p.PW;
      for (int a_blk = 0; a_blk < nblk; ++a_blk) {
        const int dt = a_blk / p.P, pl = a_blk - dt * p.P;
        mbar_wait(emptyA + 8 * sa, pha);
        mbar_expect_tx(fullA + 8 * sa, (uint32_t)p.a_tx);
        tma_load_5d(a_buf + sa * p.a_bytes, &tmA, fullA + 8 * sa, pl * 4, -3, hq - p.ph, f + dt - p.pt, b);
        if (++sa == NA) { sa = 0; pha ^= 1; }
      }
    }
  } else if (warp == 0 && lane == 0) {
    // ------------------------------------------- TMA producer: weights ---------------------------------
    int sb = 0;
    uint32_t phb = 1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int j = 0; j < nblk; ++j) {
        const int dt = j / p.P, pl = j - dt * p.P;
        for (int dh = 0; dh < p.kh; ++dh) {
          mbar_wait(emptyB + 8 * sb, phb);
          mbar_expect_tx(fullB + 8 * sb, (uint32_t)p.b_bytes);
          tma_load_2d(b_buf + sb * p.b_bytes, &tmW, fullB + 8 * sb, ((dt * p.kh + dh) * p.P + pl) * 32, 0);
          if (++sb == p.NB) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc_buf0 = desc_window(a_buf);
    const uint64_t bdesc_buf0 = umma_desc(b_buf);
    const uint32_t a_step = (uint32_t)(p.a_bytes >> 4), b_step = (uint32_t)(p.b_bytes >> 4);
    int sa = 0, sb = 0, ab = 0;
    uint32_t pha = 0, phb = 0, phacc = 1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tf = tile % p.tiles_f;
      const int mu_tile = tf * p.S * 128;
      const int mu0 = mu_tile - (mu_tile / p.PW) * p.PW;
      const int npos = p.H * p.PW - mu_tile;
      const int nsub = (npos >= p.S * 128) ? p.S : (npos + 127) / 128;
      mbar_wait(acc_empty + 8 * ab, phacc);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(ab * p.S * N);
      uint32_t first = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(fullA + 8 * sa, pha);
        tc_fence_after();
        uint64_t adesc = adesc_buf0 + (uint64_t)(sa * a_step + mu0);     // 16-byte units == pixels
#pragma unroll 1
        for (int dh = 0; dh < p.kh; ++dh, adesc += (uint64_t)p.PW) {
          mbar_wait(fullB + 8 * sb, phb);
          tc_fence_after();
          const uint64_t bdesc = bdesc_buf0 + (uint64_t)(sb * b_step);
          if (elect_one()) {
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
              if (s < nsub) {
#pragma unroll
                for (int k = 0; k < 4; ++k)               // K = 8: dw taps 2k, 2k+1 (x 4 channels)
                  umma_tf32(tacc + (uint32_t)(s * N), adesc + (uint64_t)(s * 128 + 2 * k), bdesc + (uint64_t)(2 * k), idesc,
                            first | (uint32_t)k);
              }
            }
            umma_commit(emptyB + 8 * sb);
          }
          __syncwarp();
          first = 1;
          if (++sb == p.NB) { sb = 0; phb ^= 1; }
        }
        if (elect_one()) umma_commit(emptyA + 8 * sa);
        __syncwarp();
        if (++sa == NA) { sa = 0; pha ^= 1; }
      }
      if (elect_one()) umma_commit(acc_full + 8 * ab);
      __syncwarp();
      if (++ab == p.AB) { ab = 0; phacc ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------- epilogue -----------------------------------------------
    const int q = warp & 3;
    const int eg = (warp - 4) >> 2;
    int ab = 0;
    uint32_t phacc = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tf = tile % p.tiles_f;
      const int f = (tile / p.tiles_f) % p.F;
      const int b = tile / (p.tiles_f * p.F);
      const int mu_tile = tf * p.S * 128;
      const int npos = p.H * p.PW - mu_tile;
      const int nsub = (npos >= p.S * 128) ? p.S : (npos + 127) / 128;
      mbar_wait(acc_full + 8 * ab, phacc);
      tc_fence_after();
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.S * N);
      for (int s = eg; s < nsub; s += 2) {
        const int mu = mu_tile + s * 128 + q * 32 + lane;
        const int h = mu / p.PW, w = mu - h * p.PW;
        const bool valid = (h < p.H) && (w < p.W);
        float* dst = p.y + ((((size_t)b * p.F + f) * p.H + h) * p.W + w) * (size_t)N;
#pragma unroll 1
        for (int c = 0; c < N / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)(s * N + c * 32), v);
          tmem_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + c * 32 + j));
              __stcs(reinterpret_cast<float4*>(dst + c * 32 + j),
                     make_float4(__uint_as_float(v[j]) + bv.x, __uint_as_float(v[j + 1]) + bv.y, __uint_as_float(v[j + 2]) + bv.z,
                                 __uint_as_float(v[j + 3]) + bv.w));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * ab);
      if (++ab == p.AB) { ab = 0; phacc ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
}

}  // namespace stem
}  // namespace dpc

extern "C" int dpc_stem_conv_tcgen05(const float* x, const float* w, const float* bias, float* y, int32_t B, int32_t F,
                                     int32_t H, int32_t W, int32_t Cpad, int32_t N, int32_t kt, int32_t kh, int32_t kw,
                                     void* stream) {
  using namespace dpc;
  using namespace dpc::stem;
  using namespace dpc::tc;
  if (Cpad % 4 != 0 || Cpad <= 0 || Cpad > 16 || (N != 32 && N != 64 && N != 128) || kw != 7 || kt < 1 || kh < 1 || !(kt & 1) ||
      !(kh & 1) || W < 8)
    return -2;
  DPC_CHECK_ARG(x && w && bias && y && B > 0 && F > 0 && H > 0);
  Params p;
  p.bias = bias; p.y = y; p.B = B; p.F = F; p.H = H; p.W = W; p.N = N;
  p.kt = kt; p.kh = kh; p.pt = kt / 2; p.ph = kh / 2; p.P = Cpad / 4;
  p.PW = W + 6;
  if (p.PW > 256) return -2;
  const int frame_pos = H * p.PW;
  int S = 512 / (2 * N);                                  // two accumulator sets
  if (S > MAXS) S = MAXS;
  if (S > (frame_pos + 127) / 128) S = (frame_pos + 127) / 128;
  p.S = S;
  p.AB = 2;
  { int need = p.AB * S * N; p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }
  // rows: offset inside the first row (< PW) + S*128 positions + (kh-1) halo rows + 8 pixels of dw reach
  const int span = p.PW - 1 + S * 128 + (kh - 1) * p.PW + 8;
  p.R = (span + p.PW - 1) / p.PW;
  if (p.R > 256) return -2;
  p.a_tx = p.R * p.PW * 16;
  p.a_bytes = (p.a_tx + 1023) / 1024 * 1024;
  p.b_bytes = N * 128;
  p.tiles_f = (frame_pos + S * 128 - 1) / (S * 128);
  const size_t budget = 227 * 1024 - 2048;
  int NB = (int)((budget - 1024 - 512 - (size_t)NA * p.a_bytes) / p.b_bytes);
  if (NB > 14) NB = 14;
  if (NB < 3) return -2;
  p.NB = NB;
  const size_t smem = (size_t)NA * p.a_bytes + (size_t)NB * p.b_bytes + 1024 + 512;
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  CUtensorMap ta, tw;
  {
    cuuint64_t dims[5] = {(cuuint64_t)Cpad, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)Cpad * 4, (cuuint64_t)W * Cpad * 4, (cuuint64_t)H * W * Cpad * 4,
                             (cuuint64_t)F * H * W * Cpad * 4};
    cuuint32_t box[5] = {4, (cuuint32_t)p.PW, (cuuint32_t)p.R, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(stem input) failed", __FILE__, (int)r);
  }
  {
    const int Ktot = kt * kh * p.P * 32;
    cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)Ktot * 4};
    cuuint32_t box[2] = {32, (cuuint32_t)N};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(stem weights) failed", __FILE__, (int)r);
  }
  const int dev = device_ordinal();
  static size_t configured_[kMaxDevices] = {};
  size_t& configured = configured_[dev];
  if (smem > configured) {
    DPC_CUDA(cudaFuncSetAttribute(stem_conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int num_sms = sm_count(dev);
  const size_t ntiles = (size_t)B * F * p.tiles_f;
  const unsigned grid = (unsigned)(ntiles < (size_t)num_sms ? ntiles : (size_t)num_sms);
  stem_conv_tc_kernel<<<grid, THREADS, smem, (cudaStream_t)stream>>>(ta, tw, p);
  DPC_LAUNCH_CHECK();
  return 0;
}
