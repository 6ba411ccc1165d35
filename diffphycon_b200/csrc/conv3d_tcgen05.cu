// 3x3x3 / pad 1 Conv3d (Block.proj, conv3d.py:189-204) on the 5th-generation tensor cores of sm_100a.
//
// Formulation: implicit GEMM, M = output voxels, N = Cout, K = 27 taps x Cin, TF32 operands, fp32 accumulate in TMEM.
// The implicit-GEMM A operand re-reads every input voxel 27 times; fed tap-by-tap from L2 this kernel is bound by the
// L2->SM path (measured: ~33 B/clk/SM), not by the tensor pipe.  So the nine in-plane taps share ONE shared-memory copy:
//   * per (dt, 32-channel chunk) a single TMA box  [32 ch] x [W+2] x [R rows]  (start (w,h) = (-1, hq-1); out-of-bounds
//     elements are zero-filled by TMA = the conv padding) lands in 128B-swizzled smem as rows of 128 B with pitch W+2;
//   * outputs are enumerated in the same padded-flat order  mu = ho*(W+2) + wo  (the two pad columns per image row are
//     computed and discarded: 3% at W=64), so tap (dh,dw) of the whole 128-row MMA operand is the SAME buffer shifted
//     by dh*(W+2)+dw rows.  The UMMA descriptor start address may be any multiple of 128 B: the 128B swizzle is a
//     function of the absolute smem address (verified on B200 with tools/umma_shift_probe.cu), no base_offset needed;
//   * one CTA owns up to S=4 consecutive 128-row sub-tiles (S accumulators side by side in TMEM), so each weight box
//     ([Cout] x [32], K-major) is amortised over S MMA groups.
// A traffic per tile drops from 27 box-equivalents to ~3.4; the kernel becomes tensor-pipe bound for Cout >= 64.
//
// Warp roles (256 threads): warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer (one lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld -> +bias -> GroupNorm partial statistics -> global store).
// Two mbarrier rings: A boxes (2 deep) and weight boxes (NB deep); tcgen05.commit releases slots and publishes the
// finished accumulators.  Every wait is bounded (trap after ~2 s) so a protocol bug cannot hang the device.
#include "common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace dpc {
namespace tc {

constexpr int KCH = 32;            // channels per K block: 32 fp32 = one 128-byte swizzle row
constexpr int ROW_BYTES = 128;
constexpr int NA = 2;              // A ring depth
constexpr int MAXS = 4;
constexpr int APARTS = 3;           // the A box of a block is fetched as APARTS row slabs

struct Params {
  const float* bias;
  float* y;
  double* gn_stats;
  int B, F, H, W;
  int C1, C2, Cout;
  int S, pitch, R, tiles_f;   // sub-tiles per CTA, smem row pitch (W+2), box rows, tiles per frame
  int NB;
  int a_bytes, b_bytes;
  int gn_groups;
  int debug;   // development switches (env DPC_TC_DEBUG): 1 = no MMA issue, 2 = no TMA loads/waits, 4 = no epilogue stores
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  while (true) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("dpc conv3d_tcgen05: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x,
             threadIdx.x, bar, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128B-swizzled operand descriptor: rows of 128 B, 8-row atoms 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major), bits [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version, bits [46,48)
  d |= (uint64_t)2 << 61;                               // layout type SWIZZLE_128B, bits [61,64)
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

__device__ long long g_tc_ts[16];   // development timestamps of one CTA (debug & 8)
#define TS(i) do { if ((p.debug & 8) && blockIdx.x == 1000) g_tc_ts[i] = clock64(); } while (0)

template <int N>
__global__ void __launch_bounds__(256, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmA1t, const __grid_constant__ CUtensorMap tmA2t,
                 const __grid_constant__ CUtensorMap tmW, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // 128B swizzle atoms need 1024-byte alignment
  const uint32_t a_buf = base;
  const uint32_t b_buf = base + NA * p.a_bytes;
  const uint32_t bars = b_buf + p.NB * p.b_bytes;         // 8-byte mbarriers
  const uint32_t fullA = bars, emptyA = bars + 8 * NA;
  const uint32_t fullB = bars + 16 * NA, emptyB = fullB + 8 * p.NB;
  const uint32_t accum_bar = emptyB + 8 * p.NB;
  const uint32_t tmem_slot = accum_bar + 8;
  __shared__ double s_stat[2][8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TS(0);
  const int tile = blockIdx.x;
  const int tf = tile % p.tiles_f;
  const int f = (tile / p.tiles_f) % p.F;
  const int b = tile / (p.tiles_f * p.F);
  const int mu_tile = tf * p.S * 128;                 // first padded-flat output position of this tile
  const int hq = mu_tile / p.pitch;                   // image row of that position
  const int mu0 = mu_tile - hq * p.pitch;             // its offset inside the row
  const int npos = p.H * p.pitch - mu_tile;           // positions left in the frame
  const int nsub = (npos >= p.S * 128) ? p.S : (npos + 127) / 128;
  const int Cin = p.C1 + p.C2;
  const int nch = Cin / KCH, nch1 = p.C1 / KCH;
  const int tmem_cols = (p.S * N <= 32) ? 32 : (p.S * N <= 64) ? 64 : (p.S * N <= 128) ? 128 : (p.S * N <= 256) ? 256 : 512;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    for (int i = 0; i < p.NB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (threadIdx.x < 16) s_stat[threadIdx.x >> 3][threadIdx.x & 7] = 0.0;
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (threadIdx.x == 0) TS(1);

  if (warp == 0 && lane == 0 && !(p.debug & 2)) {
    // ------------------------------------------- TMA producer -------------------------------------------
    // Issue order follows slot availability: the weight boxes of block j that fit the ring go first, then the A box of
    // block j+1 (its buffer frees when block j-1 retires) cut into APARTS row slabs interleaved with the remaining
    // weight boxes, so a 90 KB A transfer never sits in front of a weight box the tensor pipe is about to need.
    const int nblk = 3 * nch;
    const int rows_part = (p.R + APARTS - 1) / APARTS;
    auto issue_a_part = [&](int j, int part) {
      const int dt = j / nch, ch = j - dt * nch;
      const CUtensorMap* mp = (ch < nch1) ? &tmA1 : &tmA2;
      const int c0 = (ch < nch1) ? ch * KCH : (ch - nch1) * KCH;
      const int sa = j % NA;
      const int r0 = part * rows_part;
      if (r0 >= p.R) return;
      if (part == 0) {
        mbar_wait(emptyA + 8 * sa, ((j / NA) & 1) ^ 1);
        mbar_expect_tx(fullA + 8 * sa, (uint32_t)(p.R * p.pitch * ROW_BYTES));
      }
      const CUtensorMap* mpp = (r0 + rows_part <= p.R) ? mp : ((ch < nch1) ? &tmA1t : &tmA2t);
      tma_load_5d(a_buf + sa * p.a_bytes + (uint32_t)(r0 * p.pitch * ROW_BYTES), mpp, fullA + 8 * sa, c0, -1,
                  hq - 1 + r0, f + dt - 1, b);
    };
    int ib = 0;
    for (int part = 0; part < APARTS; ++part) issue_a_part(0, part);
    for (int j = 0; j < nblk; ++j) {
      const int dt = j / nch, ch = j - dt * nch;
      int next_part = 0;
      for (int t9 = 0; t9 < 9; ++t9) {
        const int sb = ib % p.NB;
        mbar_wait(emptyB + 8 * sb, ((ib / p.NB) & 1) ^ 1);
        mbar_expect_tx(fullB + 8 * sb, (uint32_t)p.b_bytes);
        tma_load_2d(b_buf + sb * p.b_bytes, &tmW, fullB + 8 * sb, (dt * 9 + t9) * Cin + ch * KCH, 0);
        ++ib;
        if (j + 1 < nblk && t9 + 1 >= p.NB - 1 && next_part < APARTS) issue_a_part(j + 1, next_part++);
      }
      while (j + 1 < nblk && next_part < APARTS) issue_a_part(j + 1, next_part++);
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    // The whole warp walks the pipeline (warp-uniform control flow keeps descriptors in uniform registers); one
    // elected lane issues.  Descriptors are advanced by integer adds on the 16-byte-unit start-address field:
    // +2 per K=8 step (32 B), +1024 per 128-row sub-tile, +(dh*pitch+dw)*8 per tap.
    // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    int ia = 0, ib = 0;
    uint32_t first = 0;                                                // becomes 1 after the first K block
    for (int dt = 0; dt < 3; ++dt) {
      for (int ch = 0; ch < nch; ++ch) {
        const int sa = ia % NA;
        if (!(p.debug & 2)) mbar_wait(fullA + 8 * sa, (ia / NA) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (ia == 0 && lane == 0) TS(2);
        const uint64_t adesc0 = umma_desc(a_buf + sa * p.a_bytes + (uint32_t)(mu0 * ROW_BYTES));
        for (int t9 = 0; t9 < 9; ++t9) {
          const int dh = t9 / 3, dw = t9 - dh * 3;
          const int sb = ib % p.NB;
          if (!(p.debug & 2)) mbar_wait(fullB + 8 * sb, (ib / p.NB) & 1);
          if (!(p.debug & 32)) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t bdesc = umma_desc(b_buf + sb * p.b_bytes);
          const uint64_t adesc = adesc0 + (uint64_t)((dh * p.pitch + dw) * (ROW_BYTES / 16));
          if ((p.debug & 64) ? (lane == 0) : elect_one()) {
#pragma unroll
            for (int s = 0; s < MAXS; ++s) {
              if (s < nsub && !(p.debug & 1)) {
#pragma unroll
                for (int k = 0; k < KCH / 8; ++k)
                  umma_tf32(tmem_base + (uint32_t)(s * N), adesc + (uint64_t)(s * (128 * ROW_BYTES / 16) + 2 * k),
                            bdesc + (uint64_t)(2 * k), idesc, first | (uint32_t)k);
              }
            }
            if (!(p.debug & 16)) umma_commit(emptyB + 8 * sb);     // weights slot free once these MMAs retire
          }
          __syncwarp();
          first = 1;
          ++ib;
        }
        if (elect_one()) umma_commit(emptyA + 8 * sa);
        __syncwarp();
        ++ia;
      }
    }
    if (elect_one()) umma_commit(accum_bar);
    __syncwarp();
    if (lane == 0) TS(3);
  } else if (warp >= 4) {
    // ------------------------------------------- epilogue -----------------------------------------------
    const int q = warp - 4;                       // TMEM lane quarter == warp_id % 4
    mbar_wait(accum_bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 128) TS(4);
    const bool do_stats = p.gn_stats != nullptr;
    // GroupNorm(8): group width N/8 columns, GPC groups per 32-column chunk; one (sum, sumsq) pair per group
    constexpr int GPC = 256 / N;        // 4, 2, 1 for N = 64, 128, 256
    constexpr int GW = 32 / GPC;        // columns of one group inside a chunk
    double gs[8], gq[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) gs[g] = gq[g] = 0.0;
    for (int s = 0; s < nsub; ++s) {
      const int mu = mu_tile + s * 128 + q * 32 + lane;   // padded-flat output position inside the frame
      const int h = mu / p.pitch, w = mu - h * p.pitch;
      const bool valid = (h < p.H) && (w < p.W);
      const size_t m = (((size_t)b * p.F + f) * p.H + h) * p.W + w;
      float* dst = p.y + m * N;
#pragma unroll
      for (int c = 0; c < N / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * N + c * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float o[32];
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + c * 32 + j));
          o[j] = __uint_as_float(v[j]) + bv.x;
          o[j + 1] = __uint_as_float(v[j + 1]) + bv.y;
          o[j + 2] = __uint_as_float(v[j + 2]) + bv.z;
          o[j + 3] = __uint_as_float(v[j + 3]) + bv.w;
        }
        if (valid && !(p.debug & 4)) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + c * 32 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          if (do_stats) {
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              float ps = 0.f, pq = 0.f;
#pragma unroll
              for (int j = 0; j < GW; ++j) {
                ps += o[g * GW + j];
                pq += o[g * GW + j] * o[g * GW + j];
              }
              gs[(c * GPC + g) % 8] += (double)ps;   // (c*GPC + g) < 8 by construction
              gq[(c * GPC + g) % 8] += (double)pq;
            }
          }
        }
      }
    }
    if (threadIdx.x == 128) TS(5);
    if (do_stats) {
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        double a = gs[g], bq = gq[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a += shfl_xor_double(a, o); bq += shfl_xor_double(bq, o); }
        if (lane == 0) {
          atomicAdd(&s_stat[0][g], a);
          atomicAdd(&s_stat[1][g], bq);
        }
      }
      // the four epilogue warps rendezvous on a named barrier, then publish one atomic per (group, statistic)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = threadIdx.x - 128;
      if (et < 16) {
        const int which = et >> 3, grp = et & 7;
        atomicAdd(p.gn_stats + ((size_t)b * 8 + grp) * 2 + which, s_stat[which][grp]);
      }
    }
  }
  if (threadIdx.x == 128) TS(6);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) TS(7);
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

static int make_act_map(CUtensorMap* m, const float* x, int B, int F, int H, int W, int C, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)F * H * W * C * 4};
  cuuint32_t box[5] = {(cuuint32_t)KCH, (cuuint32_t)box_w, (cuuint32_t)box_h, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(activation) failed", __FILE__, (int)r);
  return 0;
}

static int make_w_map(CUtensorMap* m, const float* w, int Kpad, int Npad, int N) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  cuuint64_t dims[2] = {(cuuint64_t)Kpad, (cuuint64_t)Npad};
  cuuint64_t strides[1] = {(cuuint64_t)Kpad * 4};
  cuuint32_t box[2] = {(cuuint32_t)KCH, (cuuint32_t)N};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(weights) failed", __FILE__, (int)r);
  return 0;
}

template <int N>
static int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& a1t, const CUtensorMap& a2t,
                  const CUtensorMap& wm, const Params& p, size_t smem, cudaStream_t st) {
  static size_t configured = 0;
  if (smem > configured) {
    DPC_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const unsigned grid = (unsigned)((size_t)p.B * p.F * p.tiles_f);
  conv3d_tc_kernel<N><<<grid, 256, smem, st>>>(a1, a2, a1t, a2t, wm, p);
  DPC_LAUNCH_CHECK();
  return 0;
}

}  // namespace tc
}  // namespace dpc

extern "C" int dpc_conv3d_tcgen05(const dpc_conv_params* pp, void* stream) {
  using namespace dpc;
  using namespace dpc::tc;
  DPC_CHECK_ARG(pp != nullptr);
  const dpc_conv_params& c = *pp;
  // ---- shape gate: anything else is served by dpc_conv_igemm (same numerics class) ----
  const int W = c.Wi, H = c.Hi, F = c.Fi;
  const bool shape_ok =
      c.ntaps == 27 && c.st == 1 && c.sh == 1 && c.sw == 1 && c.pt == 1 && c.ph == 1 && c.pw == 1 && c.Fo == F &&
      c.Ho == H && c.Wo == W && c.oh_mul == 1 && c.ow_mul == 1 && c.Hfull == H && c.Wfull == W && c.out_layout == 0 &&
      c.residual == nullptr && c.precise == 0 && c.C1 % KCH == 0 && c.C2 % KCH == 0 && c.C1 > 0 &&
      (c.Cout == 64 || c.Cout == 128 || c.Cout == 256) && c.Npad == c.Cout && W >= 2 && W + 2 <= 256 && H >= 1 &&
      c.Kpad == 27 * (c.C1 + c.C2);
  if (!shape_ok) return -2;
  if (c.gn_stats && c.gn_groups != 8) return -2;   // the epilogue is specialised for GroupNorm(8), the reference default
  DPC_CHECK_ARG(c.x1 && c.w && c.y && (c.C2 == 0 || c.x2));
  Params p;
  p.bias = c.bias; p.y = c.y; p.gn_stats = c.gn_stats; p.gn_groups = c.gn_groups;
  { const char* e = getenv("DPC_TC_DEBUG"); p.debug = e ? atoi(e) : 0; }
  p.B = c.B; p.F = F; p.H = H; p.W = W; p.C1 = c.C1; p.C2 = c.C2; p.Cout = c.Cout;
  p.pitch = W + 2;
  const int frame_pos = H * p.pitch;
  int S = 512 / c.Cout;
  if (S > MAXS) S = MAXS;
  if (S > (frame_pos + 127) / 128) S = (frame_pos + 127) / 128;
  const size_t budget = 227 * 1024 - 2048;
  p.b_bytes = c.Cout * ROW_BYTES;
  for (;; --S) {
    p.R = (p.pitch - 1 + S * 128 + 2 * p.pitch + 2 + p.pitch - 1) / p.pitch;
    p.a_bytes = ((p.R * p.pitch * ROW_BYTES + 1023) / 1024) * 1024;
    if ((size_t)NA * p.a_bytes + 3 * (size_t)p.b_bytes + 1024 + 256 <= budget || S == 1) break;
  }
  if ((size_t)NA * p.a_bytes + 3 * (size_t)p.b_bytes + 1024 + 256 > budget || p.R > 256) return -2;
  p.S = S;
  p.tiles_f = (frame_pos + S * 128 - 1) / (S * 128);
  int NB = (int)((budget - 1024 - 256 - (size_t)NA * p.a_bytes) / p.b_bytes);
  if (NB > 9) NB = 9;
  p.NB = NB;
  const size_t smem = (size_t)NA * p.a_bytes + (size_t)NB * p.b_bytes + 1024 + 256;
  // the A box is fetched as APARTS slabs of rows_part rows; the last slab may be shorter -> its own tensor map
  const int rows_part = (p.R + APARTS - 1) / APARTS;
  const int nfull = p.R / rows_part;
  const int tail_rows = p.R - nfull * rows_part;
  CUtensorMap a1, a2, a1t, a2t, wm;
  int rc = make_act_map(&a1, c.x1, c.B, F, H, W, c.C1, p.pitch, rows_part);
  if (rc) return rc;
  rc = make_act_map(&a1t, c.x1, c.B, F, H, W, c.C1, p.pitch, tail_rows > 0 ? tail_rows : rows_part);
  if (rc) return rc;
  if (c.C2) {
    rc = make_act_map(&a2, c.x2, c.B, F, H, W, c.C2, p.pitch, rows_part);
    if (rc) return rc;
    rc = make_act_map(&a2t, c.x2, c.B, F, H, W, c.C2, p.pitch, tail_rows > 0 ? tail_rows : rows_part);
    if (rc) return rc;
  } else {
    a2 = a1;
    a2t = a1t;
  }
  rc = make_w_map(&wm, c.w, c.Kpad, c.Npad, c.Cout);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (c.Cout == 64) return launch<64>(a1, a2, a1t, a2t, wm, p, smem, st);
  if (c.Cout == 128) return launch<128>(a1, a2, a1t, a2t, wm, p, smem, st);
  return launch<256>(a1, a2, a1t, a2t, wm, p, smem, st);
}

extern "C" int dpc_tc_debug_timestamps(long long* out16) {
  using namespace dpc;
  DPC_CUDA(cudaDeviceSynchronize());
  DPC_CUDA(cudaMemcpyFromSymbol(out16, dpc::tc::g_tc_ts, sizeof(long long) * 16));
  return 0;
}
