// Fused temporal-attention block of Unet3D_with_Conv3D at dim 64 (conv3d.py:165-174 LayerNorm, :293-352 Attention with
// RoPE + relative position bias, :153-157 Residual):   y = x + to_out(attn(rope(to_q(LN x)), rope(to_k(LN x)), to_v(LN x)))
// in ONE kernel, so the 384-wide qkv tensor (12.9 GB at the metric shape) and the 128-wide attention output never reach
// HBM: algorithmic traffic is one read of x and one write of y.
//
// Tile = 4 pixels x 32 frames = 128 tokens = the 128 rows of a tcgen05 MMA (row = pixel*32 + frame, so TMEM lane quarter q
// == pixel q, lane == frame).  Per tile:
//   TMA   : 8 boxes [32 ch] x [1 pixel] x [32 frames] -> 128B-swizzled smem (K-major A operand, two 32-channel chunks)
//   LN    : LayerNorm in place ((x-mean)*rstd; the gain is folded into the qkv weights on the host); the raw x goes into the
//           y accumulator columns of TMEM, so the out-projection MMAs (always accumulating) add the residual for free
//   MMA   : qkv[128 x 384] = xhat[128 x 64] . Wqkv^T   (kind::tf32, N = 256 + 128, accumulators in TMEM columns 0..383)
//   heads : one warp per (pixel, head): tcgen05.ld.16x256b hands q,k,v to the warp in the m16n8 accumulator-fragment layout;
//           RoPE(q, k) in registers (rotated q goes back to TMEM in place), v -> the warp's 4 KB smem tile; S = q K^T
//           (mma.sync m16n8k8 TF32, the d index permuted so that the TMEM fragments ARE the operands), scale + bias, base-2
//           softmax in registers, O = P V with the S fragments as the A operand; the normalised O goes to TMEM where the
//           consumed q rows were = the [128 x 32] A operand (from TMEM) of that head's out-projection
//   MMA   : y[128 x 64] += O_h[128 x 32] . Wout[:, h*32:(h+1)*32]^T   (TMEM columns 384..447 / 448..511, alternating per tile)
//   store : y -> swizzled smem -> TMA store
// Warp roles (544 threads): warp 0 = control (TMEM allocation; lane 0 issues every TMA load and every tcgen05.mma);
// warps 1-16 = row warps, warp%4 -> pixel (TMEM lane quarter), (warp-1)/4 -> head.  The row warps of heads 0,1 also run the
// LayerNorm of the NEXT tile (32-channel chunk = head index) while those of heads 2,3 run the store epilogue of THIS tile
// (double-buffered y accumulator), so both sit in front of the qkv MMA of the next tile.  Round 1 ran two heads per warp on
// 8 row warps (2 per scheduler) and was stall-bound at 14.4 K clk per tile; four warps per scheduler hide the TMEM / shared
// memory / mma.sync latencies of each other.  All waits are bounded (tc_common.cuh).
#include "tc_common.cuh"

namespace dpc {
namespace tb {

using namespace dpc::tc;

constexpr int C = 64, HID = 128, NQKV = 384, FR = 32, HEADS = 4, DH = 32;
constexpr int THREADS = 32 + 512;
constexpr float ATT_SCALE = 0.17677669529663687f;       // 32^-0.5 (conv3d.py:287)
constexpr float LOG2E = 1.4426950408889634f;

// shared-memory map (bytes from the 1024-aligned base; the dynamic window starts 1024-aligned, checked at run time)
constexpr uint32_t OFF_WQ = 0;                            // 2 chunks x [384 rows x 128 B]
constexpr uint32_t OFF_WO = 98304;                        // 4 chunks x [64 rows x 128 B]
constexpr uint32_t OFF_XA = 131072;                       // 2 buffers x 2 chunks x [128 rows x 128 B]; after the qkv MMA: V / O tiles of heads 0,1
constexpr uint32_t XA_BYTES = 32768;
constexpr uint32_t OFF_VX = OFF_XA + 2 * XA_BYTES;        // V / O tiles of heads 2,3 (2 x [128 rows x 128 B]), then their store staging
constexpr uint32_t OFF_BIAS = OFF_VX + 32768;             // [4][64]: log2(e) * bias of relative offset (j - i + 31)
constexpr uint32_t OFF_ROPEA = OFF_BIAS + 1024;           // [8 rows g][16 pairs] (cos, sin) of frame g, pair slot XOR-swizzled by 4*(g&3)
constexpr uint32_t OFF_ROPEB = OFF_ROPEA + 1024;          // [4][16 pairs] (cos, sin) of frame 8i: frame g + 8i by angle addition
constexpr uint32_t OFF_BAR = OFF_ROPEB + 512;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128;

struct Params {
  const float* rope_cos;   // [F][32]
  const float* rope_sin;
  const float* pos_bias;   // [4][F][F], function of (j - i) only
  float eps;
  int B, HW;
  int F;                   // frames <= 32: a pixel's rows F..31 of the tile are TMA zero fill, masked as keys, clipped by the store
  int dbg;                 // development: bit 0 skip the head phase, bit 1 skip S/softmax/PV only, bit 2 skip the stores (DPC_TB_DBG)
};

// cvt.rna.tf32.f32 for finite inputs in one integer add (ptxas expands the cvt into a compare and a predicated add)
__device__ __forceinline__ uint32_t rtf32(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}

template <bool FULL>
__global__ void __launch_bounds__(THREADS, 1)
temporal_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                      const __grid_constant__ CUtensorMap tmWq, const __grid_constant__ CUtensorMap tmWo, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = smem_u32(smem_raw);
  uint8_t* gbase = smem_raw;
  float* bias_s = reinterpret_cast<float*>(gbase + OFF_BIAS);
  float2* ropeA_s = reinterpret_cast<float2*>(gbase + OFF_ROPEA);
  float2* ropeB_s = reinterpret_cast<float2*>(gbase + OFF_ROPEB);
  const uint32_t bars = base + OFF_BAR;
  const uint32_t w_full = bars, x_full = bars + 8, a_ready = bars + 24, qk_full = bars + 32, o_ready = bars + 40;
  const uint32_t y_full = bars + 72, v_full = bars + 80, y_read = bars + 88, tmem_slot = bars + 104;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_b = p.HW / 4;
  const int ntiles = p.B * tiles_b;

  if (threadIdx.x == 0) {
    if (base & 1023u) {
      printf("dpc temporal block: dynamic shared memory window is not 1024-byte aligned\n");
      __trap();
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(x_full + 8 * i, 1);
    mbar_init(a_ready, 256);
    mbar_init(qk_full, 1);
    mbar_init(v_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(y_read + 8 * i, 256);
    for (int i = 0; i < 4; ++i) mbar_init(o_ready + 8 * i, 128);
    mbar_init(y_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // relative-position bias per (head, j - i + 31), in the base-2 domain of the softmax
  for (int i = threadIdx.x; i < 4 * 64; i += THREADS) {
    const int h = i >> 6, d = (i & 63) - 31;
    float v = 0.f;
    if (d < p.F && -d < p.F) v = (d >= 0) ? __ldg(p.pos_bias + (h * p.F + 0) * p.F + d) : __ldg(p.pos_bias + (h * p.F - d) * p.F + 0);
    bias_s[i] = v * LOG2E;
  }
  // RoPE angle of frame g + 8i = angle(g) + angle(8i): two small tables instead of the [32][16] one (the shared memory is full);
  // frames >= F are clamped into the table (their rows are masked as keys and never stored)
  for (int i = threadIdx.x; i < 8 * 16 + 4 * 16; i += THREADS) {
    const bool isA = i < 128;
    const int row = isA ? (i >> 4) : 8 * ((i - 128) >> 4), pr = i & 15;
    const int rc = min(row, p.F - 1) * 32 + 2 * pr;
    const float2 v = make_float2(__ldg(p.rope_cos + rc), __ldg(p.rope_sin + rc));
    if (isA) ropeA_s[row * 16 + (pr ^ (4 * (row & 3)))] = v;
    else ropeB_s[(i - 128)] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------ control thread: TMA producer + MMA issuer ------------------------------
      auto load_x = [&](int tile, int buf) {
        mbar_expect_tx(x_full + 8 * buf, XA_BYTES);
        const int b = tile / tiles_b, pix0 = (tile - b * tiles_b) * 4;
        const uint32_t dst = base + OFF_XA + buf * XA_BYTES;
        for (int c = 0; c < 2; ++c)
          for (int q = 0; q < 4; ++q) tma_load_4d(dst + c * 16384 + q * 4096, &tmX, x_full + 8 * buf, c * 32, pix0 + q, 0, b);
      };
      mbar_expect_tx(w_full, 98304 + 32768);
      for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 2; ++r)
          tma_load_2d(base + OFF_WQ + c * 49152 + r * 24576, &tmWq, w_full, c * 32, r * 192);
      for (int h = 0; h < 4; ++h) tma_load_2d(base + OFF_WO + h * 8192, &tmWo, w_full, h * 32, 0);
      load_x(blockIdx.x, 0);
      if (blockIdx.x + (int)gridDim.x < ntiles) load_x(blockIdx.x + gridDim.x, 1);

      const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t idesc256 = idesc_base | ((uint32_t)(256 >> 3) << 17);
      const uint32_t idesc128 = idesc_base | ((uint32_t)(128 >> 3) << 17);
      const uint32_t idesc64 = idesc_base | ((uint32_t)(64 >> 3) << 17);
      const uint64_t wq_desc = umma_desc(base + OFF_WQ), wo_desc = umma_desc(base + OFF_WO);
      mbar_wait(w_full, 0);
      int it = 0;
      long long pc[6] = {0, 0, 0, 0, 0, 0}, pt = clock64();
      auto lap = [&](int i) { if (p.dbg & 8) { const long long t = clock64(); pc[i] += t - pt; pt = t; } };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        const uint64_t xa_desc = umma_desc(base + OFF_XA + buf * XA_BYTES);
        mbar_wait(a_ready, it & 1);
        lap(0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // q and k first (the head warps rotate them while the v projection still runs), then v
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + 0, xa_desc + (uint64_t)(c * (16384 >> 4) + 2 * k), wq_desc + (uint64_t)(c * (49152 >> 4) + 2 * k),
                      idesc256, (uint32_t)(c | k));
        umma_commit(qk_full);
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + 256, xa_desc + (uint64_t)(c * (16384 >> 4) + 2 * k),
                      wq_desc + (uint64_t)(c * (49152 >> 4) + 2 * k + ((256 * 128) >> 4)), idesc128, (uint32_t)(c | k));
        umma_commit(v_full);
        lap(1);
        // the previous tile's x buffer (xhat, then V of heads 0,1) is free once its out-projection has retired (it was issued
        // before these MMAs): prefetch the next tile into it
        if (it >= 1 && tile + (int)gridDim.x < ntiles) {
          mbar_wait(y_full, (it - 1) & 1);
          load_x(tile + gridDim.x, buf ^ 1);
        }
        lap(2);
        const uint32_t ycol = tmem_base + 384 + 64 * buf;
        for (int h = 0; h < HEADS; ++h) {
          // O_h sits in TMEM where q_h was (lane = token, columns 32h..32h+31): the A operand comes straight from TMEM
          mbar_wait(o_ready + 8 * h, it & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32_ts(ycol, tmem_base + (uint32_t)(h * DH + 8 * k), wo_desc + (uint64_t)(h * (8192 >> 4) + 2 * k), idesc64, 1u);
        }
        umma_commit(y_full);
        lap(3);
      }
      if ((p.dbg & 8) && blockIdx.x == 0 && it > 0)
        printf("temporal block control: per tile clk: wait a_ready %lld, issue qkv %lld, wait y_full(prev)+load %lld, wait o_ready+issue out %lld (%d tiles)\n",
               pc[0] / it, pc[1] / it, pc[2] / it, pc[3] / it, it);
    }
  } else {
    // ------------------------------------------- row warps ----------------------------------------------
    const int q = warp & 3;                              // pixel of the tile == TMEM lane quarter
    const int h = (warp - 1) >> 2;                       // head; heads 0,1: LayerNorm chunk h; heads 2,3: store chunk h-2
    const int r = q * 32 + lane;                         // MMA row of this thread in the LayerNorm / store phases; lane == frame
    const int g = lane >> 2, t = lane & 3;               // mma.sync fragment coordinates
    const uint32_t sw = (uint32_t)(lane & 7);            // 128B-swizzle phase of row r (all tile regions are 1024-aligned)
    const uint32_t fvg = (uint32_t)(2 * ((g >> 1) & 3));
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* bias_h = bias_s + h * 64 + 31;
    const uint32_t ropeA = base + OFF_ROPEA + (uint32_t)(g * 128 + t * 8), ropeB = base + OFF_ROPEB + (uint32_t)(t * 8);
    const uint32_t asw = (uint32_t)(g & 3);
    int it = 0;
    bool store_pending = false;
    int prev_tile = 0;
    long long pc[6] = {0, 0, 0, 0, 0, 0}, pt = clock64();
    auto lap = [&](int i) { if (p.dbg & 8) { const long long t = clock64(); pc[i] += t - pt; pt = t; } };

    auto store_epilogue = [&](int tile, int itx) {       // heads 2,3: y (residual included) of tile `tile` -> global
      const uint32_t mine = base + OFF_VX + (uint32_t)(((h - 2) * 4 + q) * 4096);
      mbar_wait(y_full, itx & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t yv[32];
      tmem_ld32(tlane + 384 + 64 * (itx & 1) + (h - 2) * 32, yv);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(y_read + 8 * (itx & 1));               // this y buffer may take the raw x of tile itx + 2
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(mine + lane * 128 + ((j ^ sw) << 4), __uint_as_float(yv[4 * j]), __uint_as_float(yv[4 * j + 1]),
               __uint_as_float(yv[4 * j + 2]), __uint_as_float(yv[4 * j + 3]));
      fence_async_proxy();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0 && !(p.dbg & 4)) {
        const int b = tile / tiles_b, pix = (tile - b * tiles_b) * 4 + q;
        tma_store_4d(&tmY, mine, (h - 2) * 32, pix, 0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    };

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t xa = base + OFF_XA + buf * XA_BYTES;
      const uint32_t mine = h < 2 ? xa + (uint32_t)((h * 4 + q) * 4096) : base + OFF_VX + (uint32_t)(((h - 2) * 4 + q) * 4096);
      lap(5);
      if (h < 2) {
        // ---- LayerNorm over the 64 channels of token r, in place for chunk h; raw x -> the y accumulator (residual).
        //      Moments about the row's first element (shifted single pass), chunk 0 then chunk 1 in both threads of a row.
        mbar_wait(x_full + 8 * buf, (it >> 1) & 1);
        lap(0);
        float x[32];
        const uint32_t xrow = xa + h * 16384 + r * 128, orow = xa + (h ^ 1) * 16384 + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds128(xrow + ((j ^ sw) << 4));
          x[j * 4 + 0] = v.x; x[j * 4 + 1] = v.y; x[j * 4 + 2] = v.z; x[j * 4 + 3] = v.w;
        }
        if (it >= 2) {                                     // the store epilogue of tile it - 2 has read this y buffer
          mbar_wait(y_read + 8 * buf, ((it >> 1) - 1) & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        {
          uint32_t u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(x[j]);
          tmem_st32(tlane + 384 + 64 * buf + h * 32, u);
        }
        const float x0 = lds32(xa + r * 128 + (sw << 4));
        float s1 = 0.f, s2 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { x[j] -= x0; s1 += x[j]; s2 = fmaf(x[j], x[j], s2); }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds128(orow + ((j ^ sw) << 4));
          const float a = v.x - x0, b = v.y - x0, c = v.z - x0, d = v.w - x0;
          o1 += a; o2 = fmaf(a, a, o2); o1 += b; o2 = fmaf(b, b, o2); o1 += c; o2 = fmaf(c, c, o2); o1 += d; o2 = fmaf(d, d, o2);
        }
        const float t1 = h ? o1 + s1 : s1 + o1, t2 = h ? o2 + s2 : s2 + o2;   // chunk 0 + chunk 1 in both threads
        const float dm = t1 * (1.0f / 64.0f);
        const float var = fmaxf(t2 * (1.0f / 64.0f) - dm * dm, 0.f);
        const float rstd = 1.0f / sqrtf(var + p.eps);
        // the partner thread (other chunk) reads this chunk for its moments: all reads of the row pair happen before the writes
        asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(xrow + ((j ^ sw) << 4), (x[4 * j] - dm) * rstd, (x[4 * j + 1] - dm) * rstd, (x[4 * j + 2] - dm) * rstd,
                 (x[4 * j + 3] - dm) * rstd);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        fence_async_proxy();
      } else {
        // ---- store epilogue of the previous tile (its y accumulator is the other TMEM buffer) ----
        if (it > 0) {
          store_epilogue(prev_tile, it - 1);
          store_pending = true;
        }
      }
      if (h < 2) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(a_ready);
      }
      prev_tile = tile;
      lap(1);
      mbar_wait(qk_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      lap(2);
      if (h >= 2 && store_pending) {                     // the TMA store has finished reading this warp's tile
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
      }

      // ================================ one (pixel, head) per warp ================================
      if (p.dbg & 1) {
        fence_async_proxy();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(o_ready + 8 * h);
        continue;
      }
      const uint32_t tq = tlane + (uint32_t)(h * DH);
      uint32_t kb[4][4][2];                              // B fragments of the rotated key: [key block nt][d block kk]
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const uint32_t lh = (uint32_t)(hf * 16) << 16;
        {
          uint32_t kf[16], qf[16];                       // C-fragment layout: [4*block + {0,1: row g; 2,3: row g+8}]
          tmem_ld_16x256b_x4(tq + 128 + lh, kf);
          tmem_ld_16x256b_x4(tq + lh, qf);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const float2 ra = lds64(ropeA + (((uint32_t)b ^ asw) << 5));          // frame g, pair 4b + t
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {             // rows 16*hf + g and + 8; RoPE pair index 4*b + t
              const float2 rb = lds64(ropeB + (uint32_t)((2 * hf + rr) * 128 + b * 32));   // frame 8*(2hf+rr)
              const float cs = fmaf(-ra.y, rb.y, ra.x * rb.x), sn = fmaf(ra.x, rb.y, ra.y * rb.x);
              const float k0 = __uint_as_float(kf[4 * b + 2 * rr]), k1 = __uint_as_float(kf[4 * b + 2 * rr + 1]);
              const float q0 = __uint_as_float(qf[4 * b + 2 * rr]), q1 = __uint_as_float(qf[4 * b + 2 * rr + 1]);
              // t*cos + rotate_half(t)*sin, rotate_half: (x0, x1) -> (-x1, x0)   (rotary-embedding-torch 0.8.4)
              // MMA k positions (t, t+4) of d block b carry d = 8b+2t, 8b+2t+1.  Key (B operand, n = key 8nt+g with
              // nt = 2hf+rr): the accumulator-fragment layout of the TMEM load IS the B-fragment layout.
              kb[2 * hf + rr][b][0] = rtf32(fmaf(-k1, sn, k0 * cs));
              kb[2 * hf + rr][b][1] = rtf32(fmaf(k0, sn, k1 * cs));
              qf[4 * b + 2 * rr] = rtf32(fmaf(-q1, sn, q0 * cs));
              qf[4 * b + 2 * rr + 1] = rtf32(fmaf(q0, sn, q1 * cs));
            }
          }
          tmem_st_16x256b_x4(tq + lh, qf);               // rotated query back in place: part 2 reloads it as the A fragments
        }
        {
          uint32_t vf[16];
          if (hf == 0) {
            mbar_wait(v_full, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          tmem_ld_16x256b_x4(tq + 256 + lh, vf);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t vchunk = (uint32_t)(((2 * b + (t >> 1)) ^ fvg) << 4) + (uint32_t)((t & 1) * 8);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
              sts64(mine + (16 * hf + g + 8 * rr) * 128 + vchunk, __uint_as_float(rtf32(__uint_as_float(vf[4 * b + 2 * rr]))),
                    __uint_as_float(rtf32(__uint_as_float(vf[4 * b + 2 * rr + 1]))));
          }
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      __syncwarp();                                      // V rows visible to the whole warp
      lap(3);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t lm = (uint32_t)(mt * 16) << 16;
        float oc[4][4];
        float inv[2];
        if (p.dbg & 2) {
#pragma unroll
          for (int dn = 0; dn < 4; ++dn)
#pragma unroll
            for (int e = 0; e < 4; ++e) oc[dn][e] = __uint_as_float(kb[dn][e][mt]);
          inv[0] = inv[1] = 1.f;
        } else {
        // ---- rotated query rows 16mt + g, + 8 as the A operand: a0/a1 = rows g/g+8 at 2t, a2/a3 at 2t+1 ----
        uint32_t qa[4][4];
        {
          uint32_t qf[16];
          tmem_ld_16x256b_x4(tq + lm, qf);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            qa[b][0] = qf[4 * b]; qa[b][1] = qf[4 * b + 2]; qa[b][2] = qf[4 * b + 1]; qa[b][3] = qf[4 * b + 3];
          }
        }
        // ---- S = q K^T (m16n8k8 TF32) ----
        float sc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) sc[nt][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(sc[nt], qa[kk], kb[nt][kk][0], kb[nt][kk][1]);
        // ---- scale (base-2 domain) + relative bias, softmax over the 32 keys (row = 16mt + g + 8e2, key = 8nt + 2t + e);
        //      P stays unnormalised ----
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
          const int row = 16 * mt + g + 8 * e2;
          float mx = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int key = 8 * nt + 2 * t + e;
              float v = fmaf(sc[nt][2 * e2 + e], ATT_SCALE * LOG2E, bias_h[key - row]);
              if (!FULL && key >= p.F) v = -INFINITY;    // zero-filled frames beyond F are not keys
              sc[nt][2 * e2 + e] = v;
              mx = fmaxf(mx, v);
            }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float l = 0.f;
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float pv = ex2(sc[nt][2 * e2 + e] - mx);
              sc[nt][2 * e2 + e] = pv;
              l += pv;
            }
          l += __shfl_xor_sync(0xffffffffu, l, 1);
          l += __shfl_xor_sync(0xffffffffu, l, 2);
          inv[e2] = rcp_approx(l);
        }
        // ---- O = P V: the S accumulators are the A operand (k positions t, t+4 <-> keys 8kb+2t, 8kb+2t+1) ----
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
#pragma unroll
          for (int e = 0; e < 4; ++e) oc[dn][e] = 0.f;
#pragma unroll
        for (int kj = 0; kj < 4; ++kj) {
          uint32_t pa[4];
          pa[0] = rtf32(sc[kj][0]);                        // (row g,   key 2t)
          pa[1] = rtf32(sc[kj][2]);                        // (row g+8, key 2t)
          pa[2] = rtf32(sc[kj][1]);                        // (row g,   key 2t+1)
          pa[3] = rtf32(sc[kj][3]);                        // (row g+8, key 2t+1)
          const uint32_t vrow = mine + (8 * kj + 2 * t) * 128 + (uint32_t)((g & 3) * 4);
          const uint32_t fv = (uint32_t)(2 * t);           // 2*(((8kj + 2t) >> 1) & 3)
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            const uint32_t a0 = vrow + (uint32_t)(((2 * dn + (g >> 2)) ^ fv) << 4);
            const uint32_t b0 = __float_as_uint(lds32(a0)), b1 = __float_as_uint(lds32(a0 + 128));
            mma_tf32(oc[dn], pa, b0, b1);
          }
        }
        }
        // ---- normalised O rows -> TMEM where the (consumed) rotated query rows were: A operand of the out-projection ----
        {
          uint32_t ov[16];
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            ov[4 * dn + 0] = __float_as_uint(oc[dn][0] * inv[0]);
            ov[4 * dn + 1] = __float_as_uint(oc[dn][1] * inv[0]);
            ov[4 * dn + 2] = __float_as_uint(oc[dn][2] * inv[1]);
            ov[4 * dn + 3] = __float_as_uint(oc[dn][3] * inv[1]);
          }
          tmem_st_16x256b_x4(tq + lm, ov);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      fence_async_proxy();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(o_ready + 8 * h);
      lap(4);
    }
    if ((p.dbg & 8) && blockIdx.x == 0 && lane == 0 && q == 0 && it > 0)
      printf("temporal block head %d: per tile clk: wait x_full %lld, LN or epilogue %lld, wait qk_full %lld, part 1 %lld, part 2 %lld, loop %lld\n", h,
             pc[0] / it, pc[1] / it, pc[2] / it, pc[3] / it, pc[4] / it, pc[5] / it);
    if (h >= 2) {                                        // drain: the last tile's store
      if (it > 0) store_epilogue(prev_tile, it - 1);
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

static int make_tok_map(CUtensorMap* m, const float* x, int B, int HW, int F) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  // the box always spans 32 frames: frames >= F are out of bounds = zero fill on load, clipped on store
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)HW, (cuuint64_t)F, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)HW * C * 4, (cuuint64_t)F * HW * C * 4};
  cuuint32_t box[4] = {32, 1, (cuuint32_t)FR, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(tokens) failed", __FILE__, (int)r);
  return 0;
}

static int make_w_map(CUtensorMap* m, const float* w, int K, int rows, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(weights) failed", __FILE__, (int)r);
  return 0;
}

}  // namespace tb
}  // namespace dpc

extern "C" int dpc_temporal_block_fused(const float* x, const float* w_qkv, const float* w_out, const float* rope_cos,
                                        const float* rope_sin, const float* pos_bias, float* y, int32_t B, int32_t F,
                                        int32_t HW, int32_t C, int32_t heads, float eps, void* stream) {
  using namespace dpc;
  using namespace dpc::tb;
  if (F < 1 || F > FR || C != tb::C || heads != HEADS || HW % 4 != 0) return -2;   // served by the unfused kernels
  DPC_CHECK_ARG(x && w_qkv && w_out && rope_cos && rope_sin && pos_bias && y && B > 0 && HW > 0);
  CUtensorMap mx, my, mq, mo;
  int rc = make_tok_map(&mx, x, B, HW, F);
  if (rc) return rc;
  rc = make_tok_map(&my, y, B, HW, F);
  if (rc) return rc;
  rc = make_w_map(&mq, w_qkv, tb::C, NQKV, 192);
  if (rc) return rc;
  rc = make_w_map(&mo, w_out, HID, tb::C, 64);
  if (rc) return rc;
  const int dev = device_ordinal();
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[dev];
  if (!configured) {
    DPC_CUDA(cudaFuncSetAttribute(temporal_block_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    DPC_CUDA(cudaFuncSetAttribute(temporal_block_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    configured = true;
  }
  const int num_sms = sm_count(dev);
  static const int dbg = getenv("DPC_TB_DBG") ? atoi(getenv("DPC_TB_DBG")) : 0;
  Params p{rope_cos, rope_sin, pos_bias, eps, B, HW, F, dbg};
  const int ntiles = B * (HW / 4);
  const unsigned grid = (unsigned)(ntiles < num_sms ? ntiles : num_sms);
  if (F == FR)
    temporal_block_kernel<true><<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(mx, my, mq, mo, p);
  else
    temporal_block_kernel<false><<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(mx, my, mq, mo, p);
  DPC_LAUNCH_CHECK();
  return 0;
}
