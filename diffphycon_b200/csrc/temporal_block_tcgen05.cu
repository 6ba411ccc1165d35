// Fused temporal-attention block of Unet3D_with_Conv3D at dim 64 (conv3d.py:165-174 LayerNorm, :293-352 Attention with
// RoPE + relative position bias, :153-157 Residual):   y = x + to_out(attn(rope(to_q(LN x)), rope(to_k(LN x)), to_v(LN x)))
// in ONE kernel, so the 384-wide qkv tensor (12.9 GB at the metric shape) and the 128-wide attention output never reach
// HBM: algorithmic traffic is one read of x and one write of y.
//
// Tile = 4 pixels x 32 frames = 128 tokens = the 128 rows of a tcgen05 MMA (row = pixel*32 + frame, so TMEM lane quarter q
// == pixel q == epilogue warp q, lane == frame).  Per tile:
//   TMA   : 8 boxes [32 ch] x [1 pixel] x [32 frames] -> 128B-swizzled smem (K-major A operand, two 32-channel chunks)
//   rows  : LayerNorm in place ((x-mean)*rstd; the gain is folded into the qkv weights on the host), raw x parked in TMEM
//   MMA   : qkv[128 x 384] = xhat[128 x 64] . Wqkv^T   (kind::tf32, N = 256 + 128, accumulators in TMEM columns 0..383)
//   rows  : per head: tcgen05.ld.16x256b hands q,k,v to the warp in the m16n8 accumulator-fragment layout; q*scale, RoPE(q,k);
//           k/v -> per-warp swizzled smem; S = q K^T (mma.sync m16n8k8 TF32, the d index permuted so that the TMEM
//           fragments ARE the A operand) + bias, softmax in registers, O = P V with the S fragments as the A operand,
//           O -> swizzled smem tile
//   MMA   : y[128 x 64] += O_h[128 x 32] . Wout[:, h*32:(h+1)*32]^T   (TMEM columns 384..447)
//   rows  : y + raw x (TMEM columns 448..511) -> swizzled smem -> TMA store
// Warp roles (384 threads): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-11 = 256 row threads in
// two groups: both map warp%4 -> pixel and lane -> frame; group 0 owns channel chunk 0 and heads 0,1, group 1 chunk 1 and
// heads 2,3 (two warps per scheduler hide the TMEM / smem / mma.sync latencies of each other).
// The x tile is double-buffered (the next tile lands while this one is processed); everything else is sequenced by the row
// threads, so the only cross-tile hazards are the two x buffers (x_full / x_empty).  All waits are bounded (tc_common.cuh).
#include "tc_common.cuh"

namespace dpc {
namespace tb {

using namespace dpc::tc;

constexpr int C = 64, HID = 128, NQKV = 384, FR = 32, HEADS = 4, DH = 32;
constexpr int THREADS = 384;
constexpr float ATT_SCALE = 0.17677669529663687f;       // 32^-0.5 (conv3d.py:287)

// shared-memory map (bytes from the 1024-aligned base)
constexpr uint32_t OFF_WQ = 0;                            // 2 chunks x [384 rows x 128 B]
constexpr uint32_t OFF_WO = 98304;                        // 4 chunks x [64 rows x 128 B]
constexpr uint32_t OFF_XA = 131072;                       // 2 buffers x 2 chunks x [128 rows x 128 B]
constexpr uint32_t XA_BYTES = 32768;
constexpr uint32_t OFF_OB = OFF_XA + 2 * XA_BYTES;        // [128 rows x 128 B]
constexpr uint32_t OFF_ROPE = OFF_OB + 16384;             // cos[32][20], sin[32][20] (row pitch 20: conflict-free fragment reads)
constexpr uint32_t OFF_BIAS = OFF_ROPE + 5120;            // [4][64]: bias of relative offset (j - i + 31)
constexpr uint32_t OFF_EXCH = OFF_BIAS + 1024;            // [2 groups][128 rows] float2 LayerNorm partial sums
constexpr uint32_t OFF_BAR = OFF_EXCH + 2048;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 128 + 1024;     // + alignment slack

struct Params {
  const float* rope_cos;   // [F][32]
  const float* rope_sin;
  const float* pos_bias;   // [4][F][F], function of (j - i) only
  float eps;
  int B, HW;
  int F;                   // frames <= 32: a pixel's rows F..31 of the tile are TMA zero fill, masked as keys, clipped by the store
};

__global__ void __launch_bounds__(THREADS, 1)
temporal_block_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                      const __grid_constant__ CUtensorMap tmWq, const __grid_constant__ CUtensorMap tmWo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  float* rope_s = reinterpret_cast<float*>(gbase + OFF_ROPE);
  float* bias_s = reinterpret_cast<float*>(gbase + OFF_BIAS);
  const uint32_t bars = base + OFF_BAR;
  const uint32_t w_full = bars, x_full = bars + 8, x_empty = bars + 24, a_ready = bars + 40, qkv_full = bars + 48;
  const uint32_t o_ready = bars + 56, o_free = bars + 64, y_full = bars + 72, tmem_slot = bars + 80;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_b = p.HW / 4;
  const int ntiles = p.B * tiles_b;

  if (threadIdx.x == 0) {
    mbar_init(w_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(x_full + 8 * i, 1); mbar_init(x_empty + 8 * i, 8); }
    mbar_init(a_ready, 256);
    mbar_init(qkv_full, 1);
    mbar_init(o_ready, 128);
    mbar_init(o_free, 1);
    mbar_init(y_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // tables: RoPE angle per (frame, pair) and the relative-position bias per (head, j - i + 31)
  for (int i = threadIdx.x; i < 32 * 16; i += THREADS) {
    const int f = i >> 4, pr = i & 15;
    rope_s[f * 20 + pr] = f < p.F ? __ldg(p.rope_cos + f * 32 + 2 * pr) : 1.f;
    rope_s[640 + f * 20 + pr] = f < p.F ? __ldg(p.rope_sin + f * 32 + 2 * pr) : 0.f;
  }
  for (int i = threadIdx.x; i < 4 * 64; i += THREADS) {
    const int h = i >> 6, d = (i & 63) - 31;
    float v = 0.f;
    if (d < p.F && -d < p.F) v = (d >= 0) ? __ldg(p.pos_bias + (h * p.F + 0) * p.F + d) : __ldg(p.pos_bias + (h * p.F - d) * p.F + 0);
    bias_s[i] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0 && lane == 0) {
    // ------------------------------------------- TMA producer -------------------------------------------
    mbar_expect_tx(w_full, 98304 + 32768);
    for (int c = 0; c < 2; ++c)
      for (int r = 0; r < 2; ++r)
        tma_load_2d(base + OFF_WQ + c * 49152 + r * 24576, &tmWq, w_full, c * 32, r * 192);
    for (int h = 0; h < 4; ++h) tma_load_2d(base + OFF_WO + h * 8192, &tmWo, w_full, h * 32, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(x_empty + 8 * buf, ((it >> 1) & 1) ^ 1);
      mbar_expect_tx(x_full + 8 * buf, XA_BYTES);
      const int b = tile / tiles_b, pix0 = (tile - b * tiles_b) * 4;
      const uint32_t dst = base + OFF_XA + buf * XA_BYTES;
      for (int c = 0; c < 2; ++c)
        for (int q = 0; q < 4; ++q) tma_load_4d(dst + c * 16384 + q * 4096, &tmX, x_full + 8 * buf, c * 32, pix0 + q, 0, b);
    }
  } else if (warp == 1) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc256 = idesc_base | ((uint32_t)(256 >> 3) << 17);
    const uint32_t idesc128 = idesc_base | ((uint32_t)(128 >> 3) << 17);
    const uint32_t idesc64 = idesc_base | ((uint32_t)(64 >> 3) << 17);
    const uint64_t wq_desc = umma_desc(base + OFF_WQ), wo_desc = umma_desc(base + OFF_WO), ob_desc = umma_desc(base + OFF_OB);
    mbar_wait(w_full, 0);
    int it = 0;
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const uint64_t xa_desc = umma_desc(base + OFF_XA + (it & 1) * XA_BYTES);
      mbar_wait(a_ready, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t a = xa_desc + (uint64_t)(c * (16384 >> 4) + 2 * k);
            const uint64_t b = wq_desc + (uint64_t)(c * (49152 >> 4) + 2 * k);
            umma_tf32(tmem_base + 0, a, b, idesc256, (uint32_t)(c | k));
            umma_tf32(tmem_base + 256, a, b + (uint64_t)((256 * 128) >> 4), idesc128, (uint32_t)(c | k));
          }
        umma_commit(qkv_full);
      }
      __syncwarp();
      for (int sl = 0; sl < HEADS; ++sl, ++n) {          // slot order: group 0 / group 1 alternate -> heads 0, 2, 1, 3
        const int h = (sl & 1) * 2 + (sl >> 1);
        mbar_wait(o_ready, n & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_tf32(tmem_base + 384, ob_desc + (uint64_t)(2 * k), wo_desc + (uint64_t)(h * (8192 >> 4) + 2 * k), idesc64,
                      (uint32_t)(sl | k));
          umma_commit(o_free);
          if (sl == HEADS - 1) umma_commit(y_full);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------- row threads --------------------------------------------
    const int q = warp & 3;                              // pixel of the tile == TMEM lane quarter
    const int gi = (warp - 4) >> 2;                      // group: channel chunk gi, heads 2gi and 2gi+1
    const int r = q * 32 + lane;                         // MMA row; lane == frame
    const int g = lane >> 2, t = lane & 3;               // mma.sync fragment coordinates
    const uint32_t sw = (uint32_t)(lane & 7);            // 128B-swizzle phase of row r (all tile regions are 1024-aligned)
    const uint32_t fvg = (uint32_t)(2 * ((g >> 1) & 3));
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float2* exch = reinterpret_cast<float2*>(gbase + OFF_EXCH);
    int it = 0;
    int pending_buf = -1;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      const uint32_t xa = base + OFF_XA + buf * XA_BYTES;
      const uint32_t mine = xa + (uint32_t)((gi * 4 + q) * 4096);   // this warp's 4 KB: V rows, later the output box
      mbar_wait(x_full + 8 * buf, (it >> 1) & 1);
      // ---- LayerNorm over the 64 channels of token r (this thread: chunk gi), in place; raw x parked in TMEM 448..511.
      //      Moments are accumulated about the row's first element (shifted single pass) and exchanged between the groups.
      {
        float x[32];
        const uint32_t xrow = xa + gi * 16384 + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = lds128(xrow + ((j ^ sw) << 4));
          x[j * 4 + 0] = v.x; x[j * 4 + 1] = v.y; x[j * 4 + 2] = v.z; x[j * 4 + 3] = v.w;
        }
        const float x0 = lds32(xa + r * 128 + (sw << 4));
        {
          uint32_t u[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(x[j]);
          tmem_st32(tlane + 448 + gi * 32, u);
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { x[j] -= x0; s1 += x[j]; s2 = fmaf(x[j], x[j], s2); }
        exch[gi * 128 + r] = make_float2(s1, s2);
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float2 o = exch[(gi ^ 1) * 128 + r];
        const float t1 = gi ? o.x + s1 : s1 + o.x, t2 = gi ? o.y + s2 : s2 + o.y;   // chunk 0 + chunk 1 in both groups
        const float dm = t1 * (1.0f / 64.0f);
        const float var = fmaxf(t2 * (1.0f / 64.0f) - dm * dm, 0.f);
        const float rstd = 1.0f / sqrtf(var + p.eps);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(xrow + ((j ^ sw) << 4), (x[4 * j] - dm) * rstd, (x[4 * j + 1] - dm) * rstd, (x[4 * j + 2] - dm) * rstd,
                 (x[4 * j + 3] - dm) * rstd);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      fence_async_proxy();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(a_ready);
      // the previous tile's TMA store has long finished reading its buffer: hand that buffer back to the producer
      if (pending_buf >= 0 && lane == 0) {
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(x_empty + 8 * pending_buf);
      }
      __syncwarp();
      mbar_wait(qkv_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * gi + hh;
        const uint32_t n = (uint32_t)(4 * it + 2 * hh + gi);   // out-projection slot (order: heads 0, 2, 1, 3)
        uint32_t qa[2][4][4];                            // A fragments of the rotated, scaled query (tf32 bits)
        uint32_t kb[4][4][2];                            // B fragments of the rotated key: [key block nt][d block kk]
        {
          uint32_t qf[2][16], kf[2][16], vf[2][16];      // C-fragment layout: [half][4*block + {0,1: row g; 2,3: row g+8}]
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const uint32_t tl = tlane + ((uint32_t)(hf * 16) << 16) + h * DH;
            tmem_ld_16x256b_x4(tl, qf[hf]);
            tmem_ld_16x256b_x4(tl + 128, kf[hf]);
            tmem_ld_16x256b_x4(tl + 256, vf[hf]);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          __syncwarp();                                  // every lane is done reading the previous head's V rows
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              const uint32_t vchunk = (uint32_t)(((2 * b + (t >> 1)) ^ fvg) << 4) + (uint32_t)((t & 1) * 8);
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {           // rows 16*hf + g and + 8; RoPE pair index 4*b + t
                const int row = 16 * hf + g + 8 * rr;
                const float cs = rope_s[row * 20 + 4 * b + t], sn = rope_s[640 + row * 20 + 4 * b + t];
                const float q0 = __fmul_rn(__uint_as_float(qf[hf][4 * b + 2 * rr]), ATT_SCALE);
                const float q1 = __fmul_rn(__uint_as_float(qf[hf][4 * b + 2 * rr + 1]), ATT_SCALE);
                const float k0 = __uint_as_float(kf[hf][4 * b + 2 * rr]), k1 = __uint_as_float(kf[hf][4 * b + 2 * rr + 1]);
                // t*cos + rotate_half(t)*sin, rotate_half: (x0, x1) -> (-x1, x0)   (rotary-embedding-torch 0.8.4)
                // MMA k positions (t, t+4) of d block b carry d = 8b+2t, 8b+2t+1.  Query (A operand): a0/a1 = rows g/g+8
                // at 2t, a2/a3 at 2t+1.  Key (B operand, n = key 8nt+g with nt = 2hf+rr): the accumulator-fragment layout of
                // the TMEM load IS the B-fragment layout, so the rotated key never leaves the registers.
                qa[hf][b][rr] = to_tf32(__fadd_rn(__fmul_rn(q0, cs), __fmul_rn(-q1, sn)));
                qa[hf][b][2 + rr] = to_tf32(__fadd_rn(__fmul_rn(q1, cs), __fmul_rn(q0, sn)));
                kb[2 * hf + rr][b][0] = to_tf32(__fadd_rn(__fmul_rn(k0, cs), __fmul_rn(-k1, sn)));
                kb[2 * hf + rr][b][1] = to_tf32(__fadd_rn(__fmul_rn(k1, cs), __fmul_rn(k0, sn)));
                sts64(mine + row * 128 + vchunk, __uint_as_float(vf[hf][4 * b + 2 * rr]), __uint_as_float(vf[hf][4 * b + 2 * rr + 1]));
              }
            }
          __syncwarp();
        }
        // ---- S = q K^T (m16n8k8 TF32) ----
        float sc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[mt][nt][e] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_tf32(sc[mt][nt], qa[mt][kk], kb[nt][kk][0], kb[nt][kk][1]);
        // ---- + relative bias, softmax over the 32 keys (row = 16mt + g + 8e2, key = 8nt + 2t + e) ----
        float inv[2][2];
        const float* bias_h = bias_s + h * 64 + 31;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int e2 = 0; e2 < 2; ++e2) {
            const int row = 16 * mt + g + 8 * e2;
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int key = 8 * nt + 2 * t + e;
                float v = sc[mt][nt][2 * e2 + e] + bias_h[key - row];
                if (key >= p.F) v = -INFINITY;              // zero-filled frames beyond F are not keys
                sc[mt][nt][2 * e2 + e] = v;
                mx = fmaxf(mx, v);
              }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float l = 0.f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float pv = __expf(sc[mt][nt][2 * e2 + e] - mx);
                sc[mt][nt][2 * e2 + e] = pv;
                l += pv;
              }
            l += __shfl_xor_sync(0xffffffffu, l, 1);
            l += __shfl_xor_sync(0xffffffffu, l, 2);
            inv[mt][e2] = 1.0f / l;
          }
        // ---- O = P V: the S accumulators are the A operand (k positions t, t+4 <-> keys 8kb+2t, 8kb+2t+1) ----
        float oc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int dn = 0; dn < 4; ++dn)
#pragma unroll
            for (int e = 0; e < 4; ++e) oc[mt][dn][e] = 0.f;
#pragma unroll
        for (int kj = 0; kj < 4; ++kj) {
          uint32_t pa[2][4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            pa[mt][0] = to_tf32(sc[mt][kj][0] * inv[mt][0]);   // (row g,   key 2t)
            pa[mt][1] = to_tf32(sc[mt][kj][2] * inv[mt][1]);   // (row g+8, key 2t)
            pa[mt][2] = to_tf32(sc[mt][kj][1] * inv[mt][0]);   // (row g,   key 2t+1)
            pa[mt][3] = to_tf32(sc[mt][kj][3] * inv[mt][1]);   // (row g+8, key 2t+1)
          }
          const uint32_t vrow = mine + (8 * kj + 2 * t) * 128 + (uint32_t)((g & 3) * 4);
          const uint32_t fv = (uint32_t)(2 * t);             // 2*(((8kj + 2t) >> 1) & 3)
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            const uint32_t a0 = vrow + (uint32_t)(((2 * dn + (g >> 2)) ^ fv) << 4);
            const uint32_t b0 = to_tf32(lds32(a0)), b1 = to_tf32(lds32(a0 + 128));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) mma_tf32(oc[mt][dn], pa[mt], b0, b1);
          }
        }
        // ---- O rows -> swizzled A operand of the out-projection (one buffer, slots alternate between the groups) ----
        mbar_wait(o_free, (n & 1) ^ 1);                  // the previous slot's MMAs have consumed the buffer
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int dn = 0; dn < 4; ++dn) {
            const uint32_t ob = base + OFF_OB + (uint32_t)((q * 32 + 16 * mt + g) * 128) +
                                (uint32_t)(((2 * dn + (t >> 1)) ^ g) << 4) + (uint32_t)((t & 1) * 8);
            sts64(ob, oc[mt][dn][0], oc[mt][dn][1]);
            sts64(ob + 8 * 128, oc[mt][dn][2], oc[mt][dn][3]);
          }
        fence_async_proxy();
        mbar_arrive(o_ready);
      }
      // ---- y = out-projection + raw x (this thread: chunk gi) -> swizzled box in this warp's 4 KB -> TMA store ----
      mbar_wait(y_full, it & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      __syncwarp();                                      // all lanes are done with this warp's V rows
      {
        uint32_t yv[32], xv[32];
        tmem_ld32(tlane + 384 + gi * 32, yv);
        tmem_ld32(tlane + 448 + gi * 32, xv);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(mine + lane * 128 + ((j ^ sw) << 4), __uint_as_float(yv[4 * j]) + __uint_as_float(xv[4 * j]),
                 __uint_as_float(yv[4 * j + 1]) + __uint_as_float(xv[4 * j + 1]),
                 __uint_as_float(yv[4 * j + 2]) + __uint_as_float(xv[4 * j + 2]),
                 __uint_as_float(yv[4 * j + 3]) + __uint_as_float(xv[4 * j + 3]));
      }
      fence_async_proxy();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        const int b = tile / tiles_b, pix = (tile - b * tiles_b) * 4 + q;
        tma_store_4d(&tmY, mine, gi * 32, pix, 0, b);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending_buf = buf;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
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
    DPC_CUDA(cudaFuncSetAttribute(temporal_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    configured = true;
  }
  const int num_sms = sm_count(dev);
  Params p{rope_cos, rope_sin, pos_bias, eps, B, HW, F};
  const int ntiles = B * (HW / 4);
  const unsigned grid = (unsigned)(ntiles < num_sms ? ntiles : num_sms);
  temporal_block_kernel<<<grid, THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(mx, my, mq, mo, p);
  DPC_LAUNCH_CHECK();
  return 0;
}
