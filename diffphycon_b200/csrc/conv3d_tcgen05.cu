// 3x3x3 / pad 1 Conv3d (Block.proj, conv3d.py:189-204) on the 5th-generation tensor cores of sm_100a; the same kernel also
// serves 1x1x1 convs / Linear layers ("gemm" mode), the 1x4x4 stride-2 down-conv and the ConvTranspose parity classes
// (2x2-tap modes), cta_group::2 CTA pairs (PAIR) and the opt-in stacked-temporal-tap tiles (quad).
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
// Warp roles (384 threads): warp 0 = weight TMA producer (one lane), warp 3 = activation TMA producer (one lane), warp 1 =
// tcgen05.mma issuer (one lane), warp 2 = TMEM allocator, warps 4-11 = epilogue (tcgen05.ld -> +bias -> GroupNorm partial statistics -> global store).
// Two mbarrier rings: A boxes (2 deep) and weight boxes (NB deep); tcgen05.commit releases slots and publishes the
// finished accumulators.  Every wait is bounded (trap after ~2 s) so a protocol bug cannot hang the device.
#include "tc_common.cuh"

namespace dpc {
namespace tc {

constexpr int KCH = 32;            // channels per K block: 32 fp32 = one 128-byte swizzle row
constexpr int ROW_BYTES = 128;
constexpr int MAXS = 4;
constexpr int NTHREADS_TC = 384;    // warps 0-3: TMA / MMA / TMEM alloc / idle, warps 4-11: two epilogue groups
constexpr int APARTS = 3;           // the A box of a block is fetched as APARTS row slabs

struct Params {
  const float* bias;
  const float* residual;      // optional, indexed like y
  const float* res_scale;     // optional [B][ldy]: the residual enters as silu(residual * scale + shift) (folded GroupNorm + SiLU)
  const float* res_shift;
  float* y;
  double* gn_stats;
  int B, F, H, W;
  int C1, C2, Cout;
  int S, pitch, R, tiles_f;   // sub-tiles per CTA, smem row pitch, box rows, tiles per frame
  int ndw;                    // 1: one halo box (pitch W+2) serves all nine in-plane taps; 3: one box per dw (pitch W)
  int gemm;                   // 1: 1x1x1 conv / Linear layer (single tap, no halo)
  int ncol;                   // gemm: column tiles of N outputs accumulated side by side in one pass over A
  int ldy, wrow0;             // output row pitch (floats) and first weight row / output column of this launch
  int NA, NB;                 // A ring depth, weight ring depth
  int a_bytes, b_bytes;
  int gn_groups;
  int AB, tmem_cols;
  int mode;                   // 0: 3x3x3 / Linear; 1: 1x4x4 stride-2 down-conv (four parity boxes fetched with TMA traversal
                              // stride 2, 2x2 taps each); 2: one parity class of the 1x4x4 stride-2 ConvTranspose (2x2 taps)
  int cls_h, cls_w;           // mode 2: output parity class
  int Hout, Wout, o_mul;      // output frame and position scale: out(h, w) -> (h*o_mul + cls_h, w*o_mul + cls_w)
  int quad;                   // PAIR + Cout = 64: Q = 2 or 4 (0 = off); a tile covers Q consecutive output frames; the temporal taps that one
                              // input frame feeds are stacked into one MMA of N = 64 / 128 / 192 (see the MMA issuer)
  int staged;                 // conv epilogue: stores staged through shared memory (128-byte row segments)
  int dbg;                    // DPC_TC_DEBUG experiment switches (1: weight boxes fetched once, 2: A boxes fetched once)
  int nkt, ptt;               // temporal taps / temporal padding: 3, 1 for the 3x3x3 conv; 1, 0 for a 3x3 conv over images (F = 1)
  int gcols;                  // GroupNorm: channels per group of the WHOLE layer (Cout / 8); a launch may cover a column slice
};

// cta_group::2 helpers (PAIR mode): the two CTAs of a cluster run one M = 256 MMA per instruction.  Each CTA keeps its own A
// tile (128 rows) and HALF of the weight box (N/2 rows) in its shared memory, so the weight operand is read once per pair
// instead of once per CTA (the tensor pipe is fed from shared memory at < 100 B/clk; measured).  The leader
// (cluster rank 0) issues the MMAs; every TMA of either CTA completes on the leader's "full" barriers; tcgen05.commit
// multicasts the "empty" / "accumulator full" arrivals to both CTAs.
__device__ __forceinline__ uint32_t leader_addr(uint32_t local) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(0));
  return r;
}
__device__ __forceinline__ void tma_load_5d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_leader, int c0, int c1, int c2,
                                                 int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_leader, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_leader, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
// Descriptors passed as (low word, high word): the per-MMA arithmetic only moves the 14-bit start-address field of the low
// word, so the single issuing thread does 32-bit adds instead of 64-bit ones (it has ~32 clk per N = 64 instruction).
template <bool PAIR>
__device__ __forceinline__ void umma_tf32_w(uint32_t tmem_c, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                            uint32_t idesc, uint32_t accumulate) {
  if (PAIR)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_c), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_c), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int N, bool PAIR>
__global__ void __launch_bounds__(NTHREADS_TC, 1)
conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmA1t, const __grid_constant__ CUtensorMap tmA2t,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW3, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;           // 128B swizzle atoms need 1024-byte alignment
  const int NA = p.NA;
  const uint32_t a_buf = base;
  const uint32_t b_buf = base + NA * p.a_bytes;
  const uint32_t bars = b_buf + p.NB * p.b_bytes;         // 8-byte mbarriers
  const uint32_t fullA = bars, emptyA = bars + 8 * NA;
  const uint32_t fullB = bars + 16 * NA, emptyB = fullB + 8 * p.NB;
  const uint32_t acc_full = emptyB + 8 * p.NB, acc_empty = acc_full + 16;
  const uint32_t tmem_slot = acc_empty + 16;
  float* stage_all = reinterpret_cast<float*>(smem_raw + (base - raw) + NA * p.a_bytes + p.NB * p.b_bytes + 256);  // epilogue staging tiles
  __shared__ double s_part[8][16];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cin = p.C1 + p.C2;
  const int nch = Cin / KCH, nch1 = p.C1 / KCH;
  const bool quad = PAIR && p.quad;
  const int Q = quad ? p.quad : 1;                      // output frames per tile
  const int nblk = quad ? (Q + 2) * nch : p.mode == 1 ? 4 * nch : p.mode == 2 ? nch : p.gemm ? nch : p.nkt * nch * p.ndw;   // A boxes per tile
  const int ntap = p.mode ? 4 : p.gemm ? p.ncol : 9 / p.ndw;                                       // weight boxes per A box
  const int AB = p.AB;                                   // TMEM accumulator sets (2 = epilogue overlaps the next tile)
  const int Fd = quad ? p.F / Q : p.F;                  // frames (or frame quads) per sample in the tile enumeration
  const int ntiles = p.B * Fd * p.tiles_f;
  // PAIR: the two CTAs of a cluster take the same spatial tile tf of two consecutive frames (same operand offsets in both
  // shared memories, which one shared A descriptor requires); iterations then count tile pairs.
  uint32_t cta_rank = 0;
  if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  const bool leader = cta_rank == 0;
  const int it_begin = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int it_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int it_end = PAIR ? ntiles / 2 : ntiles;
  auto tile_of = [&](int it) {
    if (!PAIR) return it;
    const int bfp = it / p.tiles_f;
    return (2 * bfp + (int)cta_rank) * p.tiles_f + (it - bfp * p.tiles_f);
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
    for (int i = 0; i < p.NB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, PAIR ? 16 : 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(p.tmem_cols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();                          // both CTAs' barriers are initialised before any remote signal
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (quad) {
    // quad mode always accumulates (an MMA spans accumulators in different states), so the accumulators start at zero
    // and the epilogue re-zeroes what it has read
    if (warp >= 4 && warp < 8) {
      uint32_t z[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) z[j] = 0u;
      for (int c = 0; c < 512; c += 32) tmem_st32(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c, z);
      tmem_wait_st();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  if (warp == 3 && lane == 0) {
    // ------------------------------------------- TMA producer: activations -------------------------------
    // Own thread (the weight stream below blocks on its own ring): the A box of a block is fetched as APARTS row slabs as
    // soon as its slot frees, up to NA blocks ahead of the MMAs.
    const int rows_part = (p.R + APARTS - 1) / APARTS;
    int sa = 0;
    uint32_t pha = 1;                                     // producer parity: first pass over a fresh ring does not block
    int a_cnt = 0;
    const uint32_t fullA_l = PAIR ? leader_addr(fullA) : fullA;
    for (int a_it = it_begin; a_it < it_end; a_it += it_step) {
      const int a_tile = tile_of(a_it);
      const int tf = a_tile % p.tiles_f;
      int f = (a_tile / p.tiles_f) % Fd;
      const int b = a_tile / (p.tiles_f * Fd);
      const int hq = (tf * p.S * 128) / p.pitch;
      if (quad) f *= Q;                                      // first output frame of the tile
      for (int a_blk = 0; a_blk < nblk; ++a_blk, ++a_cnt) {
        int dt = 1, ch = a_blk, dwb = 1, halo = 0;           // gemm: centre tap only, no halo
        int w0 = 0, h0 = 0;
        if (quad) {                                          // input frame f - 1 + (a_blk / nch): dt carries the frame offset
          dt = a_blk / nch;
          ch = a_blk - dt * nch;
          dwb = 0;                                           // halo box: starts at column -1
          halo = 1;
        } else if (p.mode == 1) {                                   // parity box (ph, pw): input rows 2h + ph - 1 + 2*{0,1}, same in w
          const int par = a_blk / nch;
          ch = a_blk - par * nch;
          h0 = 2 * hq + ((par >> 1) ? -1 : 0);
          w0 = (par & 1) ? -1 : 0;
        } else if (p.mode == 2) {                            // class (cls_h, cls_w): input rows h - (1 - cls_h) + {0,1}
          h0 = hq - (1 - p.cls_h);
          w0 = -(1 - p.cls_w);
        } else if (!p.gemm) {
          dt = a_blk / (nch * p.ndw);
          const int rem = a_blk - dt * nch * p.ndw;
          ch = rem / p.ndw;
          dwb = rem - ch * p.ndw;
          halo = 1;
        }
        const bool src1 = ch < nch1;
        const int c0 = src1 ? ch * KCH : (ch - nch1) * KCH;
        mbar_wait(emptyA + 8 * sa, pha);
        const bool skip = !PAIR && (p.dbg & 2) && a_cnt >= NA;
        if (PAIR) { if (leader) mbar_expect_tx(fullA + 8 * sa, (uint32_t)(2 * p.R * p.pitch * ROW_BYTES)); }
        else if (skip) mbar_arrive(fullA + 8 * sa);
        else mbar_expect_tx(fullA + 8 * sa, (uint32_t)(p.R * p.pitch * ROW_BYTES));
        for (int r0 = 0; r0 < p.R && !skip; r0 += rows_part) {
          const CUtensorMap* mp = (r0 + rows_part <= p.R) ? (src1 ? &tmA1 : &tmA2) : (src1 ? &tmA1t : &tmA2t);
          const uint32_t dst = a_buf + sa * p.a_bytes + (uint32_t)(r0 * p.pitch * ROW_BYTES);
          if (p.mode) tma_load_5d(dst, mp, fullA + 8 * sa, c0, w0, h0 + (p.mode == 1 ? 2 * r0 : r0), f, b);
          else if (PAIR) tma_load_5d_pair(dst, mp, fullA_l + 8 * sa, c0, dwb - 1, hq - halo + r0, f + dt - p.ptt, b);
          else tma_load_5d(dst, mp, fullA + 8 * sa, c0, dwb - 1, hq - halo + r0, f + dt - p.ptt, b);
        }
        if (++sa == NA) { sa = 0; pha ^= 1; }
      }
    }
  } else if (warp == 0 && lane == 0) {
    // ------------------------------------------- TMA producer: weights ----------------------------------
    int sb = 0, b_cnt = 0;
    uint32_t phb = 1;
    const uint32_t fullB_l = PAIR ? leader_addr(fullB) : fullB;
    const int brow = PAIR ? (int)cta_rank * (N / 2) : 0;   // this CTA's half of the weight rows
    for (int it = it_begin; it < it_end; it += it_step) {
      for (int j = 0; j < nblk; ++j) {
        int dt = 0, ch = j, dwb = 0;
        int par = 0;
        if (p.mode) {
          par = j / nch;
          ch = j - par * nch;
        } else if (!p.gemm) {
          dt = j / (nch * p.ndw);
          const int rem = j - dt * nch * p.ndw;
          ch = rem / p.ndw;
          dwb = rem - ch * p.ndw;
        }
        if (quad) {
          // stacked weight rows R = dt*64 + n (a 3-D view of the packed weights): input frame index qi = j / nch feeds the
          // output frames of the quad through rows [r0, r0 + n); each CTA of the pair holds half of them
          const int qi = j / nch, qch = j - qi * nch;
          const int dt_lo = qi - Q + 1 > 0 ? qi - Q + 1 : 0, dt_hi = qi < 2 ? qi : 2;   // output frame i = qi - dt in [0, Q)
          const int r0 = dt_lo * 64;
          const int n = 64 * (dt_hi - dt_lo + 1);
          const int rmine = r0 + (int)cta_rank * (n / 2);
          for (int t = 0; t < 9; ++t) {
            mbar_wait(emptyB + 8 * sb, phb);
            if (leader) mbar_expect_tx(fullB + 8 * sb, (uint32_t)(n * ROW_BYTES));
            for (int q = 0; q < n / 64; ++q) {
              const int row = rmine + 32 * q;
              tma_load_3d_pair(b_buf + sb * p.b_bytes + q * (32 * ROW_BYTES), &tmW3, fullB_l + 8 * sb, t * Cin + qch * KCH, row & 63,
                               row >> 6);
            }
            if (++sb == p.NB) { sb = 0; phb ^= 1; }
          }
          continue;
        }
        // weight column of tap (dt, dh, dw): ((dt*3 + dh)*3 + dw)*Cin + ch*32; per box either all nine (dh,dw) or the three dh
        const int k0 = p.gemm ? ch * KCH : (dt * 9 + dwb) * Cin + ch * KCH;
        const int kstep = (p.ndw == 1) ? Cin : 3 * Cin;
        for (int t = 0; t < ntap; ++t) {
          mbar_wait(emptyB + 8 * sb, phb);
          if (PAIR) {
            if (leader) mbar_expect_tx(fullB + 8 * sb, (uint32_t)(2 * p.b_bytes));
            tma_load_2d_pair(b_buf + sb * p.b_bytes, &tmW, fullB_l + 8 * sb, k0 + t * kstep, p.wrow0 + brow);
          } else if ((p.dbg & 1) && b_cnt >= p.NB) mbar_arrive(fullB + 8 * sb);
          else {
            int kk = p.gemm ? k0 : k0 + t * kstep;
            if (p.mode == 1)        // tap (sa, sb) of parity (ph, pw) is kernel element (2 sa + 1 - ph, 2 sb + 1 - pw)
              kk = ((2 * (t >> 1) + 1 - (par >> 1)) * 4 + 2 * (t & 1) + 1 - (par & 1)) * Cin + ch * KCH;
            else if (p.mode == 2)
              kk = t * Cin + ch * KCH;
            mbar_expect_tx(fullB + 8 * sb, (uint32_t)p.b_bytes);
            tma_load_2d(b_buf + sb * p.b_bytes, &tmW, fullB + 8 * sb, kk, p.gemm ? p.wrow0 + t * N : p.wrow0);
          }
          ++b_cnt;
          if (++sb == p.NB) { sb = 0; phb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------- MMA issuer ---------------------------------------------
    // The whole warp walks the pipeline (warp-uniform control flow, no divisions in the loop); one elected lane
    // issues.  Descriptors are advanced by integer adds on the 16-byte-unit start-address field: +2 per K=8 step
    // (32 B), +1024 per 128-row sub-tile, +8 per dw, +8*pitch per dh.
    // instruction descriptor: D=f32 (1<<4), A=B=tf32 (2<<7, 2<<10), K-major both, N>>3 at bit 17, M>>4 at bit 24
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)((PAIR ? 256 : 128) >> 4) << 24);
    const uint64_t adesc_buf0 = umma_desc(a_buf);
    const uint64_t bdesc_buf0 = umma_desc(b_buf);
    const uint32_t a_lo0 = (uint32_t)adesc_buf0, a_hi = (uint32_t)(adesc_buf0 >> 32);
    const uint32_t b_lo0 = (uint32_t)bdesc_buf0, b_hi = (uint32_t)(bdesc_buf0 >> 32);
    const uint32_t a_step = (uint32_t)(p.a_bytes >> 4), b_step = (uint32_t)(p.b_bytes >> 4);
    const uint32_t dh_step = p.gemm ? 0u : (uint32_t)(p.pitch * (ROW_BYTES / 16));
    const int ndw_in = p.mode ? 2 : p.gemm ? 1 : ((p.ndw == 1) ? 3 : 1);   // dw taps served from one A box
    const int ndh_in = p.mode ? 2 : p.gemm ? p.ncol : 3;      // gemm: the "dh" loop walks the column tiles (A does not move)
    const uint32_t col_step = p.gemm ? (uint32_t)(p.S * N) : 0u;
    const int ncolt = quad ? Q : p.gemm ? p.ncol : 1;
    const int acc_stride = quad ? Q * 64 : N;               // TMEM columns between the accumulators of consecutive sub-tiles
    int sa = 0, sb = 0, ab = 0;
    uint32_t pha = 0, phb = 0, phacc = 1;
    for (int it = it_begin; it < it_end; it += it_step) {
      const int tile = tile_of(it);
      const int tf = tile % p.tiles_f;
      const int mu_tile = tf * p.S * 128;
      const int mu0 = mu_tile - (mu_tile / p.pitch) * p.pitch;
      const int npos = p.H * p.pitch - mu_tile;
      const int nsub = (npos >= p.S * 128) ? p.S : (npos + 127) / 128;
      mbar_wait(acc_empty + 8 * ab, phacc);               // epilogue has drained this accumulator set
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + (uint32_t)(ab * p.S * N * ncolt);
      uint32_t first = 0;
      for (int j = 0; j < nblk; ++j) {
        // quad: accumulators of output frames f0..f0+3 sit at columns 192, 128, 64, 0 of each 256-column sub-tile set (rows of
        // the stacked weights run dt = 0, 1, 2, i.e. towards EARLIER output frames); input frame index qi spans n columns
        uint32_t idesc_j = idesc, qcol = 0;
        if (quad) {
          const int qi = j / nch;
          const int dt_lo = qi - Q + 1 > 0 ? qi - Q + 1 : 0, dt_hi = qi < 2 ? qi : 2;
          const int n = 64 * (dt_hi - dt_lo + 1);
          qcol = (uint32_t)((Q - 1 - qi + dt_lo) * 64);      // accumulator of output frame i sits at column (Q-1-i)*64
          idesc_j = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(n >> 3) << 17);
        }
        mbar_wait(fullA + 8 * sa, pha);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t adesc_dh = a_lo0 + (uint32_t)(sa * a_step + mu0 * (ROW_BYTES / 16));
#pragma unroll 1
        for (int dh = 0; dh < ndh_in; ++dh, adesc_dh += dh_step) {
          uint32_t adesc = adesc_dh;
          const uint32_t tcol = tacc + (uint32_t)dh * col_step + qcol;
          if (p.gemm) first = (j > 0) ? 1u : 0u;
#pragma unroll 1
          for (int dw = 0; dw < ndw_in; ++dw, adesc += ROW_BYTES / 16) {
            mbar_wait(fullB + 8 * sb, phb);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t bdesc = b_lo0 + (uint32_t)(sb * b_step);
            if (elect_one()) {
#pragma unroll
              for (int s = 0; s < MAXS; ++s) {
                if (s < nsub) {
#pragma unroll
                  for (int k = 0; k < KCH / 8; ++k)
                    umma_tf32_w<PAIR>(tcol + (uint32_t)(s * acc_stride), adesc + (uint32_t)(s * (128 * ROW_BYTES / 16) + 2 * k), a_hi,
                                      bdesc + (uint32_t)(2 * k), b_hi, idesc_j, quad ? 1u : (first | (uint32_t)k));
                }
              }
              if (PAIR) umma_commit_pair(emptyB + 8 * sb);
              else umma_commit(emptyB + 8 * sb);          // weight slot is free once these MMAs retire
            }
            __syncwarp();
            first = 1;
            if (++sb == p.NB) { sb = 0; phb ^= 1; }
          }
        }
        if (elect_one()) { if (PAIR) umma_commit_pair(emptyA + 8 * sa); else umma_commit(emptyA + 8 * sa); }
        __syncwarp();
        if (++sa == NA) { sa = 0; pha ^= 1; }
      }
      if (elect_one()) { if (PAIR) umma_commit_pair(acc_full + 8 * ab); else umma_commit(acc_full + 8 * ab); }   // tile complete
      __syncwarp();
      if (++ab == AB) { ab = 0; phacc ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------- epilogue -----------------------------------------------
    const int q = warp & 3;                               // TMEM lane quarter == warp_id % 4
    const int eg = (warp - 4) >> 2;                       // epilogue group 0/1: sub-tiles s = eg, eg+2
    const bool do_stats = p.gn_stats != nullptr;
    // GroupNorm(8): group width N/8 columns, GPC groups per 32-column chunk; one (sum, sumsq) pair per group
    constexpr int GPC = 256 / N;                          // 4, 2, 1 for N = 64, 128, 256
    constexpr int GW = 32 / GPC;                          // columns of one group inside a chunk
    int ab = 0;
    uint32_t phacc = 0;
    const uint32_t acc_empty_l = PAIR ? leader_addr(acc_empty) : acc_empty;
    for (int it = it_begin; it < it_end; it += it_step) {
      const int tile = tile_of(it);
      const int tf = tile % p.tiles_f;
      const int f = ((tile / p.tiles_f) % Fd) * Q;   // quad: first output frame of the tile
      const int b = tile / (p.tiles_f * Fd);
      const int mu_tile = tf * p.S * 128;
      const int npos = p.H * p.pitch - mu_tile;
      const int nsub = (npos >= p.S * 128) ? p.S : (npos + 127) / 128;
      float fs[8], fq[8];                                 // per-thread partial sums (<= 64 values each) in fp32
#pragma unroll
      for (int g = 0; g < 8; ++g) fs[g] = fq[g] = 0.f;
      mbar_wait(acc_full + 8 * ab, phacc);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * p.S * N * (quad ? Q : p.gemm ? p.ncol : 1));
      if (p.gemm) {
        // Linear / 1x1x1 epilogue: the tile is store-bound, so rows are staged through a per-warp shared-memory tile
        // (32 rows x 36 floats, conflict-free float4 both ways) and written as 128-byte row segments: a warp store
        // covers 4 rows x 128 B instead of 32 rows x 16 B.
        float* stage = stage_all + (warp - 4) * (32 * 36);
        const size_t frame_base = ((size_t)b * p.F + f) * (size_t)(p.H * p.W);
        const int rsub = lane >> 3, cq = (lane & 7) * 4;
        // work items = 32-column chunks of every (column tile, sub-tile) accumulator, dealt alternately to the two groups;
        // the residual of the NEXT item is requested before this item is written (one round trip of latency per item
        // instead of one per 4-row group)
        const int cpa = N / 32;
        const int nitems = nsub * p.ncol * cpa;
        float4 rnext[8];
        auto item_geo = [&](int i, int& sN, int& ccol, int& mu_w) {
          const int acc = i / cpa, c = i - acc * cpa;     // accumulator (ct, s) = ct * S + s
          const int ct = acc / nsub, ss = acc - ct * nsub;
          sN = (ct * p.S + ss) * N + c * 32;
          ccol = p.wrow0 + ct * N + c * 32;
          mu_w = mu_tile + ss * 128 + q * 32;
        };
        auto load_res = [&](int i) {
          int sN, ccol, mu_w;
          item_geo(i, sN, ccol, mu_w);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int mu = mu_w + it * 4 + rsub;
            rnext[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (mu < p.H * p.W) rnext[it] = __ldcs(reinterpret_cast<const float4*>(p.residual + (frame_base + mu) * (size_t)p.ldy + ccol + cq));
          }
          if (p.res_scale) {                               // GroupNorm-apply + SiLU of the residual branch, folded per (b, channel)
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(p.res_scale + (size_t)b * p.ldy + ccol + cq));
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(p.res_shift + (size_t)b * p.ldy + ccol + cq));
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              float t0 = fmaf(rnext[it].x, a4.x, d4.x), t1 = fmaf(rnext[it].y, a4.y, d4.y);
              float t2 = fmaf(rnext[it].z, a4.z, d4.z), t3 = fmaf(rnext[it].w, a4.w, d4.w);
              rnext[it] = make_float4(__fdividef(t0, 1.0f + __expf(-t0)), __fdividef(t1, 1.0f + __expf(-t1)),
                                      __fdividef(t2, 1.0f + __expf(-t2)), __fdividef(t3, 1.0f + __expf(-t3)));
            }
          }
        };
        if (p.residual && eg < nitems) load_res(eg);
        for (int i = eg; i < nitems; i += 2) {
          int sN, ccol, mu_w;
          item_geo(i, sN, ccol, mu_w);
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)sN, v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const float* bp = p.bias ? p.bias + (ccol - p.wrow0) : nullptr;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bp) bv = __ldg(reinterpret_cast<const float4*>(bp + j));
            *reinterpret_cast<float4*>(stage + lane * 36 + j) =
                make_float4(__uint_as_float(v[j]) + bv.x, __uint_as_float(v[j + 1]) + bv.y,
                            __uint_as_float(v[j + 2]) + bv.z, __uint_as_float(v[j + 3]) + bv.w);
          }
          __syncwarp();
          float4 ov[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            ov[it] = *reinterpret_cast<const float4*>(stage + (it * 4 + rsub) * 36 + cq);
            if (p.residual) { ov[it].x += rnext[it].x; ov[it].y += rnext[it].y; ov[it].z += rnext[it].z; ov[it].w += rnext[it].w; }
          }
          if (p.residual && i + 2 < nitems) load_res(i + 2);
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int mu = mu_w + it * 4 + rsub;
            if (mu < p.H * p.W) __stcs(reinterpret_cast<float4*>(p.y + (frame_base + mu) * (size_t)p.ldy + ccol + cq), ov[it]);
          }
          __syncwarp();
        }
      } else
      for (int item = eg; item < Q * nsub; item += 2) {
        // quad: item = (sub-tile, frame of the tile); its accumulator sits at columns (s*Q + Q-1-i)*64
        const int s = item / Q;
        const int qfr = item - s * Q;
        const uint32_t acol = quad ? (uint32_t)((s * Q + Q - 1 - qfr) * 64) : (uint32_t)(s * N);
        const int mu_w = mu_tile + s * 128 + q * 32;        // first padded-flat position of this warp's 32 rows
        const size_t bf = (size_t)b * p.F + f + qfr;
        auto out_row = [&](int mu, bool& ok) -> size_t {     // output row index of position mu
          const int h = mu / p.pitch, w = mu - h * p.pitch;
          ok = (h < p.H) && (w < p.W);
          return p.mode == 2 ? (bf * p.Hout + (size_t)(h * p.o_mul + p.cls_h)) * p.Wout + (w * p.o_mul + p.cls_w)
                             : (bf * p.H + h) * p.W + w;
        };
        bool valid;
        const size_t m = out_row(mu_w + lane, valid);
        // staged stores: the warp's 32 rows x 32 columns go through a shared-memory tile and leave as 128-byte row segments
        // (a warp store covers 4 rows x 128 B instead of 32 rows x 16 B: 8x fewer L1 transactions; the direct form costs
        // ~12K clk per 384 x 128 tile, which is exposed whenever the accumulators are single-buffered)
        float* stage = stage_all + (warp - 4) * (32 * 36);
        const int rsub = lane >> 3, cq = (lane & 7) * 4;
        uint32_t roff[8];
        uint32_t vmask = 0;
        if (p.staged) {
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            bool ok;
            const size_t mr = out_row(mu_w + it * 4 + rsub, ok);
            roff[it] = (uint32_t)(mr * (size_t)p.ldy) + (uint32_t)(p.wrow0 + cq);
            vmask |= ok ? (1u << it) : 0u;
          }
        }
        float* dst = p.y + m * p.ldy + p.wrow0;
#pragma unroll
        for (int c = 0; c < N / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tacc + acol + (uint32_t)(c * 32), v);
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
          if (quad) {                                      // leave the accumulator at zero for the next tile
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = 0u;
            tmem_st32(tacc + acol + (uint32_t)(c * 32), v);
          }
          if (p.staged) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(stage + lane * 36 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 8; ++it)
              if ((vmask >> it) & 1u)
                __stcs(reinterpret_cast<float4*>(p.y + roff[it] + c * 32),
                       *reinterpret_cast<const float4*>(stage + (it * 4 + rsub) * 36 + cq));
            __syncwarp();
          } else if (valid) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              __stcs(reinterpret_cast<float4*>(dst + c * 32 + j), make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]));
          }
          if (valid && do_stats) {
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              float ps = 0.f, pq = 0.f;
#pragma unroll
              for (int j = 0; j < GW; ++j) {
                ps += o[g * GW + j];
                pq = fmaf(o[g * GW + j], o[g * GW + j], pq);
              }
              fs[(c * GPC + g) % 8] += ps;              // (c*GPC + g) < 8 by construction
              fq[(c * GPC + g) % 8] += pq;
            }
          }
        }
      }
      // accumulator set drained: hand it back to the MMA warp before the (slower) statistics reduction
      if (quad) tmem_wait_st();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(acc_empty_l + 8 * ab);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty + 8 * ab) : "memory");
      }
      if (++ab == AB) { ab = 0; phacc ^= 1; }
      if (do_stats) {
        // recursive-halving warp reduction of 16 doubles (8 sums, 8 sums of squares): 16 exchanges instead of 80
        double v[16];
#pragma unroll
        for (int g = 0; g < 8; ++g) { v[g] = (double)fs[g]; v[8 + g] = (double)fq[g]; }
#pragma unroll
        for (int half = 8, mask = 16; half >= 1; half >>= 1, mask >>= 1) {
          const bool up = (lane & mask) != 0;
#pragma unroll
          for (int i = 0; i < half; ++i) {
            const double keep = up ? v[i + half] : v[i];
            const double send = up ? v[i] : v[i + half];
            v[i] = keep + shfl_xor_double(send, mask);
          }
        }
        v[0] += shfl_xor_double(v[0], 1);                 // lane L now holds the warp total of value index L >> 1
        asm volatile("bar.sync 1, 256;" ::: "memory");    // previous tile's s_part has been consumed
        if ((lane & 1) == 0) s_part[warp - 4][lane >> 1] = v[0];
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int et = threadIdx.x - 128;
        if (et < 16) {
          const int slot = et;   // after the halving lane L holds value index L >> 1 (bit k of L selects bit k-1)
          const double tot = ((s_part[0][et] + s_part[1][et]) + (s_part[2][et] + s_part[3][et])) +
                             ((s_part[4][et] + s_part[5][et]) + (s_part[6][et] + s_part[7][et]));
          // local slot = 32/GPC... columns [grp*N/8, (grp+1)*N/8) of this launch's column slice -> group of the whole layer
          const int which = slot >> 3, grp = (p.wrow0 + (slot & 7) * (N / 8)) / p.gcols;
          atomicAdd(p.gn_stats + ((size_t)b * 8 + grp) * 2 + which, tot);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (PAIR) cluster_sync_all();                          // the peer may still signal this CTA's barriers / read its operands
  if (warp == 2) {
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
static int make_act_map(CUtensorMap* m, const float* x, int B, int F, int H, int W, int C, int box_w, int box_h, int estride = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_err(-1, "cuTensorMapEncodeTiled unavailable", __FILE__, __LINE__);
  cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)F, (cuuint64_t)B};
  cuuint64_t strides[4] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)F * H * W * C * 4};
  // estride 2: every other pixel in w and h (the box extents are in traversed elements; the tile lands compacted)
  cuuint32_t box[5] = {(cuuint32_t)KCH, (cuuint32_t)(box_w * estride), (cuuint32_t)(box_h * estride), 1, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1, 1};
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

template <int N, bool PAIR>
static int launch(const CUtensorMap& a1, const CUtensorMap& a2, const CUtensorMap& a1t, const CUtensorMap& a2t,
                  const CUtensorMap& wm, const CUtensorMap& wm3, const Params& p, size_t smem, cudaStream_t st) {
  const int dev = device_ordinal();
  static size_t configured_[kMaxDevices] = {};
  size_t& configured = configured_[dev];
  if (smem > configured) {
    DPC_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<N, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (PAIR) DPC_CUDA(cudaFuncSetAttribute(conv3d_tc_kernel<N, PAIR>, cudaFuncAttributeNonPortableClusterSizeAllowed, 0));
    configured = smem;
  }
  const size_t ntiles = (size_t)p.B * (p.quad ? p.F / p.quad : p.F) * p.tiles_f;
  const int num_sms = sm_count(dev);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(NTHREADS_TC);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (PAIR) {
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // persistent CTA pairs: as many clusters as can be co-resident (an SM pair of one TPC each)
    static int max_clusters_[kMaxDevices] = {};
    static size_t clusters_smem_[kMaxDevices] = {};
    int& max_clusters = max_clusters_[dev];
    size_t& clusters_smem = clusters_smem_[dev];
    if (!max_clusters || smem > clusters_smem) {
      cfg.gridDim = dim3((unsigned)(num_sms & ~1));
      int n = 0;
      DPC_CUDA(cudaOccupancyMaxActiveClusters(&n, conv3d_tc_kernel<N, PAIR>, &cfg));
      if (n < 1) return set_err(-1, "no co-resident CTA pair for conv3d_tc_kernel", __FILE__, __LINE__);
      max_clusters = n;
      clusters_smem = smem;
    }
    size_t pairs = ntiles / 2;
    if (pairs > (size_t)max_clusters) pairs = (size_t)max_clusters;
    cfg.gridDim = dim3((unsigned)(2 * pairs));
  } else {
    cfg.gridDim = dim3((unsigned)(ntiles < (size_t)num_sms ? ntiles : (size_t)num_sms));   // persistent: one CTA per SM
  }
  DPC_CUDA(cudaLaunchKernelEx(&cfg, conv3d_tc_kernel<N, PAIR>, a1, a2, a1t, a2t, wm, wm3, p));
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
  const bool common_ok = c.st == 1 && c.sh == 1 && c.sw == 1 && c.Fo == F && c.Ho == H && c.Wo == W && c.oh_mul == 1 &&
                         c.ow_mul == 1 && c.Hfull == H && c.Wfull == W && c.out_layout == 0 && c.precise == 0 &&
                         c.C1 % KCH == 0 && c.C2 % KCH == 0 && c.C1 > 0 && W >= 4 && W + 2 <= 256 && H >= 1 &&
                         c.Kpad == c.ntaps * (c.C1 + c.C2);
  // 3x3x3 over frames (pt = 1) or 3x3 over images (ntaps = 9, pt = 0: the 2-D networks of diffusion_2d_jellyfish.py:189-204);
  // Cout = 512 runs as two 256-column launches
  const bool conv_ok = common_ok && ((c.ntaps == 27 && c.pt == 1) || (c.ntaps == 9 && c.pt == 0)) && c.ph == 1 && c.pw == 1 &&
                       c.residual == nullptr && (c.Cout == 64 || c.Cout == 128 || c.Cout == 256 || c.Cout == 512) &&
                       c.Npad == c.Cout;
  // 1x1x1 conv / Linear: N tiles of 64 / 128 / 256 columns (e.g. the 384-wide qkv projection = 3 x 128)
  const bool gemm_ok = common_ok && c.ntaps == 1 && c.pt == 0 && c.ph == 0 && c.pw == 0 && c.gn_stats == nullptr &&
                       (c.Cout == 64 || c.Cout % 128 == 0) && c.Npad >= c.Cout;
  // 1x4x4 / stride (1,2,2) / pad (0,1,1) down-conv (conv3d.py:163) and one parity class of the ConvTranspose (conv3d.py:160)
  const bool two_ok = c.st == 1 && c.pt == 0 && c.Fo == F && c.out_layout == 0 && c.precise == 0 && c.C1 % KCH == 0 && c.C1 > 0 &&
                      c.C2 == 0 && c.residual == nullptr && c.gn_stats == nullptr &&
                      (c.Cout == 64 || c.Cout == 128 || c.Cout == 256) && c.Npad == c.Cout && c.Kpad == c.ntaps * c.C1;
  const bool down_ok = two_ok && c.ntaps == 16 && c.sh == 2 && c.sw == 2 && c.ph == 1 && c.pw == 1 && c.Ho * 2 == H &&
                       c.Wo * 2 == W && c.oh_mul == 1 && c.ow_mul == 1 && c.Hfull == c.Ho && c.Wfull == c.Wo && c.Wo >= 4 &&
                       2 * (c.Wo + 1) <= 256;
  const bool up_ok = two_ok && c.ntaps == 4 && c.sh == 1 && c.sw == 1 && c.Ho == H && c.Wo == W && c.oh_mul == 2 && c.ow_mul == 2 &&
                     c.Hfull == 2 * H && c.Wfull == 2 * W && (c.oh_off == 0 || c.oh_off == 1) && (c.ow_off == 0 || c.ow_off == 1) &&
                     c.ph == 1 - c.oh_off && c.pw == 1 - c.ow_off && W >= 4 && W + 1 <= 256;
  if (!conv_ok && !gemm_ok && !down_ok && !up_ok) return -2;
  if ((c.res_scale || c.res_shift) && !(gemm_ok && !conv_ok && c.residual && c.res_scale && c.res_shift)) return -2;
  if (c.gn_stats && c.gn_groups != 8) return -2;   // the epilogue is specialised for GroupNorm(8), the reference default
  DPC_CHECK_ARG(c.x1 && c.w && c.y && (c.C2 == 0 || c.x2));
  const bool gemm = gemm_ok && !conv_ok;
  const int mode = conv_ok || gemm_ok ? 0 : down_ok ? 1 : 2;
  // cta_group::2 pairs (two consecutive frames per cluster) for the 3x3x3 convolutions; DPC_TC_PAIR=0 disables
  const char* pair_env = getenv("DPC_TC_PAIR");          // read per call: the tests run every shape both ways
  const bool pair = !gemm && mode == 0 && !(pair_env && atoi(pair_env) == 0) && ((int64_t)c.B * F) % 2 == 0;
  const int Ntile = gemm ? (c.Cout == 64 ? 64 : (c.Cout == 256 ? 256 : 128)) : (c.Cout > 256 ? 256 : c.Cout);
  Params p;
  p.bias = c.bias; p.residual = c.residual; p.res_scale = c.res_scale; p.res_shift = c.res_shift; p.y = c.y; p.gn_stats = c.gn_stats; p.gn_groups = c.gn_groups;
  p.B = c.B; p.F = F; p.H = mode == 1 ? c.Ho : H; p.W = mode == 1 ? c.Wo : W; p.C1 = c.C1; p.C2 = c.C2; p.Cout = c.Cout;
  p.mode = mode; p.cls_h = c.oh_off; p.cls_w = c.ow_off; p.Hout = c.Hfull; p.Wout = c.Wfull; p.o_mul = c.oh_mul;
  p.gemm = gemm ? 1 : 0;
  p.nkt = (conv_ok && c.ntaps == 9) ? 1 : 3;
  p.ptt = (conv_ok && c.ntaps == 9) ? 0 : 1;
  p.gcols = c.Cout / 8;
  { static int dbg = -1; if (dbg < 0) { const char* e = getenv("DPC_TC_DEBUG"); dbg = e ? atoi(e) : 0; } p.dbg = dbg; }
  p.ldy = c.Cout;
  p.wrow0 = 0;
  // small frames: padding columns (W+2)/W and the 128-row quantisation of the padded-flat domain waste too much of
  // the tensor pipe, so fall back to one box per dw (pitch W, three times the A traffic, zero wasted rows)
  p.ndw = (mode || (!gemm && W >= 32)) ? 1 : 3;
  p.pitch = mode ? p.W + 1 : (p.ndw == 1) ? W + 2 : W;   // 2x2 taps: one halo column
  const int frame_pos = p.H * p.pitch;
  // quad (Cout = 64, large frames): four output frames per tile, temporal taps stacked into N = 64/128/192 MMAs; needs whole
  // quads per sample and an even number of quads for the CTA pairs.  Measured neutral (64->64 @64^2: 2.78 vs 2.85 ms,
  // 128->64: 5.6 vs 6.1 ms, 32^2 layers slightly slower: single-buffered accumulators, per-column MMA cost does not drop
  // with N), so it is opt-in: DPC_TC_QUAD=1.
  const char* quad_env = getenv("DPC_TC_QUAD");
  const int Qreq = quad_env ? atoi(quad_env) : 0;      // 1 or 4: four frames per tile; 2: two frames per tile
  const int Qn = Qreq == 2 ? 2 : 4;
  const bool quad = pair && p.nkt == 3 && c.Cout == 64 && p.ndw == 1 && F % Qn == 0 && ((int64_t)c.B * (F / Qn)) % 2 == 0 &&
                    (Qreq == 1 || Qreq == 2 || Qreq == 4);
  p.quad = quad ? Qn : 0;
  // gemm: up to 512 accumulator columns = ncol column tiles side by side, so A is read once for (up to) 512 outputs
  int ncol = 1;
  if (gemm) { ncol = c.Cout / Ntile; if (ncol * Ntile > 512) ncol = 512 / Ntile; }
  p.ncol = ncol;
  int S = gemm ? 256 / (ncol * Ntile) : quad ? 2 : 512 / Ntile;   // gemm tiles are store-bound: keep two accumulator sets when they fit
  if (S < 1) S = 1;
  if (S > MAXS) S = MAXS;
  if (S > (frame_pos + 127) / 128) S = (frame_pos + 127) / 128;
  const size_t budget = 227 * 1024 - 2048;
  const size_t stage_full = (size_t)8 * 32 * 36 * sizeof(float);
  size_t stage_bytes = gemm ? stage_full : 0;
  int NA = 2;
  p.b_bytes = (quad ? (Qn == 2 ? 64 : 96) : pair ? Ntile / 2 : Ntile) * ROW_BYTES;    // a pair splits every weight box between its two CTAs
  for (;; --S) {
    // rows needed: offset inside the first row (< pitch) + S*128 positions (+ two more image rows + 2 positions of halo)
    const int span = p.pitch - 1 + S * 128 + (gemm ? 0 : mode ? p.pitch + 1 : 2 * p.pitch + 2);
    p.R = (span + p.pitch - 1) / p.pitch;
    p.a_bytes = ((p.R * p.pitch * ROW_BYTES + 1023) / 1024) * 1024;
    if ((size_t)NA * p.a_bytes + 3 * (size_t)p.b_bytes + 1024 + 256 + stage_bytes <= budget || S == 1) break;
  }
  if ((size_t)NA * p.a_bytes + 3 * (size_t)p.b_bytes + 1024 + 256 + stage_bytes > budget || p.R > 256) return -2;
  // balance the sub-tiles over the tiles of a frame (e.g. 32x34 positions = 9 sub-tiles: 3+3+3 instead of 4+4+1 — a tile
  // with one sub-tile still streams every weight box)
  if (!gemm) {
    const int nsub_total = (frame_pos + 127) / 128, nt = (nsub_total + S - 1) / S;
    const int Sb = (nsub_total + nt - 1) / nt;
    if (Sb < S) {
      S = Sb;
      const int span = p.pitch - 1 + S * 128 + (mode ? p.pitch + 1 : 2 * p.pitch + 2);
      p.R = (span + p.pitch - 1) / p.pitch;
      p.a_bytes = ((p.R * p.pitch * ROW_BYTES + 1023) / 1024) * 1024;
    }
  }
  // conv epilogue staging tiles when they leave >= 4 weight slots (and 32-bit element offsets suffice)
  p.staged = 0;
  if (!gemm && (size_t)NA * p.a_bytes + 4 * (size_t)p.b_bytes + 1024 + 256 + stage_full <= budget &&
      (int64_t)c.B * F * c.Hfull * c.Wfull * c.Cout < ((int64_t)1 << 32)) {
    p.staged = 1;
    stage_bytes = stage_full;
  }
  // a third A slot (two boxes in flight behind the one being multiplied) whenever it leaves >= 4 weight slots
  if (!gemm && (size_t)3 * p.a_bytes + 4 * (size_t)p.b_bytes + 1024 + 256 + stage_bytes <= budget) NA = 3;
  p.NA = NA;
  p.S = S;
  p.AB = quad ? (2 * S * Qn * 64 <= 512 ? 2 : 1) : (2 * S * ncol * Ntile <= 512) ? 2 : 1;
  { int need = quad ? 512 : p.AB * S * ncol * Ntile; p.tmem_cols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512; }
  p.tiles_f = (frame_pos + S * 128 - 1) / (S * 128);
  int NB = (int)((budget - 1024 - 256 - stage_bytes - (size_t)NA * p.a_bytes) / p.b_bytes);
  if (NB > 9) NB = 9;
  if (NB < 2) return -2;
  p.NB = NB;
  const size_t smem = (size_t)NA * p.a_bytes + (size_t)NB * p.b_bytes + 1024 + 256 + stage_bytes;
  // the A box is fetched as APARTS slabs of rows_part rows; the last slab may be shorter -> its own tensor map
  const int rows_part = (p.R + APARTS - 1) / APARTS;
  const int nfull = p.R / rows_part;
  const int tail_rows = p.R - nfull * rows_part;
  CUtensorMap a1, a2, a1t, a2t, wm;
  const int estride = mode == 1 ? 2 : 1;
  int rc = make_act_map(&a1, c.x1, c.B, F, H, W, c.C1, p.pitch, rows_part, estride);
  if (rc) return rc;
  rc = make_act_map(&a1t, c.x1, c.B, F, H, W, c.C1, p.pitch, tail_rows > 0 ? tail_rows : rows_part, estride);
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
  rc = make_w_map(&wm, c.w, c.Kpad, c.Npad, pair ? Ntile / 2 : Ntile);
  if (rc) return rc;
  CUtensorMap wm3 = wm;
  if (quad) {
    // 3-D view of the packed weights [64][27*Cin]: (k' = (dh*3+dw)*Cin + ci, n, dt), boxes of 32 rows of one dt
    EncodeTiledFn enc = get_encode();
    const int Cin = c.C1 + c.C2;
    cuuint64_t dims[3] = {(cuuint64_t)(9 * Cin), 64, 3};
    cuuint64_t strides[2] = {(cuuint64_t)c.Kpad * 4, (cuuint64_t)(9 * Cin) * 4};
    cuuint32_t box[3] = {(cuuint32_t)KCH, 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&wm3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)c.w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_err(-1, "cuTensorMapEncodeTiled(stacked weights) failed", __FILE__, (int)r);
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int n0 = 0; n0 < c.Cout; n0 += ncol * Ntile) {
    if (gemm && c.Cout - n0 < ncol * Ntile) p.ncol = (c.Cout - n0) / Ntile;
    p.wrow0 = n0;
    p.bias = c.bias ? c.bias + n0 : nullptr;
    if (pair) {
      if (Ntile == 64) rc = launch<64, true>(a1, a2, a1t, a2t, wm, wm3, p, smem, st);
      else if (Ntile == 128) rc = launch<128, true>(a1, a2, a1t, a2t, wm, wm3, p, smem, st);
      else rc = launch<256, true>(a1, a2, a1t, a2t, wm, wm3, p, smem, st);
    } else if (Ntile == 64) rc = launch<64, false>(a1, a2, a1t, a2t, wm, wm3, p, smem, st);
    else if (Ntile == 128) rc = launch<128, false>(a1, a2, a1t, a2t, wm, wm3, p, smem, st);
    else rc = launch<256, false>(a1, a2, a1t, a2t, wm, wm3, p, smem, st);
    if (rc) return rc;
  }
  return 0;
}
