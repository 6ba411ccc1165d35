// Attention kernels of Unet3D_with_Conv3D (dim_head = 32, fp32 SIMT arithmetic like the reference's fp32 bmm/einsum):
//   dpc_temporal_attention        conv3d.py:293-352  (+ RoPE of rotary-embedding-torch 0.8.4, T5 relative bias :74-112)
//   dpc_spatial_attention         conv3d.py:449-451 + :293-352 (mid block, tokens = pixels of a frame)
//   dpc_spatial_linear_attention  conv3d.py:243-257
// qkv rows are channels-last [rows][3*heads*32]: q | k | v thirds, head-major inside a third.
#include "common.cuh"

namespace dpc {

constexpr int DH = 32;
constexpr float ATT_SCALE = 0.17677669529663687f;  // 32 ** -0.5 (conv3d.py:286)

// ------------------------------------------------------------------------------------------------------------
// temporal attention, F <= 64, tensor-core version: one warp per (sample, pixel, head); 32 queries at a time against all keys.
//   stage : coalesced 128-byte row segments of q|k|v -> scale, RoPE -> shared memory (row stride 36 floats)
//   S     : Q K^T as 2x4 m16n8k8 TF32 MMAs per k-step (ldmatrix fragments), + relative bias, key mask, row softmax
//           (row reductions stay inside a quad: 2 shuffles)
//   O     : P V with the accumulator fragments of S reused directly as the A operand: lane (g,t) holds P[row][2t],
//           P[row][2t+1] of every 8-key block, so the MMA's k index is mapped to keys (2t, 2t+1) and the V fragment is
//           read with the same permutation — no shuffles, no round trip through shared memory
//   store : O -> shared memory -> coalesced 128-byte row segments
// PRECISE = 3xTF32 error-compensated products (fp32-class, the reference's einsum precision); otherwise operands are
// rounded to TF32 with cvt.rna.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool PRECISE>
__device__ __forceinline__ void mma_op(float (&d)[4], const float (&a)[4], float b0, float b1) {
  if constexpr (!PRECISE) {
    const uint32_t au[4] = {to_tf32(a[0]), to_tf32(a[1]), to_tf32(a[2]), to_tf32(a[3])};
    mma_tf32_16x8x8(d, au, to_tf32(b0), to_tf32(b1));
  } else {
    uint32_t ab[4], as[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ab[i] = __float_as_uint(a[i]) & 0xffffe000u;
      as[i] = __float_as_uint(a[i] - __uint_as_float(ab[i]));
    }
    const uint32_t b0b = __float_as_uint(b0) & 0xffffe000u, b1b = __float_as_uint(b1) & 0xffffe000u;
    const uint32_t b0s = __float_as_uint(b0 - __uint_as_float(b0b)), b1s = __float_as_uint(b1 - __uint_as_float(b1b));
    mma_tf32_16x8x8(d, as, b0b, b1b);
    mma_tf32_16x8x8(d, ab, b0s, b1s);
    mma_tf32_16x8x8(d, ab, b0b, b1b);
  }
}

template <bool PRECISE, int NF>   // NF = 32 or 64: frames padded to a multiple of 32 (F <= NF)
__global__ void __launch_bounds__(128)
temporal_attention_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ rope_cos,
                              const float* __restrict__ rope_sin, const float* __restrict__ pos_bias,
                              float* __restrict__ out, int64_t total_warps, int F, int HW, int heads, int use_rope) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LD = DH + 4;
  constexpr int NKT = NF / 8;                         // key blocks of 8
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* Qs = sm + (size_t)warp * 3 * NF * LD;
  float* Ks = Qs + NF * LD;
  float* Vs = Ks + NF * LD;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const int g = lane >> 2, t = lane & 3;
  const int lf = lane >> 3, lc = (lane & 7) * 4;      // staging: frame-in-group, first channel of this lane's float4
  // flags bit 1: pos_bias[h][i][j] depends on j - i only (RelativePositionBias, conv3d.py:110-148): one [heads][2 NF] table in
  // shared memory instead of global loads that touch 8 cache lines per warp instruction
  const bool rel = pos_bias && (use_rope & 2);
  float* bias_s = sm + (size_t)4 * 3 * NF * LD;
  if (rel) {
    for (int i = threadIdx.x; i < heads * 2 * NF; i += 128) {
      const int h = i / (2 * NF), d = i - h * 2 * NF - (NF - 1);
      float v = 0.f;
      if (d < F && -d < F) v = (d >= 0) ? __ldg(pos_bias + ((size_t)h * F + 0) * F + d) : __ldg(pos_bias + ((size_t)h * F - d) * F + 0);
      bias_s[i] = v;
    }
    __syncthreads();
  }
  use_rope &= 1;

  for (int64_t wg = (int64_t)blockIdx.x * 4 + warp; wg < total_warps; wg += (int64_t)gridDim.x * 4) {
    const int head = (int)(wg % heads);
    const int64_t bp = wg / heads;
    const int pix = (int)(bp % HW);
    const int64_t b = bp / HW;
    // ---- stage q, k, v ----
#pragma unroll
    for (int it = 0; it < NF / 4; ++it) {
      const int f = it * 4 + lf;
      float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f), k4 = q4, v4 = q4;
      if (f < F) {
        const float* row = qkv + (((size_t)b * F + f) * HW + pix) * C3 + head * DH + lc;
        q4 = __ldcs(reinterpret_cast<const float4*>(row));
        k4 = __ldcs(reinterpret_cast<const float4*>(row + hid));
        v4 = __ldcs(reinterpret_cast<const float4*>(row + 2 * hid));
        q4.x = __fmul_rn(q4.x, ATT_SCALE); q4.y = __fmul_rn(q4.y, ATT_SCALE);
        q4.z = __fmul_rn(q4.z, ATT_SCALE); q4.w = __fmul_rn(q4.w, ATT_SCALE);
        if (use_rope) {
          const float4 cs = __ldg(reinterpret_cast<const float4*>(rope_cos + (size_t)f * DH + lc));
          const float4 sn = __ldg(reinterpret_cast<const float4*>(rope_sin + (size_t)f * DH + lc));
          // t*cos + rotate_half(t)*sin with rotate_half: (x0, x1) -> (-x1, x0), separately rounded like the reference
          float4 r;
          r.x = __fadd_rn(__fmul_rn(q4.x, cs.x), __fmul_rn(-q4.y, sn.x));
          r.y = __fadd_rn(__fmul_rn(q4.y, cs.y), __fmul_rn(q4.x, sn.y));
          r.z = __fadd_rn(__fmul_rn(q4.z, cs.z), __fmul_rn(-q4.w, sn.z));
          r.w = __fadd_rn(__fmul_rn(q4.w, cs.w), __fmul_rn(q4.z, sn.w));
          q4 = r;
          r.x = __fadd_rn(__fmul_rn(k4.x, cs.x), __fmul_rn(-k4.y, sn.x));
          r.y = __fadd_rn(__fmul_rn(k4.y, cs.y), __fmul_rn(k4.x, sn.y));
          r.z = __fadd_rn(__fmul_rn(k4.z, cs.z), __fmul_rn(-k4.w, sn.z));
          r.w = __fadd_rn(__fmul_rn(k4.w, cs.w), __fmul_rn(k4.z, sn.w));
          k4 = r;
        }
      }
      *reinterpret_cast<float4*>(Qs + f * LD + lc) = q4;
      *reinterpret_cast<float4*>(Ks + f * LD + lc) = k4;
      *reinterpret_cast<float4*>(Vs + f * LD + lc) = v4;
    }
    __syncwarp();
#pragma unroll 1
    for (int qh = 0; qh < NF / 32; ++qh) {             // 32 queries at a time (two m16 tiles) against all NF keys
      const float* Qh = Qs + qh * 32 * LD;
      // ---- S = Q K^T ----
      float s[2][NKT][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        float a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const float* qa = Qh + (mt * 16 + g) * LD + ks * 8 + t;
          a[mt][0] = qa[0]; a[mt][1] = qa[8 * LD]; a[mt][2] = qa[4]; a[mt][3] = qa[8 * LD + 4];
        }
#pragma unroll
        for (int nt = 0; nt < NKT; ++nt) {
          const float* kb = Ks + (nt * 8 + g) * LD + ks * 8 + t;   // B[k = d][n = key] = K[key][d]
          const float b0 = kb[0], b1 = kb[4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_op<PRECISE>(s[mt][nt], a[mt], b0, b1);
        }
      }
      // ---- bias, key mask, softmax over keys (row = qh*32 + mt*16 + g + 8*h, col = nt*8 + 2t + e) ----
      float inv[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int row = qh * 32 + mt * 16 + g + 8 * h;
          const float* brow = (!rel && pos_bias && row < F) ? pos_bias + ((size_t)head * F + row) * F : nullptr;
          const float* brel = bias_s + head * 2 * NF + (NF - 1) - row;
          float mx = -INFINITY;
#pragma unroll
          for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = nt * 8 + 2 * t + e;
              float v = s[mt][nt][2 * h + e];
              if (rel) v += brel[col];
              else if (brow && col < F) v += __ldg(brow + col);
              if (col >= F) v = -INFINITY;
              s[mt][nt][2 * h + e] = v;
              mx = fmaxf(mx, v);
            }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          float l = 0.f;
#pragma unroll
          for (int nt = 0; nt < NKT; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const float pv = PRECISE ? expf(s[mt][nt][2 * h + e] - mx) : __expf(s[mt][nt][2 * h + e] - mx);
              s[mt][nt][2 * h + e] = pv;
              l += pv;
            }
          l += __shfl_xor_sync(0xffffffffu, l, 1);
          l += __shfl_xor_sync(0xffffffffu, l, 2);
          inv[mt][h] = 1.0f / l;
        }
      // ---- O = P V (k index of the MMA mapped to keys 8*kb + 2t, 8*kb + 2t + 1) ----
      float o[2][4][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
#pragma unroll
          for (int e = 0; e < 4; ++e) o[mt][dn][e] = 0.f;
#pragma unroll
      for (int kb = 0; kb < NKT; ++kb) {
        float a[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          a[mt][0] = s[mt][kb][0] * inv[mt][0];   // (row g,   key 2t)
          a[mt][1] = s[mt][kb][2] * inv[mt][1];   // (row g+8, key 2t)
          a[mt][2] = s[mt][kb][1] * inv[mt][0];   // (row g,   key 2t+1)
          a[mt][3] = s[mt][kb][3] * inv[mt][1];   // (row g+8, key 2t+1)
        }
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          const float* vb = Vs + (kb * 8 + 2 * t) * LD + dn * 8 + g;
          const float b0 = vb[0], b1 = vb[LD];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) mma_op<PRECISE>(o[mt][dn], a[mt], b0, b1);
        }
      }
      __syncwarp();   // every lane is done reading this half of Qs: reuse it as the output staging tile
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          float* od = Qs + (qh * 32 + mt * 16 + g) * LD + dn * 8 + 2 * t;
          *reinterpret_cast<float2*>(od) = make_float2(o[mt][dn][0], o[mt][dn][1]);
          *reinterpret_cast<float2*>(od + 8 * LD) = make_float2(o[mt][dn][2], o[mt][dn][3]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < NF / 4; ++it) {
      const int f = it * 4 + lf;
      if (f < F) {
        const float4 v = *reinterpret_cast<const float4*>(Qs + f * LD + lc);
        __stcs(reinterpret_cast<float4*>(out + (((size_t)b * F + f) * HW + pix) * hid + head * DH + lc), v);
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------------------
// mid-block spatial softmax attention on tensor cores (TF32 mma.sync m16n8k8, flash-attention style): one CTA per
// (frame, head, block of 256 queries), a warp owns 32 queries (two m16 tiles); K/V stream through shared memory in blocks
// of 64 keys with an online softmax; the score accumulators are re-used as the A operand of P.V through the k-index
// permutation (keys 2t, 2t+1 <-> k positions t, t+4), so P never leaves the registers.  Shared rows have pitch 36 floats:
// both B-fragment access patterns (K: row g, column t; V: row 2t, column g) are bank-conflict free.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
spatial_attention_mma_kernel(const float* __restrict__ qkv, float* __restrict__ out, int HW, int heads) {
  constexpr int KB = 64, PITCH = 36;
  __shared__ __align__(16) float Ks[KB * PITCH];
  __shared__ __align__(16) float Vs[KB * PITCH];
  const int head = blockIdx.y;
  const int64_t bf = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const float* base = qkv + (size_t)bf * HW * C3 + head * DH;
  const int q0 = blockIdx.x * 256 + warp * 32;           // first query of this warp
  const bool active = q0 < HW;                           // HW % 32 == 0: a warp is entirely in or out
  // Q fragments, scaled, TF32: qa[mt][kk] = rows q0 + 16 mt + {g, g+8}, columns 8 kk + {t, t+4}
  uint32_t qa[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = q0 + 16 * mt + g + 8 * (e & 1), col = 8 * kk + t + 4 * (e >> 1);
        qa[mt][kk][e] = active ? to_tf32(__fmul_rn(__ldg(base + (size_t)row * C3 + col), ATT_SCALE)) : 0u;
      }
  float oc[2][4][4];
  float mx[2][2], l[2][2];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int dn = 0; dn < 4; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) oc[mt][dn][e] = 0.f;
    mx[mt][0] = mx[mt][1] = -INFINITY;
    l[mt][0] = l[mt][1] = 0.f;
  }
  for (int j0 = 0; j0 < HW; j0 += KB) {
    __syncthreads();
    // 64 keys x (32 k + 32 v) floats = 1024 float4, 4 per thread; TF32-rounded once here
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + 256 * i;
      const int tk = idx >> 4, part = idx & 15;
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + (size_t)(j0 + tk) * C3 + (part < 8 ? hid : 2 * hid) + (part & 7) * 4));
      float* dst = (part < 8 ? Ks : Vs) + tk * PITCH + (part & 7) * 4;
      *reinterpret_cast<float4*>(dst) = make_float4(__uint_as_float(to_tf32(v.x)), __uint_as_float(to_tf32(v.y)),
                                                    __uint_as_float(to_tf32(v.z)), __uint_as_float(to_tf32(v.w)));
    }
    __syncthreads();
    if (!active) continue;
    // ---- S = q K^T for 64 keys ----
    float sc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[mt][nt][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b0 = __float_as_uint(Ks[(8 * nt + g) * PITCH + 8 * kk + t]);
        const uint32_t b1 = __float_as_uint(Ks[(8 * nt + g) * PITCH + 8 * kk + t + 4]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_tf32_16x8x8(sc[mt][nt], qa[mt][kk], b0, b1);
      }
    // ---- online softmax: rows 16 mt + g (elements 0,1) and + 8 (elements 2,3) ----
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        float tm = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) tm = fmaxf(tm, fmaxf(sc[mt][nt][2 * e2], sc[mt][nt][2 * e2 + 1]));
        tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 1));
        tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, 2));
        const float nm = fmaxf(mx[mt][e2], tm);
        const float corr = __expf(mx[mt][e2] - nm);      // first block: exp(-inf) = 0
        mx[mt][e2] = nm;
        float ps = 0.f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          const float p0 = __expf(sc[mt][nt][2 * e2] - nm), p1 = __expf(sc[mt][nt][2 * e2 + 1] - nm);
          sc[mt][nt][2 * e2] = p0;
          sc[mt][nt][2 * e2 + 1] = p1;
          ps += p0 + p1;
        }
        l[mt][e2] = l[mt][e2] * corr + ps;               // this lane's share of the row sum (reduced over the quad at the end)
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
          oc[mt][dn][2 * e2] *= corr;
          oc[mt][dn][2 * e2 + 1] *= corr;
        }
      }
    // ---- O += P V: score fragments are the A operand (k positions t, t+4 <-> keys 8 kb + 2t, 2t+1) ----
#pragma unroll
    for (int kb = 0; kb < 8; ++kb) {
      uint32_t pa[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        pa[mt][0] = to_tf32(sc[mt][kb][0]);              // (row g,   key 2t)
        pa[mt][1] = to_tf32(sc[mt][kb][2]);              // (row g+8, key 2t)
        pa[mt][2] = to_tf32(sc[mt][kb][1]);              // (row g,   key 2t+1)
        pa[mt][3] = to_tf32(sc[mt][kb][3]);              // (row g+8, key 2t+1)
      }
#pragma unroll
      for (int dn = 0; dn < 4; ++dn) {
        const uint32_t b0 = __float_as_uint(Vs[(8 * kb + 2 * t) * PITCH + 8 * dn + g]);
        const uint32_t b1 = __float_as_uint(Vs[(8 * kb + 2 * t + 1) * PITCH + 8 * dn + g]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) mma_tf32_16x8x8(oc[mt][dn], pa[mt], b0, b1);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int e2 = 0; e2 < 2; ++e2) {
        float ls = l[mt][e2];
        ls += __shfl_xor_sync(0xffffffffu, ls, 1);
        ls += __shfl_xor_sync(0xffffffffu, ls, 2);
        const float inv = 1.0f / ls;
        float* orow = out + ((size_t)bf * HW + q0 + 16 * mt + g + 8 * e2) * hid + head * DH;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn)
          *reinterpret_cast<float2*>(orow + 8 * dn + 2 * t) = make_float2(oc[mt][dn][2 * e2] * inv, oc[mt][dn][2 * e2 + 1] * inv);
      }
  }
}

// ------------------------------------------------------------------------------------------------------------
// mid-block spatial softmax attention: one thread per query token, K/V streamed through shared memory in tiles of 32
// tokens with an online softmax.  grid = (ceil(HW/128), heads, B*F).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
spatial_attention_kernel(const float* __restrict__ qkv, float* __restrict__ out, int HW, int heads) {
  __shared__ __align__(16) float Ks[32][DH];
  __shared__ __align__(16) float Vs[32][DH];
  const int head = blockIdx.y;
  const int64_t bf = blockIdx.z;
  const int tok = blockIdx.x * 128 + threadIdx.x;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const float* base = qkv + (size_t)bf * HW * C3 + head * DH;
  const bool active = tok < HW;
  float q[DH], o[DH];
#pragma unroll
  for (int c = 0; c < DH; ++c) { q[c] = 0.f; o[c] = 0.f; }
  if (active) {
    const float* row = base + (size_t)tok * C3;
#pragma unroll
    for (int c = 0; c < DH; c += 4) {
      float4 a = __ldcs(reinterpret_cast<const float4*>(row + c));
      q[c] = __fmul_rn(a.x, ATT_SCALE); q[c + 1] = __fmul_rn(a.y, ATT_SCALE);
      q[c + 2] = __fmul_rn(a.z, ATT_SCALE); q[c + 3] = __fmul_rn(a.w, ATT_SCALE);
    }
  }
  float mx = -INFINITY, l = 0.f;
  for (int j0 = 0; j0 < HW; j0 += 32) {
    __syncthreads();
    // 32 tokens x (32 k + 32 v) floats = 512 float4, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + 128 * i;  // 0..511
      const int t = idx >> 4, part = idx & 15;     // part 0..7 -> K, 8..15 -> V
      const int j = j0 + t;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < HW) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)j * C3 + (part < 8 ? hid : 2 * hid) + (part & 7) * 4));
      if (part < 8) *reinterpret_cast<float4*>(&Ks[t][(part & 7) * 4]) = v;
      else *reinterpret_cast<float4*>(&Vs[t][(part & 7) * 4]) = v;
    }
    __syncthreads();
    if (!active) continue;
    float s[32];
    float tmx = -INFINITY;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        float4 k4 = *reinterpret_cast<const float4*>(&Ks[t][c]);
        acc = fmaf(q[c], k4.x, acc);
        acc = fmaf(q[c + 1], k4.y, acc);
        acc = fmaf(q[c + 2], k4.z, acc);
        acc = fmaf(q[c + 3], k4.w, acc);
      }
      s[t] = (j0 + t < HW) ? acc : -INFINITY;
      tmx = fmaxf(tmx, s[t]);
    }
    const float nmx = fmaxf(mx, tmx);
    const float corr = expf(mx - nmx);  // first tile: exp(-inf) = 0
    l *= corr;
#pragma unroll
    for (int c = 0; c < DH; ++c) o[c] *= corr;
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      const float pj = expf(s[t] - nmx);  // masked tokens: exp(-inf) = 0
      l += pj;
#pragma unroll
      for (int c = 0; c < DH; c += 4) {
        float4 v4 = *reinterpret_cast<const float4*>(&Vs[t][c]);
        o[c] = fmaf(pj, v4.x, o[c]);
        o[c + 1] = fmaf(pj, v4.y, o[c + 1]);
        o[c + 2] = fmaf(pj, v4.z, o[c + 2]);
        o[c + 3] = fmaf(pj, v4.w, o[c + 3]);
      }
    }
    mx = nmx;
  }
  if (active) {
    const float inv = 1.0f / l;
    float* orow = out + ((size_t)bf * HW + tok) * hid + head * DH;
#pragma unroll
    for (int c = 0; c < DH; c += 4)
      *reinterpret_cast<float4*>(orow + c) = make_float4(o[c] * inv, o[c + 1] * inv, o[c + 2] * inv, o[c + 3] * inv);
  }
}

// ------------------------------------------------------------------------------------------------------------
// spatial linear attention (conv3d.py:243-257), two kernels.
// pass A: context[d][e] = sum_n softmax_n(k)[n][d] * v[n][e] per (frame, head).  One CTA (128 threads) per (frame, head).
//   1) column max of k over the pixels: float4 loads, 8 lanes cover the 128-byte head slice of one pixel (coalesced)
//   2) tiles of 64 pixels of k and v are staged in shared memory with coalesced float4 loads, k is exponentiated in
//      place, then thread (d = lane, e-block = warp) owns 8 context entries: per pixel one conflict-free LDS of
//      exp(k), two broadcast LDS.128 of v and 8 FMAs.  No cross-warp reduction, 8 accumulators per thread.
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
linattn_context_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, int HW, int heads, float* __restrict__ kstat,
                       float vscale) {
  constexpr int TP = 64;                       // pixels per tile
  __shared__ __align__(16) float s_k[2][TP][DH];
  __shared__ __align__(16) float s_v[2][TP][DH];
  __shared__ float s_red[16][DH];
  __shared__ float s_max[DH];
  __shared__ float s_sum[DH];
  const int head = blockIdx.x % heads;
  const int64_t bf = blockIdx.x / heads;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const float* kbase = qkv + (size_t)bf * HW * C3 + hid + head * DH;
  const float* vbase = kbase + hid;
  const int prow = tid >> 3, pc = (tid & 7) * 4;   // 16 pixel rows x 8 float4 per pass

  // ---- 1) max over pixels ----
  float4 mx4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  for (int n = prow; n < HW; n += 16) {
    const float4 k4 = __ldg(reinterpret_cast<const float4*>(kbase + (size_t)n * C3 + pc));
    mx4.x = fmaxf(mx4.x, k4.x); mx4.y = fmaxf(mx4.y, k4.y); mx4.z = fmaxf(mx4.z, k4.z); mx4.w = fmaxf(mx4.w, k4.w);
  }
  s_red[prow][pc] = mx4.x; s_red[prow][pc + 1] = mx4.y; s_red[prow][pc + 2] = mx4.z; s_red[prow][pc + 3] = mx4.w;
  __syncthreads();
  if (tid < DH) {
    float m = s_red[0][tid];
#pragma unroll
    for (int r = 1; r < 16; ++r) m = fmaxf(m, s_red[r][tid]);
    s_max[tid] = m;
  }
  __syncthreads();
  const float4 cmax = *reinterpret_cast<const float4*>(&s_max[pc]);

  // ---- 2) context accumulation over tiles ----
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  float ksum = 0.f;
  const int ntiles = (HW + TP - 1) / TP;
  float4 kreg[4], vreg[4];
  auto load_tile = [&](int tile) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = tile * TP + prow + 16 * i;
      if (n < HW) {
        kreg[i] = __ldg(reinterpret_cast<const float4*>(kbase + (size_t)n * C3 + pc));
        vreg[i] = __ldg(reinterpret_cast<const float4*>(vbase + (size_t)n * C3 + pc));
      } else {
        kreg[i] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);   // exp -> 0
        vreg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = prow + 16 * i;
      *reinterpret_cast<float4*>(&s_k[buf][r][pc]) = make_float4(expf(kreg[i].x - cmax.x), expf(kreg[i].y - cmax.y),
                                                                  expf(kreg[i].z - cmax.z), expf(kreg[i].w - cmax.w));
      *reinterpret_cast<float4*>(&s_v[buf][r][pc]) = vreg[i];
    }
  };
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int tile = 0; tile < ntiles; ++tile) {
    const int buf = tile & 1;
    if (tile + 1 < ntiles) load_tile(tile + 1);        // global loads of the next tile overlap the FMAs below
#pragma unroll 8
    for (int n = 0; n < TP; ++n) {
      const float ek = s_k[buf][n][lane];
      const float4 v0 = *reinterpret_cast<const float4*>(&s_v[buf][n][warp * 8]);
      const float4 v1 = *reinterpret_cast<const float4*>(&s_v[buf][n][warp * 8 + 4]);
      if (warp == 0) ksum += ek;
      acc[0] = fmaf(ek, v0.x, acc[0]); acc[1] = fmaf(ek, v0.y, acc[1]);
      acc[2] = fmaf(ek, v0.z, acc[2]); acc[3] = fmaf(ek, v0.w, acc[3]);
      acc[4] = fmaf(ek, v1.x, acc[4]); acc[5] = fmaf(ek, v1.y, acc[5]);
      acc[6] = fmaf(ek, v1.z, acc[6]); acc[7] = fmaf(ek, v1.w, acc[7]);
    }
    if (tile + 1 < ntiles) store_tile(buf ^ 1);
    __syncthreads();
  }
  if (warp == 0) s_sum[lane] = ksum;
  __syncthreads();
  const float inv = vscale / s_sum[lane];     // vscale: the 2-D variant's v / (h*w) (jf.py:219), 1 for conv3d.py:243-257
  if (kstat && warp == 0) {                   // softmax_n(k) statistics for the backward pass: [frame*heads + head][d][max, sum]
    kstat[((size_t)blockIdx.x * DH + lane) * 2 + 0] = s_max[lane];
    kstat[((size_t)blockIdx.x * DH + lane) * 2 + 1] = s_sum[lane];
  }
  float* dst = ctx + (size_t)blockIdx.x * DH * DH + lane * DH + warp * 8;   // [d][e]
  *reinterpret_cast<float4*>(dst) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
}

// pass B: out[n][e] = sum_d context[d][e] * (softmax_d(q[n])[d] * scale).  One CTA = 128 pixels of one (frame, head):
// q rows are staged through shared memory with coalesced loads, thread-per-pixel math, coalesced stores.
__global__ void __launch_bounds__(128)
linattn_apply_kernel(const float* __restrict__ qkv, const float* __restrict__ ctx, float* __restrict__ out, int HW,
                     int heads) {
  __shared__ __align__(16) float s_ctx[DH][DH];
  __shared__ __align__(16) float s_q[128][DH + 4];
  const int head = blockIdx.y;
  const int64_t bf = blockIdx.z;
  const int tid = threadIdx.x;
  const float* cp = ctx + ((size_t)bf * heads + head) * DH * DH;
  for (int i = tid; i < DH * DH; i += 128) (&s_ctx[0][0])[i] = __ldg(cp + i);
  const int C3 = 3 * heads * DH, hid = heads * DH;
  const int tok0 = blockIdx.x * 128;
  const int prow = tid >> 3, pc = (tid & 7) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = prow + 16 * i;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tok0 + r < HW) v = __ldcs(reinterpret_cast<const float4*>(qkv + ((size_t)bf * HW + tok0 + r) * C3 + head * DH + pc));
    *reinterpret_cast<float4*>(&s_q[r][pc]) = v;
  }
  __syncthreads();
  float q[DH];
  float mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < DH; c += 4) {
    const float4 a = *reinterpret_cast<const float4*>(&s_q[tid][c]);
    q[c] = a.x; q[c + 1] = a.y; q[c + 2] = a.z; q[c + 3] = a.w;
    mx = fmaxf(mx, fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
  }
  float l = 0.f;
#pragma unroll
  for (int c = 0; c < DH; ++c) {
    q[c] = expf(q[c] - mx);
    l += q[c];
  }
  const float inv = 1.0f / l;
  float o[DH];
#pragma unroll
  for (int e = 0; e < DH; ++e) o[e] = 0.f;
#pragma unroll
  for (int d = 0; d < DH; ++d) {
    const float qs = __fmul_rn(q[d] * inv, ATT_SCALE);
#pragma unroll
    for (int e = 0; e < DH; e += 4) {
      const float4 c4 = *reinterpret_cast<const float4*>(&s_ctx[d][e]);
      o[e] = fmaf(qs, c4.x, o[e]);
      o[e + 1] = fmaf(qs, c4.y, o[e + 1]);
      o[e + 2] = fmaf(qs, c4.z, o[e + 2]);
      o[e + 3] = fmaf(qs, c4.w, o[e + 3]);
    }
  }
#pragma unroll
  for (int e = 0; e < DH; e += 4) *reinterpret_cast<float4*>(&s_q[tid][e]) = make_float4(o[e], o[e + 1], o[e + 2], o[e + 3]);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = prow + 16 * i;
    if (tok0 + r < HW)
      __stcs(reinterpret_cast<float4*>(out + ((size_t)bf * HW + tok0 + r) * hid + head * DH + pc),
             *reinterpret_cast<const float4*>(&s_q[r][pc]));
  }
}

}  // namespace dpc

extern "C" int dpc_temporal_attention(const float* qkv, const float* rope_cos, const float* rope_sin,
                                      const float* pos_bias, float* out, int32_t B, int32_t F, int32_t HW,
                                      int32_t heads, int32_t use_rope, int32_t precise, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(qkv && out && B > 0 && F > 0 && F <= 64 && HW > 0 && heads > 0);
  DPC_CHECK_ARG(!(use_rope & 1) || (rope_cos && rope_sin));
  const int64_t total = (int64_t)B * HW * heads;
  int64_t blocks = (total + 3) / 4;
  const int64_t cap = 148LL * 64;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  // F <= 32 and 32 < F <= 64 both run the tensor-core kernel (frames padded to 32 / 64, padded keys masked)
  const int NFr = F <= 32 ? 32 : 64;
  DPC_CHECK_ARG(heads <= 16);
  const size_t smem = ((size_t)4 * 3 * NFr * 36 + (size_t)heads * 2 * NFr) * sizeof(float);
  static bool configured_[kMaxDevices] = {};
  bool& configured = configured_[device_ordinal()];
  if (!configured) {
    const int s32 = (4 * 3 * 32 * 36 + 16 * 64) * (int)sizeof(float), s64 = (4 * 3 * 64 * 36 + 16 * 128) * (int)sizeof(float);
    DPC_CUDA(cudaFuncSetAttribute(temporal_attention_mma_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, s32));
    DPC_CUDA(cudaFuncSetAttribute(temporal_attention_mma_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, s32));
    DPC_CUDA(cudaFuncSetAttribute(temporal_attention_mma_kernel<false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
    DPC_CUDA(cudaFuncSetAttribute(temporal_attention_mma_kernel<true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
    configured = true;
  }
  if (F <= 32) {
    if (precise)
      temporal_attention_mma_kernel<true, 32><<<(unsigned)blocks, 128, smem, st>>>(qkv, rope_cos, rope_sin, pos_bias, out, total, F,
                                                                                  HW, heads, use_rope);
    else
      temporal_attention_mma_kernel<false, 32><<<(unsigned)blocks, 128, smem, st>>>(qkv, rope_cos, rope_sin, pos_bias, out, total, F,
                                                                                   HW, heads, use_rope);
  } else {
    if (precise)
      temporal_attention_mma_kernel<true, 64><<<(unsigned)blocks, 128, smem, st>>>(qkv, rope_cos, rope_sin, pos_bias, out, total, F,
                                                                                  HW, heads, use_rope);
    else
      temporal_attention_mma_kernel<false, 64><<<(unsigned)blocks, 128, smem, st>>>(qkv, rope_cos, rope_sin, pos_bias, out, total, F,
                                                                                   HW, heads, use_rope);
  }
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_spatial_attention(const float* qkv, float* out, int32_t BF, int32_t HW, int32_t heads, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(qkv && out && BF > 0 && BF <= 65535 && HW > 0 && heads > 0);
  dim3 grid((unsigned)((HW + 127) / 128), (unsigned)heads, (unsigned)BF);
  spatial_attention_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(qkv, out, HW, heads);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_spatial_attention_mma(const float* qkv, float* out, int32_t BF, int32_t HW, int32_t heads, void* stream) {
  using namespace dpc;
  if (HW % 64 != 0) return -2;                            // served by dpc_spatial_attention (fp32 SIMT)
  DPC_CHECK_ARG(qkv && out && BF > 0 && BF <= 65535 && HW > 0 && heads > 0);
  dim3 grid((unsigned)((HW + 255) / 256), (unsigned)heads, (unsigned)BF);
  spatial_attention_mma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(qkv, out, HW, heads);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_spatial_linear_attention(const float* qkv, float* ctx_ws, float* out, int32_t BF, int32_t HW,
                                            int32_t heads, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(qkv && ctx_ws && out && BF > 0 && BF <= 65535 && HW > 0 && heads > 0);
  cudaStream_t st = (cudaStream_t)stream;
  linattn_context_kernel<<<(unsigned)((int64_t)BF * heads), 128, 0, st>>>(qkv, ctx_ws, HW, heads, nullptr, 1.0f);
  DPC_LAUNCH_CHECK();
  dim3 grid((unsigned)((HW + 127) / 128), (unsigned)heads, (unsigned)BF);
  linattn_apply_kernel<<<grid, 128, 0, st>>>(qkv, ctx_ws, out, HW, heads);
  DPC_LAUNCH_CHECK();
  return 0;
}

extern "C" int dpc_spatial_linear_attention_ex(const float* qkv, float* ctx_ws, float* kstat, float* out, int32_t BF, int32_t HW,
                                               int32_t heads, float v_scale, void* stream) {
  using namespace dpc;
  DPC_CHECK_ARG(qkv && ctx_ws && out && BF > 0 && BF <= 65535 && HW > 0 && heads > 0 && heads <= 65535);
  cudaStream_t st = (cudaStream_t)stream;
  linattn_context_kernel<<<(unsigned)((int64_t)BF * heads), 128, 0, st>>>(qkv, ctx_ws, HW, heads, kstat, v_scale);
  DPC_LAUNCH_CHECK();
  dim3 grid((unsigned)((HW + 127) / 128), (unsigned)heads, (unsigned)BF);
  linattn_apply_kernel<<<grid, 128, 0, st>>>(qkv, ctx_ws, out, HW, heads);
  DPC_LAUNCH_CHECK();
  return 0;
}
