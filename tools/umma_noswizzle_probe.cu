// Hardware probe (development aid, not part of the library): K-major NO-SWIZZLE UMMA descriptor over a raw channels-last image
// row ([pixels][4 floats], 16 B per pixel) with OVERLAPPING core matrices: 8-row groups 128 B apart, the two K core matrices of
// one MMA 16 B apart, i.e. A[m][k] = raw[4m + k] (a sliding window = the im2col of a 1 x 1 x kw conv without copying).
// mode 0: LBO = 16 B, SBO = 128 B; mode 1: LBO = 128 B, SBO = 16 B (which field is which for K-major no-swizzle).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o probe tools/umma_noswizzle_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (true) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    if (clock64() - t0 > 2000000000LL) { printf("timeout\n"); __trap(); }
  }
}
__device__ __forceinline__ uint64_t desc(uint32_t addr, uint32_t base_off) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_off & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_ns(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;                                             // layout type 0: no swizzle
}

constexpr int ROWS = 256, N = 64;

__global__ void __launch_bounds__(128, 1) probe(const float* rawA, const __grid_constant__ CUtensorMap tmB,
                                                float* D, int rshift, int use_base_off) {
  extern __shared__ uint8_t raw[];
  uint32_t base = (s32(raw) + 1023u) & ~1023u;
  uint32_t a_buf = base, b_buf = base + ROWS * 128, bar = b_buf + N * 128, bar2 = bar + 8, slot = bar + 16;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar2));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  {
    float* sa = reinterpret_cast<float*>(raw + (a_buf - s32(raw)));
    for (int i = threadIdx.x; i < ROWS * 32; i += 128) sa[i] = rawA[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(N * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(b_buf), "l"(&tmB), "r"(bar), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t a0 = a_buf + rshift * 16;
    for (int k = 0; k < 4; ++k) {
      uint32_t aa = a0 + k * 32;
      uint64_t da = use_base_off ? desc_ns(aa, 128, 16) : desc_ns(aa, 16, 128), db = desc(b_buf + k * 32, 0);
      uint32_t acc = k;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar2) : "memory");
  }
  mbar_wait(bar2, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 32) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(warp * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

static float trunc_tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u &= 0xffffe000u; memcpy(&x, &u, 4); return x; }

int main() {
  typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  Enc enc = (Enc)fp;
  float *hA = (float*)malloc(ROWS * 32 * 4), *hB = (float*)malloc(N * 32 * 4), *hD = (float*)malloc(128 * N * 4);
  srand(1);
  for (int i = 0; i < ROWS * 32; ++i) hA[i] = (rand() % 2001 - 1000) / 1000.0f;
  for (int i = 0; i < N * 32; ++i) hB[i] = (rand() % 2001 - 1000) / 1000.0f;
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, ROWS * 32 * 4)); CK(cudaMalloc(&dB, N * 32 * 4)); CK(cudaMalloc(&dD, 128 * N * 4));
  CK(cudaMemcpy(dA, hA, ROWS * 32 * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB, N * 32 * 4, cudaMemcpyHostToDevice));
  CUtensorMap tB;
  cuuint64_t dimsB[2] = {32, N}, str[1] = {128};
  cuuint32_t boxB[2] = {32, N}, es[2] = {1, 1};
  if (enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimsB, str, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) { printf("encB fail\n"); return 1; }
  size_t smem = ROWS * 128 + N * 128 + 1024 + 64;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int shifts[] = {0, 1, 2, 3, 5, 7, 8, 9, 64, 66, 67, 127};
  for (int mode = 0; mode < 2; ++mode)
    for (int si = 0; si < 12; ++si) {
      int r = shifts[si];
      CK(cudaMemset(dD, 0, 128 * N * 4));
      probe<<<1, 128, smem>>>(dA, tB, dD, r, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, r, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < 32; ++k) ref += (double)trunc_tf32(hA[4 * (r + i) + k]) * (double)trunc_tf32(hB[n * 32 + k]);
          double d = fabs(ref - hD[i * N + n]);
          if (d > maxerr) maxerr = d;
        }
      printf("mode=%d shift=%3d pixels: max |err| = %.3e %s\n", mode, r, maxerr, maxerr < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
