// Probe: the register layout of tcgen05.ld.16x256b against a pattern written with tcgen05.st.32x32b (thread = lane/row,
// register = column).  Prints, for a few threads, which (row, col) each returned register holds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_ld_probe tmem_ld_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe() {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = slot;
  const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
  // write: row r = warp*32+lane, col c -> value r*100 + c  (32 columns)
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = (warp * 32 + lane) * 100 + c;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(tl), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[8];
    // 16x256b.x2: 16 lanes x 16 columns starting at column 8, lanes [half*16, half*16+16) of this warp's quarter
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(tl + ((uint32_t)(half * 16) << 16) + 8));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (warp == 1 && (lane < 6 || lane == 31))
      printf("half %d lane %2d: %u %u %u %u | %u %u %u %u\n", half, lane, r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7]);
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(64) : "memory");
}
int main() {
  probe<<<1, 128>>>();
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
