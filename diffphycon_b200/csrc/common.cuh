// Shared helpers for the sm_100a kernels of libdpc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dpc_b200.h"

namespace dpc {

// thread-local last-error text behind dpc_last_error()
char* err_buf();
int set_err(int code, const char* what, const char* file, int line);

#define DPC_CHECK_ARG(cond)                                                        \
  do {                                                                             \
    if (!(cond)) return dpc::set_err(-1, "argument check failed: " #cond, __FILE__, __LINE__); \
  } while (0)

#define DPC_CUDA(expr)                                                             \
  do {                                                                             \
    cudaError_t e__ = (expr);                                                      \
    if (e__ != cudaSuccess) return dpc::set_err((int)e__, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

#define DPC_LAUNCH_CHECK() DPC_CUDA(cudaGetLastError())

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double shfl_xor_double(double v, int o) {
  int lo = __double2loint(v), hi = __double2hiint(v);
  lo = __shfl_xor_sync(0xffffffffu, lo, o);
  hi = __shfl_xor_sync(0xffffffffu, hi, o);
  return __hiloint2double(hi, lo);
}

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + expf(-v)); }

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Per-device launch state.  cudaFuncSetAttribute(MaxDynamicSharedMemorySize) and the SM count belong to ONE device, so every
// cache of them is indexed by the ordinal of the device that is current at launch time (the binding makes the tensors'
// device current around every call).
constexpr int kMaxDevices = 64;
inline int device_ordinal() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) return 0;
  return d;
}
inline int sm_count(int dev) {
  static int n[kMaxDevices] = {};
  if (!n[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) n[dev] = v;
  }
  return n[dev] > 0 ? n[dev] : 1;
}

}  // namespace dpc
