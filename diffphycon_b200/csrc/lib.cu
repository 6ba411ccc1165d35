// Library-level entry points: ABI version, last-error text, device check.
#include "common.cuh"

#include <string.h>

namespace dpc {

char* err_buf() {
  static thread_local char buf[512] = "";
  return buf;
}

int set_err(int code, const char* what, const char* file, int line) {
  snprintf(err_buf(), 512, "%s (%s:%d, code %d)", what, file, line, code);
  return code;
}

}  // namespace dpc

extern "C" int dpc_abi_version(void) { return DPC_ABI_VERSION; }

extern "C" const char* dpc_last_error(void) { return dpc::err_buf(); }

extern "C" int dpc_device_is_sm100(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return -(int)e;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return -(int)e;
  return (prop.major == 10 && prop.minor == 0) ? 1 : 0;
}
