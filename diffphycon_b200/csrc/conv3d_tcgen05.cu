// placeholder until the TMA/tcgen05 kernel lands: reports "shape not supported" so the host uses dpc_conv_igemm.
#include "common.cuh"
extern "C" int dpc_conv3d_tcgen05(const dpc_conv_params* p, void* stream) {
  (void)p; (void)stream;
  return -2;
}
