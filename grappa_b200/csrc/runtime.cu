// Device-property cache and small runtime helpers shared by all kernels.
#include "common.cuh"

namespace gb {
int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}
}  // namespace gb

extern "C" int grappa_b200_sm_count(void) {
  int dev = 0, n = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) {
    gb::set_error("grappa_b200_sm_count: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return GB_ERR_CUDA;
  }
  return n;
}
