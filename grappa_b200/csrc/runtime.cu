// Device-property cache and small runtime helpers shared by all kernels.
#include <stdlib.h>

#include "common.cuh"

#include <atomic>
namespace gb {
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }
int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
  cached = n;
  return n;
}
long long launches();
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

extern "C" int64_t grappa_b200_launch_count(void) { return gb::launches(); }
