// Shared helpers for the grappa_b200 sm_100a kernels: error reporting, launch checks, warp utils.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/grappa_b200.h"

namespace gb {

// thread-local last-error string returned by grappa_b200_last_error()
void set_error(const char* fmt, ...);

#define GB_CHECK_CUDA(expr)                                                                 \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      gb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return GB_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define GB_CHECK_LAUNCH()                                                                   \
  do {                                                                                      \
    gb::count_launch();                                                                     \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      gb::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return GB_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

#define GB_REQUIRE(cond, ...)                \
  do {                                       \
    if (!(cond)) {                           \
      gb::set_error(__VA_ARGS__);            \
      return GB_ERR_INVALID;                 \
    }                                        \
  } while (0)

void count_launch();  // bumps the process-wide kernel launch counter (grappa_b200_launch_count)
int sm_count();  // cached multiprocessor count of the current device (148 on B200)

#ifdef __CUDACC__
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
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }
// d elu(x)/dx expressed through y = elu(x):  y > 0 ? 1 : y + 1
__device__ __forceinline__ float elu1_grad_from_out(float y) { return y > 0.f ? 1.f : y + 1.f; }

// Counter-based dropout mask: keep iff hash(seed, index) >= p * 2^32.  Stateless, so the backward
// pass regenerates the identical mask from (seed, index).
__device__ __forceinline__ uint32_t mix32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
// seed of this launch = static seed + run-time device counter * odd constant (fresh masks per graph replay)
__device__ __forceinline__ uint64_t seed_with_offset(uint64_t seed, const uint64_t* offset) {
  return offset ? seed + __ldg(reinterpret_cast<const unsigned long long*>(offset)) * 0xD1342543DE82EF95ull : seed;
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, uint32_t thresh, float inv_keep) {
  return mix32(seed, idx) >= thresh ? inv_keep : 0.f;
}
#endif

}  // namespace gb
