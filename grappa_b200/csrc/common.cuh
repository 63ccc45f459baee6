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

// true the first time it is called for `mask` on the CURRENT device: function attributes (max dynamic shared memory) are
// per device, so one-off kernel set-up is tracked per device ordinal, not per process
inline bool first_use_on_device(unsigned long long& mask) {
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
}

#ifdef __CUDACC__
// Programmatic dependent launch (PDL).  Every kernel calls pdl_trigger() first thing: once all CTAs of a grid have done
// so (or exited), a following kernel that was launched with the programmatic-stream-serialization attribute may start
// its CTAs -- they run their prologue (barrier init, TMEM allocation, descriptor prefetch) while this grid's tail is
// still executing -- and must call pdl_wait() before touching global memory: it returns when the preceding grid has
// completed and its writes are visible.  Only the tensor-core GEMM is launched with the attribute (heaviest prologue,
// 230 launches per step); for kernels launched normally both calls are no-ops.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// Counter-based dropout mask.  Stateless, so the backward pass regenerates the identical mask from (seed, index).
// Elements are handled in aligned groups of four consecutive indices: two 32-bit finalisers (murmur3 fmix32, 32-bit
// multiplies only -- the 64-bit splitmix this replaces cost ~35 instructions per element and made the fused GEMM
// epilogues 2x longer than the mainloop) give 4 x 16 random bits per group; keep iff bits16 >= p * 2^16.
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
// seed of this launch = static seed + run-time device counter * odd constant (fresh masks per graph replay)
__device__ __forceinline__ uint64_t seed_with_offset(uint64_t seed, const uint64_t* offset) {
  return offset ? seed + __ldg(reinterpret_cast<const unsigned long long*>(offset)) * 0xD1342543DE82EF95ull : seed;
}
// key of group `g` (= element index >> 2); distinct for distinct g below 2^31 groups
__device__ __forceinline__ uint32_t dropout_key(uint64_t seed, uint64_t g) {
  uint32_t k = ((uint32_t)g * 0x9E3779B1u + (uint32_t)seed) ^ ((uint32_t)(seed >> 32) * 0x85EBCA77u);
  if (g >> 32) k ^= (uint32_t)(g >> 32) * 0xC2B2AE3Du;
  return k;
}
// 16-bit threshold from the 32-bit one the host computes (p * 2^32)
__device__ __forceinline__ uint32_t dropout_thresh16(uint32_t thresh) {
  const uint32_t t = (thresh >> 16) + ((thresh >> 15) & 1u);
  return t == 0u ? 1u : (t > 65535u ? 65535u : t);
}
// the four scale factors of the aligned group that starts at element index `base` (base % 4 == 0)
__device__ __forceinline__ float4 dropout_scale4(uint64_t seed, uint64_t base, uint32_t thresh, float inv_keep) {
  const uint32_t k = dropout_key(seed, base >> 2);
  const uint32_t h0 = fmix32(k), h1 = fmix32(k ^ 0x80000000u);
  const uint32_t t = dropout_thresh16(thresh);
  return make_float4((h0 & 0xffffu) >= t ? inv_keep : 0.f, (h0 >> 16) >= t ? inv_keep : 0.f,
                     (h1 & 0xffffu) >= t ? inv_keep : 0.f, (h1 >> 16) >= t ? inv_keep : 0.f);
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, uint32_t thresh, float inv_keep) {
  const uint32_t k = dropout_key(seed, idx >> 2);
  const uint32_t h = fmix32((idx & 2) ? (k ^ 0x80000000u) : k);
  const uint32_t b = (idx & 1) ? (h >> 16) : (h & 0xffffu);
  return b >= dropout_thresh16(thresh) ? inv_keep : 0.f;
}
#endif

}  // namespace gb
