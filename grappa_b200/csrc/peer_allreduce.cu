// Gradient all-reduce over NVLink peer memory (data-parallel training, one process per GPU).
//
// Every rank's flat gradient buffer lives in a cudaMalloc allocation that all other ranks of the node map through CUDA
// IPC, so a kernel can load and store any rank's gradients directly over NVLink / NVSwitch.  One launch reduces one span
// in two shots without leaving the kernel:
//   1. ready barrier  : CTA b of every rank tells CTA b of every peer that the rank's gradients of this span are final
//                       (flag words in peer memory, monotonically increasing epochs -- nothing is ever reset);
//   2. reduce-scatter : rank r owns slice r of the span; it sums the G copies of that slice (16-byte loads from the peers,
//                       fixed order 0..G-1, so the result does not depend on timing) ...
//   3. all-gather     : ... and stores the sum straight into every rank's buffer (slices are disjoint: nobody reads what
//                       another rank writes);
//   4. done barrier   : a rank's kernel ends only after every peer's CTA b has finished writing into its buffer.
// Against NCCL's ring (2 (G-1) dependent steps per all-reduce, LL protocol at 50 % payload, 13-29 kernels of 50-200 us
// each that hold their SMs while they wait for the slowest rank) this moves (G-1)/G of the span in and out once, on a
// grid of a few CTAs that the caller sizes.  Every slice is reduced by exactly ONE rank and broadcast, so all ranks hold
// bit-identical gradients afterwards (bench.py `param_sync`).
//
// Spin waits give up after ~30 s (error word set; later launches do not wait at all) so that a rank that died cannot
// hang the others' GPUs.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "common.cuh"
#include "grappa_b200.h"

namespace gb {

constexpr int AR_THREADS = 512;
constexpr int AR_UNROLL = 2;
constexpr unsigned long long AR_SPIN_LIMIT = 60000000000ull;   // clock64 ticks (~30 s: a peer may be capturing a CUDA graph)

__device__ __forceinline__ void st_flag(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_flag(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// flags of rank q: [GB_MAX_PEERS sources][GB_PEER_MAX_CTAS] words; word [src][b] is written by CTA b of rank src
__device__ __forceinline__ bool cross_rank_barrier(const gb_peer_allreduce_args& a, uint32_t value, uint32_t* err) {
  __syncthreads();
  bool ok = true;
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_flag(a.flags[peer] + a.rank * GB_PEER_MAX_CTAS + blockIdx.x, value);
    const uint32_t* mine = a.flags[a.rank] + peer * GB_PEER_MAX_CTAS + blockIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_flag(mine) - value) < 0 && *(volatile uint32_t*)err == 0u) {   // after one time-out nobody waits again
      if ((unsigned long long)(clock64() - t0) > AR_SPIN_LIMIT) {
        *err = 1u;
        ok = false;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  return ok;
}

// One thread per peer: tell it `value` in word [rank][slot] of its flag block, wait for its `value` in ours.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const gb_peer_allreduce_args a, int slot) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = ++a.epoch[slot];
  __syncthreads();
  const uint32_t e = s_epoch;
  uint32_t* err = a.epoch + GB_PEER_MAX_CTAS;
  if ((int)threadIdx.x < a.world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_flag(a.flags[peer] + a.rank * GB_PEER_MAX_CTAS + slot, e);
    const uint32_t* mine = a.flags[a.rank] + peer * GB_PEER_MAX_CTAS + slot;
    const long long t0 = clock64();
    while ((int32_t)(ld_flag(mine) - e) < 0 && *(volatile uint32_t*)err == 0u) {
      if ((unsigned long long)(clock64() - t0) > AR_SPIN_LIMIT) {
        *err = 1u;
        break;
      }
      __nanosleep(64);
    }
  }
}

// BARRIERS = true : one self-contained launch (ready barrier, reduce + broadcast, done barrier).
// BARRIERS = false: the middle part only; the launcher brackets it with two one-warp peer_barrier_kernel launches, so
//                   that the wait for the slowest rank does not hold `ctas` whole SMs (this kernel needs all registers
//                   of an SM; measured: its launches last ~60 us at 2 GPUs of which ~10 us move data).
template <bool BARRIERS>
__global__ void __launch_bounds__(AR_THREADS) peer_allreduce_kernel(const gb_peer_allreduce_args a) {
  __shared__ uint32_t s_epoch;
  uint32_t e = 0;
  uint32_t* err = a.epoch + GB_PEER_MAX_CTAS;
  if (BARRIERS) {
    if (threadIdx.x == 0) s_epoch = ++a.epoch[blockIdx.x];   // every CTA counts its own launches (same sequence on all ranks)
    __syncthreads();
    e = s_epoch;
    cross_rank_barrier(a, 2u * e - 1u, err);   // all ranks' gradients of this span are final
  }

  // slice of this rank, in float4 units (the span starts 16-byte aligned: flat offsets are multiples of 4 floats)
  const long long n4 = (a.count + 3) / 4;
  const long long per = (n4 + a.world - 1) / a.world;
  const long long lo = min(n4, per * a.rank), hi = min(n4, lo + per);
  const long long base4 = a.start / 4;
  const long long stride = (long long)gridDim.x * AR_THREADS;
  // Per thread and iteration: AR_UNROLL 16-byte loads from EVERY rank are issued before the first one is consumed
  // (a 512-thread CTA keeps up to 128 KB in flight -- NVLink reads take ~2-4 us), then summed in rank order.
  for (long long i = lo + (long long)blockIdx.x * AR_THREADS + threadIdx.x; i < hi; i += AR_UNROLL * stride) {
    float4 v[GB_MAX_PEERS][AR_UNROLL];
    bool on[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) on[u] = i + u * stride < hi;
#pragma unroll
    for (int p = 0; p < GB_MAX_PEERS; ++p) {
      if (p < a.world) {
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u)
          if (on[u]) v[p][u] = ld_peer(reinterpret_cast<const float4*>(a.data[p]) + base4 + i + u * stride);
      }
    }
    float4 acc[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int p = 0; p < GB_MAX_PEERS; ++p) {          // fixed order: the sum is reproducible
      if (p < a.world) {
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u)
          if (on[u]) {
            acc[u].x += v[p][u].x; acc[u].y += v[p][u].y; acc[u].z += v[p][u].z; acc[u].w += v[p][u].w;
          }
      }
    }
#pragma unroll
    for (int p = 0; p < GB_MAX_PEERS; ++p) {
      if (p < a.world) {
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u)
          if (on[u]) st_peer(reinterpret_cast<float4*>(a.data[p]) + base4 + i + u * stride, acc[u]);
      }
    }
  }
  if (BARRIERS) cross_rank_barrier(a, 2u * e, err);        // every peer has written its slice into this rank's buffer
}

}  // namespace gb

using namespace gb;

extern "C" int grappa_b200_ipc_alloc(int64_t bytes, void** ptr, gb_ipc_handle* handle) {
  if (bytes <= 0 || !ptr || !handle) { set_error("ipc_alloc: bad arguments"); return GB_ERR_INVALID; }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) { set_error("ipc_alloc: cudaMalloc(%lld) -> %s", (long long)bytes, cudaGetErrorString(e)); return GB_ERR_CUDA; }
  e = cudaMemset(p, 0, (size_t)bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); set_error("ipc_alloc: %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(gb_ipc_handle), "handle size");
  std::memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return GB_OK;
}

extern "C" int grappa_b200_ipc_open(const gb_ipc_handle* handle, void** ptr) {
  if (!handle || !ptr) { set_error("ipc_open: bad arguments"); return GB_ERR_INVALID; }
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("ipc_open: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
  return GB_OK;
}

extern "C" int grappa_b200_ipc_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { set_error("ipc_close -> %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
  return GB_OK;
}

extern "C" int grappa_b200_ipc_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { set_error("ipc_free -> %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
  return GB_OK;
}

extern "C" int grappa_b200_peer_allreduce(const gb_peer_allreduce_args* a, void* stream) {
  if (!a || a->world < 1 || a->world > GB_MAX_PEERS || a->rank < 0 || a->rank >= a->world || a->count < 0 || (a->start & 3) ||
      a->ctas < 1 || a->ctas > GB_PEER_MAX_CTAS || !a->epoch) {
    set_error("peer_allreduce: bad arguments (world %d rank %d start %lld count %lld ctas %d)", a ? a->world : -1, a ? a->rank : -1,
              a ? (long long)a->start : 0ll, a ? (long long)a->count : 0ll, a ? a->ctas : -1);
    return GB_ERR_INVALID;
  }
  for (int p = 0; p < a->world; ++p)
    if (!a->data[p] || !a->flags[p]) { set_error("peer_allreduce: rank %d has no mapping of rank %d", a->rank, p); return GB_ERR_INVALID; }
  if (a->count == 0) return GB_OK;
  if (a->split) {
    // slots GB_PEER_MAX_CTAS - 2 / - 1 of the flag blocks and epoch counters belong to the two barrier kernels
    if (a->ctas > GB_PEER_MAX_CTAS - 2) { set_error("peer_allreduce: split mode supports up to %d CTAs", GB_PEER_MAX_CTAS - 2); return GB_ERR_INVALID; }
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a, GB_PEER_MAX_CTAS - 2);
    count_launch();
    peer_allreduce_kernel<false><<<a->ctas, AR_THREADS, 0, (cudaStream_t)stream>>>(*a);
    count_launch();
    peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*a, GB_PEER_MAX_CTAS - 1);
  } else {
    peer_allreduce_kernel<true><<<a->ctas, AR_THREADS, 0, (cudaStream_t)stream>>>(*a);
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("peer_allreduce launch -> %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
  return GB_OK;
}
