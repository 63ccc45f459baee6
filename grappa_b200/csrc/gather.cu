// K9 / K11 data movement: tuple gather (+ positional-encoding column) and the symmetriser's permuted
// concatenation, with deterministic gather-style backward passes (no atomics).
// Reference: interaction_parameters.py:173-178 (atom_feats[idxs].transpose(0,1).contiguous()),
// perm_equiv_transformer.py:134-141 (PE concat), :246-262 (stack of permuted copies).
#include "common.cuh"

namespace gb {

// x[l*T+t, 0:F] = p[idx[t*L+l], 0:F]; x[l*T+t, F:E] = pe[l]     (one warp per output row)
__global__ void __launch_bounds__(256) tuple_gather_fwd_kernel(const float* __restrict__ p, int ldp,
                                                               const int* __restrict__ idx, const float* __restrict__ pe,
                                                               float* __restrict__ x, int T, int L, int F, int E) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= (long long)L * T) return;
  const int l = (int)(row / T), t = (int)(row - (long long)l * T);
  const int atom = __ldg(idx + (size_t)t * L + l);
  const float* src = p + (size_t)atom * ldp;
  float* dst = x + (size_t)row * E;
  const bool vec = (ldp % 4 == 0) && (E % 4 == 0) && (((uintptr_t)p & 15) == 0) && (((uintptr_t)x & 15) == 0);
  if (vec) {
    const int nv = E >> 2;
    for (int c = lane; c < nv; c += 32) {
      float4 v;
      if (c * 4 + 3 < F) {
        v = __ldg(reinterpret_cast<const float4*>(src) + c);
      } else {
        float tmp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) tmp[i] = (c * 4 + i < F) ? __ldg(src + c * 4 + i) : (pe ? __ldg(pe + l) : 0.f);
        v = make_float4(tmp[0], tmp[1], tmp[2], tmp[3]);
      }
      reinterpret_cast<float4*>(dst)[c] = v;
    }
  } else {
    for (int c = lane; c < E; c += 32) dst[c] = c < F ? __ldg(src + c) : (pe ? __ldg(pe + l) : 0.f);
  }
}

// dp[n, 0:F] (+)= sum over incidences j in [inv_ptr[n], inv_ptr[n+1]): ent = t*L + l -> dx[l*T+t, 0:F]
__global__ void __launch_bounds__(256) tuple_gather_bwd_kernel(const float* __restrict__ dx, const int* __restrict__ inv_ptr,
                                                               const int* __restrict__ inv_ent, float* __restrict__ dp,
                                                               int ldp, int n_atoms, int T, int L, int F, int E,
                                                               int accumulate) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= n_atoms) return;
  const int j0 = __ldg(inv_ptr + n), j1 = __ldg(inv_ptr + n + 1);
  for (int c = lane; c < ldp; c += 32) {
    float acc = 0.f;
    if (c < F) {
      for (int j = j0; j < j1; ++j) {
        const int ent = __ldg(inv_ent + j);
        const int t = ent / L, l = ent - t * L;
        acc += __ldg(dx + ((size_t)l * T + t) * E + c);
      }
    }
    float* o = dp + (size_t)n * ldp + c;
    *o = (accumulate && c < F) ? *o + acc : acc;   // padding columns (c >= F) are zeroed
  }
}

// vectorised variant: one CTA per atom, one float4 column group per thread, incidence loop unrolled by 4 so
// that several 16-byte loads are in flight per thread (needs E % 4 == 0, ldp % 4 == 0, 16-byte aligned bases)
__global__ void __launch_bounds__(128) tuple_gather_bwd_vec_kernel(const float* __restrict__ dx, const int* __restrict__ inv_ptr,
                                                                   const int* __restrict__ inv_ent, float* __restrict__ dp,
                                                                   int ldp, int T, int L, int F, int E, int accumulate) {
  pdl_trigger();
  const int n = blockIdx.x;
  const int j0 = __ldg(inv_ptr + n), j1 = __ldg(inv_ptr + n + 1);
  const int nv = ldp >> 2, ev = E >> 2;
  const float4* dx4 = reinterpret_cast<const float4*>(dx);
  for (int c = threadIdx.x; c < nv; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c * 4 < F) {
      int j = j0;
      for (; j + 3 < j1; j += 4) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int ent = __ldg(inv_ent + j + u);
          const int t = ent / L, l = ent - t * L;
          v[u] = __ldg(dx4 + ((size_t)l * T + t) * ev + c);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
      }
      for (; j < j1; ++j) {
        const int ent = __ldg(inv_ent + j);
        const int t = ent / L, l = ent - t * L;
        const float4 v = __ldg(dx4 + ((size_t)l * T + t) * ev + c);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      // columns >= F of the last group carry the positional encoding (not a function of p): zero gradient
      if (c * 4 + 3 >= F) {
        if (c * 4 + 1 >= F) acc.y = 0.f;
        if (c * 4 + 2 >= F) acc.z = 0.f;
        acc.w = 0.f;
      }
    }
    float4* o = reinterpret_cast<float4*>(dp + (size_t)n * ldp) + c;
    if (accumulate) {
      const float4 old = *o;
      // padding columns (>= F) are zeroed, as in the scalar kernel
      acc.x += old.x;
      acc.y = (c * 4 + 1 < F) ? acc.y + old.y : 0.f;
      acc.z = (c * 4 + 2 < F) ? acc.z + old.z : 0.f;
      acc.w = (c * 4 + 3 < F) ? acc.w + old.w : 0.f;
      if (c * 4 >= F) acc.x = 0.f;
    }
    *o = acc;
  }
}

__global__ void __launch_bounds__(256) perm_concat_fwd_kernel(const float* __restrict__ x, float* __restrict__ s,
                                                              gb_perms perms, int T, int L, int E) {
  pdl_trigger();
  // one thread per float4 of the output [n_perm*T, L*E]
  const int nv = E >> 2;
  const long long total = (long long)perms.n_perm * T * L * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % nv);
    long long r = i / nv;
    const int j = (int)(r % L);
    r /= L;
    const int t = (int)(r % T), p = (int)(r / T);
    const int l = perms.perm[p][j];
    reinterpret_cast<float4*>(s)[i] = __ldg(reinterpret_cast<const float4*>(x) + ((size_t)l * T + t) * nv + c);
  }
}

__global__ void __launch_bounds__(256) perm_concat_bwd_kernel(const float* __restrict__ ds, float* __restrict__ dx,
                                                              gb_perms perms, int T, int L, int E) {
  pdl_trigger();
  // one thread per float4 of dx [L*T, E]: sum over permutations of the slot that read position l
  const int nv = E >> 2;
  const long long total = (long long)L * T * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % nv);
    long long r = i / nv;
    const int t = (int)(r % T), l = (int)(r / T);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < perms.n_perm; ++p) {
      for (int j = 0; j < L; ++j) {
        if (perms.perm[p][j] == l) {
          float4 v = __ldg(reinterpret_cast<const float4*>(ds) + (((size_t)p * T + t) * L + j) * nv + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
    }
    reinterpret_cast<float4*>(dx)[i] = acc;
  }
}

}  // namespace gb

using namespace gb;

extern "C" int grappa_b200_tuple_gather_fwd(const float* p, int32_t ldp, const int32_t* idx, const float* pe, float* x,
                                            int32_t T, int32_t L, int32_t F, int32_t E, void* stream_) {
  GB_REQUIRE(T >= 0 && L >= 1 && L <= 4 && F >= 0 && F <= E && ldp >= F, "tuple_gather_fwd: bad shape T=%d L=%d F=%d E=%d ldp=%d", T, L, F, E, ldp);
  if (T == 0) return GB_OK;
  GB_REQUIRE(p && idx && x, "tuple_gather_fwd: NULL pointer");
  const long long rows = (long long)L * T;
  tuple_gather_fwd_kernel<<<(int)((rows + 7) / 8), 256, 0, (cudaStream_t)stream_>>>(p, ldp, idx, pe, x, T, L, F, E);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_tuple_gather_bwd(const float* dx, const int32_t* inv_ptr, const int32_t* inv_ent, float* dp,
                                            int32_t ldp, int32_t n_atoms, int32_t T, int32_t L, int32_t F, int32_t E,
                                            int32_t accumulate, void* stream_) {
  GB_REQUIRE(T >= 0 && L >= 1 && L <= 4 && F >= 0 && F <= E && ldp >= F, "tuple_gather_bwd: bad shape");
  if (n_atoms == 0) return GB_OK;
  GB_REQUIRE(inv_ptr && dp && (T == 0 || (dx && inv_ent)), "tuple_gather_bwd: NULL pointer");
  const bool vec = (E % 4 == 0) && (ldp % 4 == 0) && (((uintptr_t)dx & 15) == 0) && (((uintptr_t)dp & 15) == 0);
  if (vec)
    tuple_gather_bwd_vec_kernel<<<n_atoms, 128, 0, (cudaStream_t)stream_>>>(dx, inv_ptr, inv_ent, dp, ldp, T, L, F, E, accumulate);
  else
    tuple_gather_bwd_kernel<<<(n_atoms + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(dx, inv_ptr, inv_ent, dp, ldp, n_atoms, T, L, F, E, accumulate);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

static int check_perms(const char* who, const gb_perms* perms, int T, int L, int E) {
  GB_REQUIRE(perms != nullptr, "%s: perms is NULL", who);
  GB_REQUIRE(perms->n_perm >= 1 && perms->n_perm <= 6, "%s: n_perm must be 1..6", who);
  GB_REQUIRE(T >= 0 && L >= 1 && L <= 4 && E > 0 && E % 4 == 0, "%s: bad shape T=%d L=%d E=%d", who, T, L, E);
  for (int p = 0; p < perms->n_perm; ++p)
    for (int j = 0; j < L; ++j) GB_REQUIRE(perms->perm[p][j] >= 0 && perms->perm[p][j] < L, "%s: bad permutation", who);
  return GB_OK;
}

extern "C" int grappa_b200_perm_concat_fwd(const float* x, float* s, const gb_perms* perms, int32_t T, int32_t L,
                                           int32_t E, void* stream_) {
  int rc = check_perms("perm_concat_fwd", perms, T, L, E);
  if (rc) return rc;
  if (T == 0) return GB_OK;
  GB_REQUIRE(x && s, "perm_concat_fwd: NULL pointer");
  long long total = (long long)perms->n_perm * T * L * (E / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  perm_concat_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(x, s, *perms, T, L, E);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_perm_concat_bwd(const float* ds, float* dx, const gb_perms* perms, int32_t T, int32_t L,
                                           int32_t E, void* stream_) {
  int rc = check_perms("perm_concat_bwd", perms, T, L, E);
  if (rc) return rc;
  if (T == 0) return GB_OK;
  GB_REQUIRE(ds && dx, "perm_concat_bwd: NULL pointer");
  long long total = (long long)T * L * (E / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 16) blocks = sm_count() * 16;
  perm_concat_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(ds, dx, *perms, T, L, E);
  GB_CHECK_LAUNCH();
  return GB_OK;
}
