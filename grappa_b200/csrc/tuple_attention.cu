// K10 core: multi-head self-attention over the L <= 4 atoms of each tuple, fully in registers.
//
// Replaces the scaled-dot-product part of torch.nn.MultiheadAttention as used by DottedAttWithMLP
// (reference models/network_utils.py:105,122; semantics SURVEY.md appendix A.1): tokens are rows
// l*T + t of qkv[L*T, 3E] (q | k | v), heads are contiguous head_dim slices, softmax over the L keys.
// A group of head_dim/4 lanes owns one (tuple, head): every lane keeps a float4 of q, k, v for all L
// tokens; the L x L score matrix is formed with segmented shuffle reductions.  Backward recomputes
// the probabilities instead of storing them.
#include "common.cuh"

namespace gb {

__device__ __forceinline__ float seg_sum_t(float v, int gs) {
  for (int o = gs >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot4t(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 fma4(float s, float4 a, float4 b) {
  return make_float4(fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z), fmaf(s, a.w, b.w));
}

template <int L>
__global__ void __launch_bounds__(256) tuple_attn_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ out, int T,
                                                             int heads, int HD) {
  pdl_trigger();
  const int gs = HD >> 2;
  const int groups_per_warp = 32 / gs;
  const int lane = threadIdx.x & 31;
  const long long warp = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long item = warp * groups_per_warp + lane / gs;
  const long long n_items = (long long)T * heads;
  const bool valid = item < n_items;
  const int t = valid ? (int)(item / heads) : 0;
  const int h = valid ? (int)(item % heads) : 0;
  const int E = heads * HD;
  const int col = h * HD + (lane & (gs - 1)) * 4;
  const float scale = rsqrtf((float)HD);
  float4 q[L], k[L], v[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float* row = qkv + ((size_t)l * T + t) * 3 * E + col;
    if (valid) {
      q[l] = __ldg(reinterpret_cast<const float4*>(row));
      k[l] = __ldg(reinterpret_cast<const float4*>(row + E));
      v[l] = __ldg(reinterpret_cast<const float4*>(row + 2 * E));
    } else {
      q[l] = k[l] = v[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int l = 0; l < L; ++l) {
    float s[L], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      s[j] = seg_sum_t(dot4t(q[l], k[j]), gs) * scale;
      mx = fmaxf(mx, s[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) { s[j] = expf(s[j] - mx); den += s[j]; }
    const float id = 1.f / den;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < L; ++j) o = fma4(s[j] * id, v[j], o);
    if (valid) *reinterpret_cast<float4*>(out + ((size_t)l * T + t) * E + col) = o;
  }
}

template <int L>
__global__ void __launch_bounds__(256) tuple_attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dout,
                                                             float* __restrict__ dqkv, int T, int heads, int HD) {
  pdl_trigger();
  const int gs = HD >> 2;
  const int groups_per_warp = 32 / gs;
  const int lane = threadIdx.x & 31;
  const long long warp = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long item = warp * groups_per_warp + lane / gs;
  const long long n_items = (long long)T * heads;
  const bool valid = item < n_items;
  const int t = valid ? (int)(item / heads) : 0;
  const int h = valid ? (int)(item % heads) : 0;
  const int E = heads * HD;
  const int col = h * HD + (lane & (gs - 1)) * 4;
  const float scale = rsqrtf((float)HD);
  float4 q[L], k[L], v[L], go[L], dq[L], dk[L], dv[L];
#pragma unroll
  for (int l = 0; l < L; ++l) {
    const float* row = qkv + ((size_t)l * T + t) * 3 * E + col;
    if (valid) {
      q[l] = __ldg(reinterpret_cast<const float4*>(row));
      k[l] = __ldg(reinterpret_cast<const float4*>(row + E));
      v[l] = __ldg(reinterpret_cast<const float4*>(row + 2 * E));
      go[l] = __ldg(reinterpret_cast<const float4*>(dout + ((size_t)l * T + t) * E + col));
    } else {
      q[l] = k[l] = v[l] = go[l] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    dq[l] = dk[l] = dv[l] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int l = 0; l < L; ++l) {
    float p[L], dp[L], mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      p[j] = seg_sum_t(dot4t(q[l], k[j]), gs) * scale;
      mx = fmaxf(mx, p[j]);
    }
    float den = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) { p[j] = expf(p[j] - mx); den += p[j]; }
    const float id = 1.f / den;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < L; ++j) {
      p[j] *= id;
      dp[j] = seg_sum_t(dot4t(go[l], v[j]), gs);
      dot += p[j] * dp[j];
      dv[j] = fma4(p[j], go[l], dv[j]);
    }
#pragma unroll
    for (int j = 0; j < L; ++j) {
      const float dsv = p[j] * (dp[j] - dot) * scale;
      dq[l] = fma4(dsv, k[j], dq[l]);
      dk[j] = fma4(dsv, q[l], dk[j]);
    }
  }
  if (valid) {
#pragma unroll
    for (int l = 0; l < L; ++l) {
      float* row = dqkv + ((size_t)l * T + t) * 3 * E + col;
      *reinterpret_cast<float4*>(row) = dq[l];
      *reinterpret_cast<float4*>(row + E) = dk[l];
      *reinterpret_cast<float4*>(row + 2 * E) = dv[l];
    }
  }
}

static int check(const char* who, int T, int L, int heads, int HD) {
  GB_REQUIRE(T >= 0 && heads > 0, "%s: bad shape", who);
  GB_REQUIRE(L >= 1 && L <= 4, "%s: tuple length must be 1..4 (got %d)", who, L);
  GB_REQUIRE(HD % 4 == 0 && HD <= 128 && ((HD / 4) & (HD / 4 - 1)) == 0,
             "%s: head_dim must be 4, 8, 16, 32, 64 or 128 (got %d)", who, HD);
  return GB_OK;
}

}  // namespace gb

using namespace gb;

extern "C" int grappa_b200_tuple_attention_fwd(const float* qkv, float* out, int32_t T, int32_t L, int32_t heads,
                                               int32_t head_dim, void* stream_) {
  int rc = check("tuple_attention_fwd", T, L, heads, head_dim);
  if (rc) return rc;
  if (T == 0) return GB_OK;
  GB_REQUIRE(qkv && out, "tuple_attention_fwd: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long items = (long long)T * heads;
  const int per_block = 8 * (32 / (head_dim / 4));
  const int blocks = (int)((items + per_block - 1) / per_block);
  switch (L) {
    case 1: tuple_attn_fwd_kernel<1><<<blocks, 256, 0, stream>>>(qkv, out, T, heads, head_dim); break;
    case 2: tuple_attn_fwd_kernel<2><<<blocks, 256, 0, stream>>>(qkv, out, T, heads, head_dim); break;
    case 3: tuple_attn_fwd_kernel<3><<<blocks, 256, 0, stream>>>(qkv, out, T, heads, head_dim); break;
    default: tuple_attn_fwd_kernel<4><<<blocks, 256, 0, stream>>>(qkv, out, T, heads, head_dim); break;
  }
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_tuple_attention_bwd(const float* qkv, const float* dout, float* dqkv, int32_t T, int32_t L,
                                               int32_t heads, int32_t head_dim, void* stream_) {
  int rc = check("tuple_attention_bwd", T, L, heads, head_dim);
  if (rc) return rc;
  if (T == 0) return GB_OK;
  GB_REQUIRE(qkv && dout && dqkv, "tuple_attention_bwd: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const long long items = (long long)T * heads;
  const int per_block = 8 * (32 / (head_dim / 4));
  const int blocks = (int)((items + per_block - 1) / per_block);
  switch (L) {
    case 1: tuple_attn_bwd_kernel<1><<<blocks, 256, 0, stream>>>(qkv, dout, dqkv, T, heads, head_dim); break;
    case 2: tuple_attn_bwd_kernel<2><<<blocks, 256, 0, stream>>>(qkv, dout, dqkv, T, heads, head_dim); break;
    case 3: tuple_attn_bwd_kernel<3><<<blocks, 256, 0, stream>>>(qkv, dout, dqkv, T, heads, head_dim); break;
    default: tuple_attn_bwd_kernel<4><<<blocks, 256, 0, stream>>>(qkv, dout, dqkv, T, heads, head_dim); break;
  }
  GB_CHECK_LAUNCH();
  return GB_OK;
}
