// K4: graph attention over bonded neighbours, CSR-segmented by destination atom.
//
// Replaces the three DGL kernels behind DotGatConv (SDDMM u_dot_v, edge_softmax, SpMM u_mul_e + sum;
// call site reference models/graph_attention.py:283) and their autograd with
//   fwd : one warp per destination atom, neighbour rows gathered with 16-byte loads, online softmax in
//         registers, fp32 accumulation; saves the attention weights alpha[E,H]
//   bwd : pass 1 (per destination) softmax backward -> ds[E,H]; pass 2 (per atom) pure gather over the
//         symmetric CSR using the reverse-edge table -- deterministic, no atomics.
// Row layout: ft[n, H*D]; a lane owns float4 chunks c = lane + 32*it; the D/4 lanes of one head reduce
// their partial dot products with segmented shuffles.
#include "common.cuh"

namespace gb {

__device__ __forceinline__ float seg_sum(float v, int gs) {
  for (int o = gs >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

__global__ void __launch_bounds__(256) edge_attn_fwd_kernel(const float* __restrict__ ft, const int* __restrict__ indptr,
                                                            const int* __restrict__ esrc, float* __restrict__ out,
                                                            float* __restrict__ alpha, int n_nodes, int H, int D) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  // one warp per (atom, group of 32 float4 columns): the column groups are independent (a head never straddles one),
  // and a warp per atom walking its 4 groups one after the other left the 1,664-atom GNN kernels latency-bound
  const int nchunks = (H * D) >> 2, gs = D >> 2;
  const int ncg = (nchunks + 31) >> 5;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int v = wid / ncg;
  if (v >= n_nodes) return;
  const float scale = rsqrtf((float)D);
  const int e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const float4* ft4 = reinterpret_cast<const float4*>(ft);
  const bool leader = (lane & (gs - 1)) == 0;
  {
    const int c0 = (wid - v * ncg) << 5;
    const int c = c0 + lane;
    const bool valid = c < nchunks;
    const int h = valid ? (c << 2) / D : 0;
    const float4 q = valid ? __ldg(ft4 + (size_t)v * nchunks + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    float m = -INFINITY, l = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const int u = __ldg(esrc + e);
      const float4 k = valid ? __ldg(ft4 + (size_t)u * nchunks + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float s = seg_sum(dot4(q, k), gs) * scale;
      const float mn = fmaxf(m, s);
      const float corr = expf(m - mn), p = expf(s - mn);
      l = l * corr + p;
      acc.x = acc.x * corr + p * k.x;
      acc.y = acc.y * corr + p * k.y;
      acc.z = acc.z * corr + p * k.z;
      acc.w = acc.w * corr + p * k.w;
      m = mn;
      if (valid && leader && alpha) alpha[(size_t)e * H + h] = s;
    }
    const float il = l > 0.f ? 1.f / l : 0.f;
    if (valid)
      reinterpret_cast<float4*>(out)[(size_t)v * nchunks + c] = make_float4(acc.x * il, acc.y * il, acc.z * il, acc.w * il);
    if (valid && leader && alpha)
      for (int e = e0; e < e1; ++e) alpha[(size_t)e * H + h] = expf(alpha[(size_t)e * H + h] - m) * il;
  }
}

// pass 1: ds[e,h] = alpha_e (dalpha_e - sum_e' alpha_e' dalpha_e') / sqrt(D),  dalpha_e = <dout_v, ft_u>
__global__ void __launch_bounds__(256) edge_attn_bwd1_kernel(const float* __restrict__ ft, const float* __restrict__ alpha,
                                                             const float* __restrict__ dout, const int* __restrict__ indptr,
                                                             const int* __restrict__ esrc, float* __restrict__ ds,
                                                             int n_nodes, int H, int D) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int nchunks = (H * D) >> 2, gs = D >> 2;
  const int ncg = (nchunks + 31) >> 5;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int v = wid / ncg;
  if (v >= n_nodes) return;
  const float scale = rsqrtf((float)D);
  const int e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const float4* ft4 = reinterpret_cast<const float4*>(ft);
  const bool leader = (lane & (gs - 1)) == 0;
  {
    const int c0 = (wid - v * ncg) << 5;
    const int c = c0 + lane;
    const bool valid = c < nchunks;
    const int h = valid ? (c << 2) / D : 0;
    const float4 g = valid ? __ldg(reinterpret_cast<const float4*>(dout) + (size_t)v * nchunks + c)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    float dsum = 0.f;
    for (int e = e0; e < e1; ++e) {
      const int u = __ldg(esrc + e);
      const float4 k = valid ? __ldg(ft4 + (size_t)u * nchunks + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float da = seg_sum(dot4(g, k), gs);
      const float a = valid ? __ldg(alpha + (size_t)e * H + h) : 0.f;
      dsum += a * da;
      if (valid && leader) ds[(size_t)e * H + h] = da;
    }
    if (valid && leader)
      for (int e = e0; e < e1; ++e)
        ds[(size_t)e * H + h] = __ldg(alpha + (size_t)e * H + h) * (ds[(size_t)e * H + h] - dsum) * scale;
  }
}

// pass 2: dft_u = sum over in-edges e = (w -> u), r = reverse(e) = (u -> w):
//           (ds_e + ds_r) * ft_w + alpha_r * dout_w
__global__ void __launch_bounds__(256) edge_attn_bwd2_kernel(const float* __restrict__ ft, const float* __restrict__ alpha,
                                                             const float* __restrict__ dout, const float* __restrict__ ds,
                                                             const int* __restrict__ indptr, const int* __restrict__ esrc,
                                                             const int* __restrict__ erev, float* __restrict__ dft,
                                                             int n_nodes, int H, int D) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int nchunks = (H * D) >> 2;
  const int ncg = (nchunks + 31) >> 5;
  const int wid = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int u = wid / ncg;
  if (u >= n_nodes) return;
  const int e0 = __ldg(indptr + u), e1 = __ldg(indptr + u + 1);
  const float4* ft4 = reinterpret_cast<const float4*>(ft);
  const float4* do4 = reinterpret_cast<const float4*>(dout);
  for (int c = ((wid - u * ncg) << 5) + lane; c < nchunks; c += nchunks) {
    const int h = (c << 2) / D;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const int w = __ldg(esrc + e), r = __ldg(erev + e);
      const float4 fw = __ldg(ft4 + (size_t)w * nchunks + c);
      const float4 gw = __ldg(do4 + (size_t)w * nchunks + c);
      const float sc = ds[(size_t)e * H + h] + ds[(size_t)r * H + h];
      const float ar = __ldg(alpha + (size_t)r * H + h);
      acc.x += sc * fw.x + ar * gw.x;
      acc.y += sc * fw.y + ar * gw.y;
      acc.z += sc * fw.z + ar * gw.z;
      acc.w += sc * fw.w + ar * gw.w;
    }
    reinterpret_cast<float4*>(dft)[(size_t)u * nchunks + c] = acc;
  }
}

static int check_shape(const char* who, int n, int H, int D) {
  GB_REQUIRE(n >= 0 && H > 0 && D > 0, "%s: bad shape", who);
  GB_REQUIRE(D % 4 == 0 && D <= 128 && ((D / 4) & (D / 4 - 1)) == 0,
             "%s: head dim must be 4, 8, 16, 32, 64 or 128 (got %d)", who, D);
  return GB_OK;
}

// SAGEConv('mean') aggregation (grappa-1.0 ResidualConvBlock, reference models/graph_attention.py:314-415 with
// dgl.nn.SAGEConv): out[v] = 1/deg(v) * sum_{u in N(v)} x[u]  (mode 0, forward), and its transpose on the symmetric bonded
// graph  out[u] = sum_{v in N(u)} x[v] / deg(v)  (mode 1, backward) -- both pure gathers over the CSR by destination: one
// warp per atom, 16-byte loads, fixed summation order, no atomics.
__global__ void __launch_bounds__(256) neighbor_mean_kernel(const float* __restrict__ x, int ldx, const int* __restrict__ indptr,
                                                            const int* __restrict__ esrc, float* __restrict__ out, int ldo,
                                                            int n_nodes, int width, int mode) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int v = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (v >= n_nodes) return;
  const int e0 = __ldg(indptr + v), e1 = __ldg(indptr + v + 1);
  const float inv_own = e1 > e0 ? 1.f / (float)(e1 - e0) : 0.f;
  for (int c = lane * 4; c < width; c += 128) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = e0; e < e1; ++e) {
      const int u = __ldg(esrc + e);
      const float4 t = __ldg(reinterpret_cast<const float4*>(x + (size_t)u * ldx + c));
      float w = 1.f;
      if (mode == 1) {
        const int du = __ldg(indptr + u + 1) - __ldg(indptr + u);
        w = du > 0 ? 1.f / (float)du : 0.f;
      }
      acc.x += w * t.x; acc.y += w * t.y; acc.z += w * t.z; acc.w += w * t.w;
    }
    if (mode == 0) { acc.x *= inv_own; acc.y *= inv_own; acc.z *= inv_own; acc.w *= inv_own; }
    *reinterpret_cast<float4*>(out + (size_t)v * ldo + c) = acc;
  }
}

}  // namespace gb

using namespace gb;

extern "C" int grappa_b200_neighbor_mean(const float* x, int32_t ldx, const int32_t* indptr, const int32_t* esrc, float* out,
                                         int32_t ldo, int32_t n_nodes, int32_t width, int32_t mode, void* stream_) {
  GB_REQUIRE(n_nodes >= 0 && width > 0 && width % 4 == 0, "neighbor_mean: width must be a positive multiple of 4");
  GB_REQUIRE(ldx >= width && ldo >= width && ldx % 4 == 0 && ldo % 4 == 0, "neighbor_mean: row pitches must be multiples of 4 floats");
  GB_REQUIRE(mode == 0 || mode == 1, "neighbor_mean: mode is 0 (mean over in-neighbours) or 1 (its transpose)");
  if (n_nodes == 0) return GB_OK;
  GB_REQUIRE(x && indptr && esrc && out, "neighbor_mean: NULL pointer");
  neighbor_mean_kernel<<<(n_nodes + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(x, ldx, indptr, esrc, out, ldo, n_nodes, width, mode);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_edge_attention_fwd(const float* ft, const int32_t* indptr, const int32_t* esrc, float* out,
                                              float* alpha, int32_t n_nodes, int32_t heads, int32_t dim, void* stream_) {
  int rc = check_shape("edge_attention_fwd", n_nodes, heads, dim);
  if (rc) return rc;
  if (n_nodes == 0) return GB_OK;
  GB_REQUIRE(ft && indptr && esrc && out, "edge_attention_fwd: NULL pointer");
  const int ncg = (heads * dim / 4 + 31) / 32;
  edge_attn_fwd_kernel<<<(n_nodes * ncg + 7) / 8, 256, 0, (cudaStream_t)stream_>>>(ft, indptr, esrc, out, alpha, n_nodes, heads, dim);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_edge_attention_bwd(const float* ft, const float* alpha, const float* dout,
                                              const int32_t* indptr, const int32_t* esrc, const int32_t* erev, float* ds,
                                              float* dft, int32_t n_nodes, int32_t heads, int32_t dim, void* stream_) {
  int rc = check_shape("edge_attention_bwd", n_nodes, heads, dim);
  if (rc) return rc;
  if (n_nodes == 0) return GB_OK;
  GB_REQUIRE(ft && alpha && dout && indptr && esrc && erev && ds && dft, "edge_attention_bwd: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int ncg = (heads * dim / 4 + 31) / 32;
  edge_attn_bwd1_kernel<<<(n_nodes * ncg + 7) / 8, 256, 0, stream>>>(ft, alpha, dout, indptr, esrc, ds, n_nodes, heads, dim);
  GB_CHECK_LAUNCH();
  edge_attn_bwd2_kernel<<<(n_nodes * ncg + 7) / 8, 256, 0, stream>>>(ft, alpha, dout, ds, indptr, esrc, erev, dft, n_nodes, heads, dim);
  GB_CHECK_LAUNCH();
  return GB_OK;
}
