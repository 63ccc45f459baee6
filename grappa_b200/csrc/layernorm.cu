// LayerNorm forward / backward (dx) and the deterministic column reductions that give the parameter
// gradients of LayerNorm (gamma, beta) and of every Linear bias.
// torch.nn.LayerNorm semantics: biased variance, eps inside the sqrt, affine
// (reference models/graph_attention.py:258,265; models/network_utils.py:38,100).
#include "common.cuh"

namespace gb {

// one warp per row; the row lives in registers (VPT float4 per lane), two-pass mean / variance
template <int VPT>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, float* __restrict__ y,
                                                            float* __restrict__ mean, float* __restrict__ rstd, int rows,
                                                            int cols, float eps) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols >> 2;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * cols);
  float4 v[VPT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    v[i] = c < nvec ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mu = warp_sum(s) / cols;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float a = v[i].x - mu, b = v[i].y - mu, c2 = v[i].z - mu, d = v[i].w - mu;
      q += a * a + b * b + c2 * c2 + d * d;
    }
  }
  const float rs = rsqrtf(warp_sum(q) / cols + eps);
  if (lane == 0) {
    if (mean) mean[row] = mu;
    if (rstd) rstd[row] = rs;
  }
  float4* yr = reinterpret_cast<float4*>(y + (size_t)row * cols);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      float4 b = __ldg(reinterpret_cast<const float4*>(beta) + c);
      float4 o;
      o.x = (v[i].x - mu) * rs * g.x + b.x;
      o.y = (v[i].y - mu) * rs * g.y + b.y;
      o.z = (v[i].z - mu) * rs * g.z + b.z;
      o.w = (v[i].w - mu) * rs * g.w + b.w;
      yr[c] = o;
    }
  }
}

// dx = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy * gamma
template <int VPT>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const float* __restrict__ gamma, float* __restrict__ dx,
                                                            int rows, int cols) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = cols >> 2;
  const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * cols);
  const float4* dr = reinterpret_cast<const float4*>(dy + (size_t)row * cols);
  float4 xh[VPT], g[VPT];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      float4 xv = __ldg(xr + c), dv = __ldg(dr + c), gm = __ldg(reinterpret_cast<const float4*>(gamma) + c);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      g[i] = make_float4(dv.x * gm.x, dv.y * gm.y, dv.z * gm.z, dv.w * gm.w);
      s1 += g[i].x + g[i].y + g[i].z + g[i].w;
      s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
    } else {
      xh[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  s1 = warp_sum(s1) / cols;
  s2 = warp_sum(s2) / cols;
  float4* o = reinterpret_cast<float4*>(dx + (size_t)row * cols);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      o[c] = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                         rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
    }
  }
}

// Column reduction, ONE kernel: block = 32 columns x 8 row lanes, grid.y row slices.  Every block writes its
// partial sums; the last block to finish a column group (ticket counter, self-resetting) adds the partials in
// slice order, so the result is deterministic and no second launch is needed.
__global__ void __launch_bounds__(256) col_reduce_kernel(const float* __restrict__ dy, int ld, const float* __restrict__ x,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         float* __restrict__ part_sum, float* __restrict__ part_xhat,
                                                         unsigned int* __restrict__ tickets, float* __restrict__ out_sum,
                                                         float* __restrict__ out_xhat, int rows, int cols, int accumulate) {
  pdl_trigger();
  __shared__ float sh[2][8][33];
  __shared__ bool is_last;
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const int slices = gridDim.y;
  float a = 0.f, b = 0.f;
  if (col < cols) {
    const int stride = slices * 8;
    int r = blockIdx.y * 8 + ry;
    // 4 independent rows in flight per thread
    for (; r + 3 * stride < rows; r += 4 * stride) {
      float d[4], xv[4], mu[4], rs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) d[u] = __ldg(dy + (size_t)(r + u * stride) * ld + col);
      if (x) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          xv[u] = __ldg(x + (size_t)(r + u * stride) * cols + col);
          mu[u] = __ldg(mean + r + u * stride);
          rs[u] = __ldg(rstd + r + u * stride);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        a += d[u];
        if (x) b += d[u] * (xv[u] - mu[u]) * rs[u];
      }
    }
    for (; r < rows; r += stride) {
      const float d = __ldg(dy + (size_t)r * ld + col);
      a += d;
      if (x) b += d * (__ldg(x + (size_t)r * cols + col) - __ldg(mean + r)) * __ldg(rstd + r);
    }
  }
  sh[0][ry][cx] = a;
  sh[1][ry][cx] = b;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa += sh[0][i][cx]; sb += sh[1][i][cx]; }
    if (slices == 1) {
      if (out_sum) out_sum[col] = accumulate ? out_sum[col] + sa : sa;
      if (x && out_xhat) out_xhat[col] = accumulate ? out_xhat[col] + sb : sb;
    } else {
      __stcg(part_sum + (size_t)blockIdx.y * cols + col, sa);
      if (x) __stcg(part_xhat + (size_t)blockIdx.y * cols + col, sb);
    }
  }
  if (slices == 1) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(tickets + blockIdx.x, 1u) == (unsigned int)(slices - 1);
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  a = 0.f; b = 0.f;
  if (col < cols) {
    for (int s2 = ry; s2 < slices; s2 += 8) {
      a += __ldcg(part_sum + (size_t)s2 * cols + col);
      if (x) b += __ldcg(part_xhat + (size_t)s2 * cols + col);
    }
  }
  sh[0][ry][cx] = a;
  sh[1][ry][cx] = b;
  __syncthreads();
  if (ry == 0 && col < cols) {
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { sa += sh[0][i][cx]; sb += sh[1][i][cx]; }
    if (out_sum) out_sum[col] = accumulate ? out_sum[col] + sa : sa;
    if (x && out_xhat) out_xhat[col] = accumulate ? out_xhat[col] + sb : sb;
  }
  if (threadIdx.x == 0) tickets[blockIdx.x] = 0u;   // ready for the next launch on this workspace
}

// ------------------------------------------------------------------------------------------------
// Fused backward kernels that also emit per-CTA column partial sums (parameter gradients), finalised later
// by ONE finalize_colsums launch per stage (deterministic: fixed row->CTA assignment, partials added in CTA order).
// ------------------------------------------------------------------------------------------------
// LayerNorm: dx as layernorm_bwd_kernel + partial[b][0][c] = sum_rows dy, partial[b][1][c] = sum_rows dy * xhat
template <int VPT>
__global__ void __launch_bounds__(256) layernorm_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, float* __restrict__ dx,
                                                                  float* __restrict__ partial, int rows, int cols,
                                                                  int rows_per_cta) {
  pdl_trigger();
  extern __shared__ float sm[];   // [8 warps][2][cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nvec = cols >> 2;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows, r_begin + rows_per_cta);
  float4 pb[VPT], pg[VPT];
  float4 gm[VPT];
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    pb[i] = pg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c = lane + i * 32;
    gm[i] = c < nvec ? __ldg(reinterpret_cast<const float4*>(gamma) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int row = r_begin + warp; row < r_end; row += 8) {
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * cols);
    const float4* dr = reinterpret_cast<const float4*>(dy + (size_t)row * cols);
    float4 xh[VPT], g[VPT];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = lane + i * 32;
      if (c < nvec) {
        const float4 xv = __ldg(xr + c), dv = __ldg(dr + c);
        xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
        g[i] = make_float4(dv.x * gm[i].x, dv.y * gm[i].y, dv.z * gm[i].z, dv.w * gm[i].w);
        s1 += g[i].x + g[i].y + g[i].z + g[i].w;
        s2 += g[i].x * xh[i].x + g[i].y * xh[i].y + g[i].z * xh[i].z + g[i].w * xh[i].w;
        pb[i].x += dv.x; pb[i].y += dv.y; pb[i].z += dv.z; pb[i].w += dv.w;
        pg[i].x += dv.x * xh[i].x; pg[i].y += dv.y * xh[i].y; pg[i].z += dv.z * xh[i].z; pg[i].w += dv.w * xh[i].w;
      } else {
        xh[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    s1 = warp_sum(s1) / cols;
    s2 = warp_sum(s2) / cols;
    float4* o = reinterpret_cast<float4*>(dx + (size_t)row * cols);
#pragma unroll
    for (int i = 0; i < VPT; ++i) {
      const int c = lane + i * 32;
      if (c < nvec)
        o[c] = make_float4(rs * (g[i].x - s1 - xh[i].x * s2), rs * (g[i].y - s1 - xh[i].y * s2),
                           rs * (g[i].z - s1 - xh[i].z * s2), rs * (g[i].w - s1 - xh[i].w * s2));
    }
  }
  float4* smv = reinterpret_cast<float4*>(sm);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int c = lane + i * 32;
    if (c < nvec) {
      smv[(warp * 2 + 0) * nvec + c] = pb[i];
      smv[(warp * 2 + 1) * nvec + c] = pg[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * cols; i += 256) {     // i = set * cols + c
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) a += sm[(size_t)w * 2 * cols + i];
    partial[(size_t)blockIdx.x * 2 * cols + i] = a;
  }
}

// dx = dy * keep/(1-p) * elu'(act_out) (as act_dropout_bwd_kernel) + partial[b][c] = sum over the CTA's rows of dx
__global__ void __launch_bounds__(256) act_dropout_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ act_out,
                                                                    float* __restrict__ dx, float* __restrict__ partial,
                                                                    int rows, int cols, int rows_per_cta, uint32_t thresh,
                                                                    float inv_keep, uint64_t seed0,
                                                                    const uint64_t* __restrict__ seed_off) {
  pdl_trigger();
  __shared__ float4 sh[256];
  const uint64_t seed = seed_with_offset(seed0, seed_off);
  // Rows wider than 1024 floats are cut into gridDim.y column slices of <= 256 float4 groups (GNN feed-forward: 2048).
  const int ncg_row = cols >> 2;                                   // float4 groups per row
  const int ncg = (ncg_row + gridDim.y - 1) / gridDim.y;           // ... per slice (<= 256)
  const int cg0 = blockIdx.y * ncg;
  const int nrl = 256 / ncg;                 // row lanes
  const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg;
  const int r_begin = blockIdx.x * rows_per_cta;
  const int r_end = min(rows, r_begin + rows_per_cta);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < nrl && cg0 + cg < ncg_row) {
    for (int row = r_begin + rl; row < r_end; row += nrl) {
      const size_t i4 = (size_t)row * ncg_row + cg0 + cg;
      float4 v = __ldg(reinterpret_cast<const float4*>(dy) + i4);
      if (thresh) {
        const uint64_t base = (uint64_t)i4 * 4ull;
        const float4 d = dropout_scale4(seed, base, thresh, inv_keep);
        v.x *= d.x; v.y *= d.y; v.z *= d.z; v.w *= d.w;
      }
      if (act_out) {
        const float4 y = __ldg(reinterpret_cast<const float4*>(act_out) + i4);
        v.x *= elu1_grad_from_out(y.x); v.y *= elu1_grad_from_out(y.y);
        v.z *= elu1_grad_from_out(y.z); v.w *= elu1_grad_from_out(y.w);
      }
      reinterpret_cast<float4*>(dx)[i4] = v;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < ncg && cg0 + threadIdx.x < ncg_row) {
    float4 a = sh[threadIdx.x];
    for (int l = 1; l < nrl; ++l) {
      const float4 o = sh[l * ncg + threadIdx.x];
      a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
    }
    reinterpret_cast<float4*>(partial)[(size_t)blockIdx.x * ncg_row + cg0 + threadIdx.x] = a;
  }
}

// out[c] (+)= sum_{p < n_part} partial[p * stride + c]   for every descriptor; grid = (column chunks of 64, n_desc).
// 4 lanes share a column (partial rows p = lane, lane + 4, ...), combined in a fixed order.
__global__ void __launch_bounds__(256) finalize_colsums_kernel(gb_colsum_batch batch) {
  pdl_trigger();
  const gb_colsum_desc d = batch.desc[blockIdx.y];
  const int c = blockIdx.x * 64 + (threadIdx.x >> 2);
  const int sub = threadIdx.x & 3;
  float acc = 0.f;
  if (c < d.cols) {
    const float* p = d.partial + c;
    int i = sub;
    for (; i + 12 < d.n_part; i += 16) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(p + (size_t)(i + 4 * u) * d.stride);
#pragma unroll
      for (int u = 0; u < 4; ++u) acc += v[u];
    }
    for (; i < d.n_part; i += 4) acc += __ldg(p + (size_t)i * d.stride);
  }
  const float a1 = __shfl_down_sync(0xffffffffu, acc, 1);
  const float a2 = __shfl_down_sync(0xffffffffu, acc, 2);
  const float a3 = __shfl_down_sync(0xffffffffu, acc, 3);
  if (sub == 0 && c < d.cols) {
    const float tot = ((acc + a1) + a2) + a3;
    d.out[c] = d.accumulate ? d.out[c] + tot : tot;
  }
}

static int col_slices(int rows) {
  int s = (rows + 63) / 64;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

}  // namespace gb

using namespace gb;

#define GB_LN_DISPATCH(KERNEL, ...)                                                     \
  do {                                                                                  \
    const int nvec = cols / 4;                                                          \
    const int blocks = (rows + 7) / 8;                                                  \
    if (nvec <= 32) KERNEL<1><<<blocks, 256, 0, stream>>>(__VA_ARGS__);                 \
    else if (nvec <= 64) KERNEL<2><<<blocks, 256, 0, stream>>>(__VA_ARGS__);            \
    else if (nvec <= 128) KERNEL<4><<<blocks, 256, 0, stream>>>(__VA_ARGS__);           \
    else if (nvec <= 256) KERNEL<8><<<blocks, 256, 0, stream>>>(__VA_ARGS__);           \
    else if (nvec <= 384) KERNEL<12><<<blocks, 256, 0, stream>>>(__VA_ARGS__);          \
    else KERNEL<16><<<blocks, 256, 0, stream>>>(__VA_ARGS__);                           \
  } while (0)

extern "C" int grappa_b200_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean,
                                         float* rstd, int32_t rows, int32_t cols, float eps, void* stream_) {
  GB_REQUIRE(rows >= 0 && cols > 0, "layernorm_fwd: bad shape %d x %d", rows, cols);
  GB_REQUIRE(cols % 4 == 0 && cols <= 2048, "layernorm_fwd: cols must be a multiple of 4 and <= 2048 (got %d)", cols);
  if (rows == 0) return GB_OK;
  GB_REQUIRE(x && gamma && beta && y, "layernorm_fwd: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  GB_LN_DISPATCH(layernorm_fwd_kernel, x, gamma, beta, y, mean, rstd, rows, cols, eps);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                                         const float* gamma, float* dx, int32_t rows, int32_t cols, void* stream_) {
  GB_REQUIRE(rows >= 0 && cols > 0, "layernorm_bwd: bad shape %d x %d", rows, cols);
  GB_REQUIRE(cols % 4 == 0 && cols <= 2048, "layernorm_bwd: cols must be a multiple of 4 and <= 2048 (got %d)", cols);
  if (rows == 0) return GB_OK;
  GB_REQUIRE(dy && x && mean && rstd && gamma && dx, "layernorm_bwd: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  GB_LN_DISPATCH(layernorm_bwd_kernel, dy, x, mean, rstd, gamma, dx, rows, cols);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

// workspace layout: [256 B ticket counters (MUST be zero on first use; the kernel resets them)] [partials]
extern "C" int64_t grappa_b200_col_reduce_workspace(int32_t rows, int32_t cols) {
  return 256 + (int64_t)2 * col_slices(rows) * cols * sizeof(float);
}

extern "C" int grappa_b200_col_reduce(const float* dy, int32_t ld, const float* x, const float* mean, const float* rstd,
                                      float* out_sum, float* out_xhat, float* workspace, int32_t rows, int32_t cols,
                                      int32_t accumulate, void* stream_) {
  GB_REQUIRE(rows >= 0 && cols > 0 && ld >= cols, "col_reduce: bad shape %d x %d (ld %d)", rows, cols, ld);
  GB_REQUIRE(workspace != nullptr, "col_reduce: workspace is NULL");
  GB_REQUIRE(!x || (mean && rstd), "col_reduce: x given without mean/rstd");
  cudaStream_t stream = (cudaStream_t)stream_;
  GB_REQUIRE(cols <= 2048, "col_reduce: cols must be <= 2048 (got %d)", cols);
  if (rows == 0) {
    if (!accumulate) {
      if (out_sum) GB_CHECK_CUDA(cudaMemsetAsync(out_sum, 0, sizeof(float) * cols, stream));
      if (out_xhat && x) GB_CHECK_CUDA(cudaMemsetAsync(out_xhat, 0, sizeof(float) * cols, stream));
    }
    return GB_OK;
  }
  const int slices = col_slices(rows);
  unsigned int* tickets = reinterpret_cast<unsigned int*>(workspace);
  float* ps = workspace + 64;
  float* px = x ? ps + (size_t)slices * cols : nullptr;
  dim3 grid((cols + 31) / 32, slices);
  col_reduce_kernel<<<grid, 256, 0, stream>>>(dy, ld, x, mean, rstd, ps, px, tickets, out_sum, out_xhat, rows, cols, accumulate);
  GB_CHECK_LAUNCH();
  return GB_OK;
}


extern "C" int grappa_b200_layernorm_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd,
                                               const float* gamma, float* dx, float* partial, int32_t n_cta, int32_t rows,
                                               int32_t cols, void* stream_) {
  GB_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0 && cols <= 512, "layernorm_bwd_fused: cols must be a multiple of 4 and <= 512 (got %d x %d)", rows, cols);
  GB_REQUIRE(n_cta >= 1, "layernorm_bwd_fused: n_cta must be >= 1");
  GB_REQUIRE(dy && x && mean && rstd && gamma && dx && partial, "layernorm_bwd_fused: NULL pointer");
  cudaStream_t stream = (cudaStream_t)stream_;
  const int rpc = (rows + n_cta - 1) / n_cta;
  const size_t smem = (size_t)8 * 2 * cols * sizeof(float);
  const int nvec = cols / 4;
  if (nvec <= 32) layernorm_bwd_fused_kernel<1><<<n_cta, 256, smem, stream>>>(dy, x, mean, rstd, gamma, dx, partial, rows, cols, rpc);
  else if (nvec <= 64) layernorm_bwd_fused_kernel<2><<<n_cta, 256, smem, stream>>>(dy, x, mean, rstd, gamma, dx, partial, rows, cols, rpc);
  else layernorm_bwd_fused_kernel<4><<<n_cta, 256, smem, stream>>>(dy, x, mean, rstd, gamma, dx, partial, rows, cols, rpc);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_act_dropout_bwd_fused(const float* dy, const float* act_out, float* dx, float* partial,
                                                 int32_t n_cta, int32_t rows, int32_t cols, float p, uint64_t seed,
                                                 const uint64_t* seed_offset, void* stream_) {
  GB_REQUIRE(p >= 0.f && p < 1.f, "act_dropout_bwd_fused: p must be in [0,1)");
  GB_REQUIRE(rows > 0 && cols >= 4 && cols % 4 == 0 && cols <= 8192, "act_dropout_bwd_fused: cols must be a multiple of 4 and <= 8192 (got %d x %d)", rows, cols);
  GB_REQUIRE(n_cta >= 1 && dy && dx && partial, "act_dropout_bwd_fused: bad arguments");
  uint32_t thresh = 0;
  if (p > 0.f) {
    double t = (double)p * 4294967296.0;
    thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    if (thresh == 0) thresh = 1;
  }
  const int rpc = (rows + n_cta - 1) / n_cta;
  const int slices = (cols / 4 + 255) / 256;
  act_dropout_bwd_fused_kernel<<<dim3(n_cta, slices), 256, 0, (cudaStream_t)stream_>>>(dy, act_out, dx, partial, rows, cols, rpc, thresh, 1.f / (1.f - p), seed, seed_offset);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_finalize_colsums(const gb_colsum_batch* batch, void* stream_) {
  GB_REQUIRE(batch != nullptr && batch->n >= 0 && batch->n <= GB_COLSUM_MAX, "finalize_colsums: bad batch");
  if (batch->n == 0) return GB_OK;
  int max_cols = 0;
  for (int i = 0; i < batch->n; ++i) {
    const gb_colsum_desc& d = batch->desc[i];
    GB_REQUIRE(d.partial && d.out && d.cols > 0 && d.n_part > 0 && d.stride >= d.cols, "finalize_colsums: bad descriptor %d", i);
    if (d.cols > max_cols) max_cols = d.cols;
  }
  dim3 grid((max_cols + 63) / 64, batch->n);
  finalize_colsums_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(*batch);
  GB_CHECK_LAUNCH();
  return GB_OK;
}
