// fp32 FFMA GEMM (CUDA cores) with the fused epilogue -- the bit-faithful fp32 path used for 1e-5
// parity against the reference and for shapes the tcgen05 path cannot take (unaligned rows, K = 85).
// 128x128x8 tiles, 256 threads, 8x8 micro-tiles split 2x2 so shared-memory reads are conflict-free,
// register-staged double buffering, optional deterministic split-K through a workspace.
#include "common.cuh"
#include "gemm_epilogue.cuh"

namespace gb {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4;

// rows x K operand tile -> smem[k][row].  TR = 0: global is [rows, K] (K contiguous); TR = 1: [K, rows].
template <int TR>
struct TileLoader {
  float v[4];
  __device__ __forceinline__ void load(const float* __restrict__ g, int ld, int row0, int k0, int rows, int k_end,
                                       int tid, bool vec) {
    if (TR == 0) {
      const int r = row0 + (tid >> 1), k = k0 + (tid & 1) * 4;
      if (vec && r < rows && k + 3 < k_end) {
        float4 t = __ldg(reinterpret_cast<const float4*>(g + (size_t)r * ld + k));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (r < rows && k + i < k_end) ? __ldg(g + (size_t)r * ld + k + i) : 0.f;
      }
    } else {
      const int k = k0 + (tid >> 5), r = row0 + (tid & 31) * 4;
      if (vec && k < k_end && r + 3 < rows) {
        float4 t = __ldg(reinterpret_cast<const float4*>(g + (size_t)k * ld + r));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (k < k_end && r + i < rows) ? __ldg(g + (size_t)k * ld + r + i) : 0.f;
      }
    }
  }
  __device__ __forceinline__ void store(float (*s)[BM + PAD], int tid) const {
    if (TR == 0) {
      const int r = tid >> 1, kq = (tid & 1) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) s[kq + i][r] = v[i];
    } else {
      const int k = tid >> 5, rq = (tid & 31) * 4;
      *reinterpret_cast<float4*>(&s[k][rq]) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
};

template <int TA, int TB>
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, int M,
                                                    int N, int K, int lda, int ldb, bool vec_a, bool vec_b,
                                                    int k_chunk, float* __restrict__ partial, Epilogue ep) {
  pdl_trigger();
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_chunk;
  const int k_end = min(K, k_begin + k_chunk);
  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  TileLoader<TA> la;
  TileLoader<TB> lb;
  la.load(A, lda, m0, k_begin, M, k_end, tid, vec_a);
  lb.load(B, ldb, n0, k_begin, N, k_end, tid, vec_b);
  // TileLoader indexes rows relative to row0 through its own arithmetic: shift to tile-local on store
  la.store(As[0], tid);
  lb.store(Bs[0], tid);
  __syncthreads();
  int buf = 0;
  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    const bool more = k0 + BK < k_end;
    if (more) {
      la.load(A, lda, m0, k0 + BK, M, k_end, tid, vec_a);
      lb.load(B, ldb, n0, k0 + BK, N, k_end, tid, vec_b);
    }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      la.store(As[buf ^ 1], tid);
      lb.store(Bs[buf ^ 1], tid);
    }
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      if (partial) partial[((size_t)blockIdx.z * M + m) * N + n] = acc[i][j];
      else ep.store(acc[i][j], m, n);
    }
  }
}

__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                                                            Epilogue ep) {
  pdl_trigger();
  const size_t total = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(size_t)z * total + i];
    ep.store(s, (int)(i / N), (int)(i % N));
  }
}

// float4 variant (N % 4 == 0 and a 16-byte aligned epilogue): 4 slices in flight per thread, fixed summation order
__global__ void __launch_bounds__(256) splitk_reduce_vec_kernel(const float4* __restrict__ partial, int splits, int M, int N,
                                                                Epilogue ep) {
  pdl_trigger();
  const int nv = N >> 2;
  const size_t total = (size_t)M * nv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 3 < splits; z += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(partial + (size_t)(z + u) * total + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; z < splits; ++z) {
      const float4 v = __ldcs(partial + (size_t)z * total + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const int row = (int)(i / nv), col = (int)(i - (size_t)row * nv) * 4;
    ep.template store4<false>(s, row, col);
  }
}

// one launch for the partials of a grouped GEMM: blockIdx.y = problem
struct ReduceGroup {
  int n;
  const float4* partial[4];
  int splits[4], M[4], N[4];
  Epilogue ep[4];
};
__global__ void __launch_bounds__(256) splitk_reduce_grouped_kernel(const __grid_constant__ ReduceGroup g) {
  pdl_trigger();
  const int pi = blockIdx.y;
  const float4* __restrict__ partial = g.partial[pi];
  const int splits = g.splits[pi], N = g.N[pi];
  if (splits <= 1) return;
  const int nv = N >> 2;
  const size_t total = (size_t)g.M[pi] * nv;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 3 < splits; z += 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcs(partial + (size_t)(z + u) * total + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; z < splits; ++z) {
      const float4 v = __ldcs(partial + (size_t)z * total + i);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    const int row = (int)(i / nv), col = (int)(i - (size_t)row * nv) * 4;
    g.ep[pi].template store4<false>(s, row, col);
  }
}

static bool reduce_vec_ok(const float* partial, int N, const Epilogue& ep) {
  return (N % 4 == 0) && (((uintptr_t)partial & 15) == 0) && (((uintptr_t)ep.bias | (uintptr_t)ep.mul_elu_out |
          (uintptr_t)ep.residual | (uintptr_t)ep.C | (uintptr_t)ep.act_out) & 15) == 0 &&
         ((ep.ldc | (ep.mul_elu_out ? ep.ldm : 0) | (ep.residual ? ep.ldr : 0) | (ep.act_out ? ep.ldact : 0)) & 3) == 0;
}

int launch_splitk_reduce(const float* partial, int splits, int M, int N, const Epilogue& ep, cudaStream_t stream);

int launch_splitk_reduce_grouped(int n, const float* const* partial, const int* splits, const int* Ms, const int* Ns,
                                 const Epilogue* eps, cudaStream_t stream) {
  bool vec = n <= 4;
  size_t max_total = 0;
  for (int i = 0; i < n && vec; ++i) {
    if (splits[i] <= 1) continue;
    vec = reduce_vec_ok(partial[i], Ns[i], eps[i]);
    const size_t t = (size_t)Ms[i] * Ns[i] / 4;
    max_total = t > max_total ? t : max_total;
  }
  if (!vec) {
    for (int i = 0; i < n; ++i)
      if (splits[i] > 1) {
        int rc = launch_splitk_reduce(partial[i], splits[i], Ms[i], Ns[i], eps[i], stream);
        if (rc) return rc;
      }
    return GB_OK;
  }
  ReduceGroup g;
  g.n = n;
  for (int i = 0; i < 4; ++i) {
    g.partial[i] = i < n ? reinterpret_cast<const float4*>(partial[i]) : nullptr;
    g.splits[i] = i < n ? splits[i] : 0;
    g.M[i] = i < n ? Ms[i] : 0;
    g.N[i] = i < n ? Ns[i] : 0;
    if (i < n) g.ep[i] = eps[i]; else g.ep[i] = eps[0];
  }
  int bx = (int)((max_total + 255) / 256);
  const int cap = sm_count() * 8 / n > 1 ? sm_count() * 8 / n : 1;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  splitk_reduce_grouped_kernel<<<dim3(bx, n), 256, 0, stream>>>(g);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

int launch_splitk_reduce(const float* partial, int splits, int M, int N, const Epilogue& ep, cudaStream_t stream) {
  const bool vec = (N % 4 == 0) && (((uintptr_t)partial & 15) == 0) && (((uintptr_t)ep.bias | (uintptr_t)ep.mul_elu_out |
                    (uintptr_t)ep.residual | (uintptr_t)ep.C | (uintptr_t)ep.act_out) & 15) == 0 &&
                   ((ep.ldc | (ep.mul_elu_out ? ep.ldm : 0) | (ep.residual ? ep.ldr : 0) | (ep.act_out ? ep.ldact : 0)) & 3) == 0;
  size_t total = (size_t)M * N / (vec ? 4 : 1);
  int blocks = (int)((total + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  if (vec)
    splitk_reduce_vec_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(partial), splits, M, N, ep);
  else
    splitk_reduce_kernel<<<blocks, 256, 0, stream>>>(partial, splits, M, N, ep);
  GB_CHECK_LAUNCH();
  return GB_OK;
}


// ---- skinny shapes ---------------------------------------------------------------------------------
// The last symmetriser layer has 2 (bond / angle) or 6 (gated torsion) outputs: its forward GEMM has N <= 8, its
// input gradient K <= 8 and its weight gradient M <= 8.  A 128x128 tile wastes > 90 % of its work on them (45-60 us per
// launch, measured); these three kernels stream the one large operand exactly once instead.
constexpr int SK_MAX = 8;

// C[m, n] = sum_k A[m, k] * B[n, k]   (N <= 8): one warp per row, B staged in shared memory
__global__ void __launch_bounds__(256) skinny_n_kernel(const float* __restrict__ A, const float* __restrict__ B, int M, int N,
                                                       int K, int lda, int ldb, Epilogue ep) {
  pdl_trigger();
  extern __shared__ float sB[];   // [N][K]
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) sB[i] = __ldg(B + (size_t)(i / K) * ldb + (i % K));
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += warps) {
    float acc[SK_MAX];
#pragma unroll
    for (int n = 0; n < SK_MAX; ++n) acc[n] = 0.f;
    const float* a = A + (size_t)m * lda;
    for (int k = lane; k < K; k += 32) {
      const float av = __ldg(a + k);
#pragma unroll
      for (int n = 0; n < SK_MAX; ++n)
        if (n < N) acc[n] = fmaf(av, sB[n * K + k], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < SK_MAX; ++n)
      if (n < N) acc[n] = warp_sum(acc[n]);
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < SK_MAX; ++n)
        if (n < N) ep.store(acc[n], m, n);
    }
  }
}

// C[m, n] = sum_{k < K} A[m, k] * B[k, n]   (K <= 8, B row-major [K, N]): one thread per output element
__global__ void __launch_bounds__(256) skinny_k_kernel(const float* __restrict__ A, const float* __restrict__ B, int M, int N,
                                                       int K, int lda, int ldb, Epilogue ep) {
  pdl_trigger();
  const size_t total = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / N), n = (int)(i - (size_t)m * N);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < SK_MAX; ++k)
      if (k < K) acc = fmaf(__ldg(A + (size_t)m * lda + k), __ldg(B + (size_t)k * ldb + n), acc);
    ep.store(acc, m, n);
  }
}

// C[i, n] = sum_k A[k, i] * B[k, n]   (M <= 8; A is [K, M], B is [K, N]): column sums over K in two deterministic stages.
// Stage 1: CTA (chunk, column block) accumulates its K rows; stage 2: chunks are folded in order and the epilogue runs.
__global__ void __launch_bounds__(256) skinny_m_stage1_kernel(const float* __restrict__ A, const float* __restrict__ B, int M,
                                                              int N, int K, int lda, int ldb, int rows_per_chunk,
                                                              float* __restrict__ partial) {
  pdl_trigger();
  __shared__ float sA[64][SK_MAX];
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int k0 = blockIdx.y * rows_per_chunk, k1 = min(K, k0 + rows_per_chunk);
  float acc[SK_MAX];
#pragma unroll
  for (int i = 0; i < SK_MAX; ++i) acc[i] = 0.f;
  for (int kb = k0; kb < k1; kb += 64) {
    const int nk = min(64, k1 - kb);
    __syncthreads();
    for (int t = threadIdx.x; t < nk * M; t += blockDim.x) sA[t / M][t % M] = __ldg(A + (size_t)(kb + t / M) * lda + (t % M));
    __syncthreads();
    if (n < N) {
      for (int kk = 0; kk < nk; ++kk) {
        const float b = __ldg(B + (size_t)(kb + kk) * ldb + n);
#pragma unroll
        for (int i = 0; i < SK_MAX; ++i)
          if (i < M) acc[i] = fmaf(sA[kk][i], b, acc[i]);
      }
    }
  }
  if (n < N) {
#pragma unroll
    for (int i = 0; i < SK_MAX; ++i)
      if (i < M) partial[((size_t)blockIdx.y * M + i) * N + n] = acc[i];
  }
}

static bool skinny_gemm(const gb_gemm_args* a, const Epilogue& ep, cudaStream_t stream, int* rc) {
  const int M = a->M, N = a->N, K = a->K;
  *rc = GB_OK;
  if (K == 0) return false;
  if (!a->trans_a && !a->trans_b && N <= SK_MAX && (size_t)N * K * 4 <= 32768 && M >= 64) {
    int blocks = (M + 7) / 8;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    skinny_n_kernel<<<blocks, 256, (size_t)N * K * 4, stream>>>(a->A, a->B, M, N, K, a->lda, a->ldb, ep);
    count_launch();
    return true;
  }
  if (!a->trans_a && a->trans_b && K <= SK_MAX && M >= 64) {
    size_t total = (size_t)M * N;
    int blocks = (int)((total + 255) / 256);
    if (blocks > sm_count() * 16) blocks = sm_count() * 16;
    skinny_k_kernel<<<blocks, 256, 0, stream>>>(a->A, a->B, M, N, K, a->lda, a->ldb, ep);
    count_launch();
    return true;
  }
  if (a->trans_a && a->trans_b && M <= SK_MAX && a->workspace && K >= 256) {
    const int col_blocks = (N + 255) / 256;
    // few, long chunks: the fold over chunks is a serial loop per output element (ncu: 296 chunks made the 2-CTA reduce
    // 35 us), while 48 CTAs already stream a 7 MB operand in a few microseconds
    int chunks = (2 * sm_count()) / col_blocks;
    if (chunks > 48) chunks = 48;
    if (chunks > (K + 63) / 64) chunks = (K + 63) / 64;
    const long long by_ws = a->workspace_bytes / ((long long)M * N * 4);
    if (chunks > by_ws) chunks = (int)by_ws;
    if (chunks < 1) return false;
    const int rows_per_chunk = ((K + chunks - 1) / chunks + 63) / 64 * 64;
    chunks = (K + rows_per_chunk - 1) / rows_per_chunk;
    skinny_m_stage1_kernel<<<dim3(col_blocks, chunks), 256, 0, stream>>>(a->A, a->B, M, N, K, a->lda, a->ldb, rows_per_chunk, a->workspace);
    count_launch();
    *rc = launch_splitk_reduce(a->workspace, chunks, M, N, ep, stream);
    return true;
  }
  return false;
}

int gemm_simt(const gb_gemm_args* a, cudaStream_t stream) {
  Epilogue ep = make_epilogue(a);
  const int M = a->M, N = a->N, K = a->K;
  {
    int rc;
    if (skinny_gemm(a, ep, stream, &rc)) {
      if (rc) return rc;
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) { set_error("skinny gemm launch -> %s", cudaGetErrorString(e)); return GB_ERR_CUDA; }
      return GB_OK;
    }
  }
  const int gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  // split-K when the output grid cannot fill the machine and K is long (weight gradients)
  int splits = 1;
  const int sms = sm_count();
  if (a->workspace && gx * gy < sms && K >= 1024) {
    splits = (2 * sms + gx * gy - 1) / (gx * gy);
    int max_by_k = K / 256;
    if (splits > max_by_k) splits = max_by_k;
    long long max_by_ws = a->workspace_bytes / ((long long)M * N * 4);
    if (splits > max_by_ws) splits = (int)max_by_ws;
    if (splits < 1) splits = 1;
  }
  int k_chunk = ((K + splits - 1) / splits + BK - 1) / BK * BK;
  if (k_chunk <= 0) k_chunk = BK;
  splits = (K + k_chunk - 1) / k_chunk;
  if (splits < 1) splits = 1;
  const bool a16 = ((uintptr_t)a->A % 16 == 0) && (a->lda % 4 == 0);
  const bool b16 = ((uintptr_t)a->B % 16 == 0) && (a->ldb % 4 == 0);
  float* partial = splits > 1 ? a->workspace : nullptr;
  dim3 grid(gx, gy, splits);
#define GB_LAUNCH(TA, TB) \
  sgemm_kernel<TA, TB><<<grid, 256, 0, stream>>>(a->A, a->B, M, N, K, a->lda, a->ldb, a16, b16, k_chunk, partial, ep)
  if (!a->trans_a && !a->trans_b) GB_LAUNCH(0, 0);
  else if (!a->trans_a && a->trans_b) GB_LAUNCH(0, 1);
  else if (a->trans_a && !a->trans_b) GB_LAUNCH(1, 0);
  else GB_LAUNCH(1, 1);
#undef GB_LAUNCH
  GB_CHECK_LAUNCH();
  if (splits > 1) return launch_splitk_reduce(partial, splits, M, N, ep, stream);
  return GB_OK;
}

int gemm_tcgen05(const gb_gemm_args* a, cudaStream_t stream, bool* handled);
int gemm_tcgen05_grouped(const gb_gemm_args* list, int n, cudaStream_t stream, bool* handled);
bool gemm_tcgen05_can_fuse_colsum(const gb_gemm_args* a);

}  // namespace gb

extern "C" int grappa_b200_gemm(const gb_gemm_args* a, void* stream_);

extern "C" int grappa_b200_gemm_can_fuse_colsum(const gb_gemm_args* a) {
  if (!a || a->M <= 0 || a->N <= 0 || !(a->precision >= 1 && a->precision <= 3)) return 0;
  return gb::gemm_tcgen05_can_fuse_colsum(a) ? 1 : 0;
}

extern "C" int grappa_b200_gemm_grouped(const gb_gemm_args* list, int32_t n, void* stream_) {
  GB_REQUIRE(list != nullptr || n == 0, "gemm_grouped: list is NULL");
  GB_REQUIRE(n >= 0, "gemm_grouped: negative count");
  cudaStream_t stream = (cudaStream_t)stream_;
  int i = 0;
  while (i < n) {
    // longest run (<= 4) of non-empty tensor-core problems that one persistent launch can serve
    int j = i;
    while (j < n && j - i < 4 && list[j].M > 0 && list[j].N > 0 && list[j].precision >= 1 && list[j].precision <= 3 &&
           (list[j].precision == 3) == (list[i].precision == 3)) ++j;
    if (j - i >= 2) {
      bool handled = false;
      int rc = gb::gemm_tcgen05_grouped(list + i, j - i, stream, &handled);
      if (rc != GB_OK) return rc;
      if (handled) { i = j; continue; }
    }
    int rc = grappa_b200_gemm(list + i, stream_);
    if (rc != GB_OK) return rc;
    ++i;
  }
  return GB_OK;
}

extern "C" int grappa_b200_gemm(const gb_gemm_args* a, void* stream_) {
  GB_REQUIRE(a != nullptr, "gemm: args is NULL");
  GB_REQUIRE(a->M >= 0 && a->N >= 0 && a->K >= 0, "gemm: negative dimension");
  if (a->M == 0 || a->N == 0) return GB_OK;
  GB_REQUIRE(a->A && a->B && a->C, "gemm: NULL operand");
  GB_REQUIRE(a->lda >= (a->trans_a ? a->M : a->K) && a->ldb >= (a->trans_b ? a->N : a->K) && a->ldc >= a->N,
             "gemm: leading dimension too small (M=%d N=%d K=%d lda=%d ldb=%d ldc=%d ta=%d tb=%d)", a->M, a->N, a->K,
             a->lda, a->ldb, a->ldc, a->trans_a, a->trans_b);
  GB_REQUIRE(a->dropout_p >= 0.f && a->dropout_p < 1.f, "gemm: dropout_p must be in [0,1)");
  cudaStream_t stream = (cudaStream_t)stream_;
  GB_REQUIRE(a->precision >= 0 && a->precision <= 3, "gemm: precision must be 0 (fp32), 1 (tf32), 2 (auto tf32) or 3 (bf16x3)");
  if (a->precision >= 1) {
    bool handled = false;
    int rc = gb::gemm_tcgen05(a, stream, &handled);
    if (rc != GB_OK) return rc;
    if (handled) return GB_OK;
    GB_REQUIRE(a->precision != 1, "gemm: tcgen05 path cannot take this shape/alignment (M=%d N=%d K=%d lda=%d ldb=%d)",
               a->M, a->N, a->K, a->lda, a->ldb);
  }
  GB_REQUIRE(a->colsum == nullptr, "gemm: fused column sums need the tensor-core path (ask grappa_b200_gemm_can_fuse_colsum)");
  return gb::gemm_simt(a, stream);
}
