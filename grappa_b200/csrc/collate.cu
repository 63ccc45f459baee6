// Device-side batch assembly: replaces the per-batch host work of the reference's
// data/GraphDataLoader.py:23-73 (collate_fn), utils/dgl_utils.py:132-171 (set_number_confs: conformation sub-sampling /
// padding) and utils/dgl_utils.py:11-60 (batch: per-type concatenation, `idxs += atom offset`) -- plus the index tables
// of pack.PackedBatch -- by ONE kernel over a dataset that lives in HBM.
//
// Everything a batch consists of is a concatenation of per-molecule pieces, optionally shifted by a per-molecule offset:
// node features / labels (plain rows), conformation fields (rows x selected conformations), tuple indices (+ atom offset),
// and the per-molecule index tables precomputed once per dataset (inverse incidence CSR, bonded-graph CSR, reverse-edge
// table, conflict-free schedule; + tuple / edge / atom offsets).  A job describes one such concatenation; the kernel
// maps every output word to (molecule of the batch, local element) with a binary search over the batch offsets.
#include "common.cuh"

namespace gb {

__device__ __forceinline__ int find_mol(const int32_t* __restrict__ off, int n, int row) {
  int lo = 0, hi = n;   // off[lo] <= row < off[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(off + mid) <= row) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) collate_kernel(const gb_collate_args* __restrict__ args) {
  const gb_collate_job& j = args->job[blockIdx.y];
  const int B = args->B;
  const int32_t* __restrict__ mol = args->mol;
  const int kind = j.kind;
  if (kind == 4) {   // plain copy of `n_rows * row_words` words (host-computed tables shipped in the upload buffer)
    const long long n = (long long)j.n_rows * j.row_words;
    const int32_t* s = reinterpret_cast<const int32_t*>(j.src);
    int32_t* d = reinterpret_cast<int32_t*>(j.dst);
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) d[e] = s[e];
    return;
  }
  const int words = kind == 3 ? j.row_words * j.n_confs_out : j.row_words;   // output words per row
  const long long total = (long long)j.n_rows * words;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(e / words);
    const int w = (int)(e - (long long)row * words);
    const int b = find_mol(j.dst_off, B, row);
    const int local = row - __ldg(j.dst_off + b);
    const int m = __ldg(mol + b);
    const long long base = j.src_off[m];
    if (kind == 3) {
      const int tail = j.row_words;                 // words per (row, conformation): 3 for xyz / gradients, 1 for energies
      const int c = w / tail, t = w - c * tail;
      const int sc = __ldg(j.csel + (size_t)b * j.n_confs_out + c);
      const int nc = __ldg(j.confs + m);
      reinterpret_cast<int32_t*>(j.dst)[e] =
          __ldg(reinterpret_cast<const int32_t*>(j.src) + base + ((long long)local * nc + sc) * tail + t);
    } else if (kind == 2) {                         // int64 elements + per-molecule offset
      const long long v = reinterpret_cast<const long long*>(j.src)[(base + local) * (words >> 1) + (w >> 1)];
      if ((w & 1) == 0) reinterpret_cast<long long*>(j.dst)[e >> 1] = v + (j.add ? (long long)__ldg(j.add + b) : 0ll);
    } else {
      int32_t v = __ldg(reinterpret_cast<const int32_t*>(j.src) + (base + local) * words + w);
      if (kind == 1 && j.add && v >= 0) v += __ldg(j.add + b);     // negative entries (idle schedule slots) stay as they are
      reinterpret_cast<int32_t*>(j.dst)[e] = v;
    }
  }
}

}  // namespace gb

extern "C" int grappa_b200_collate(const gb_collate_args* dev_args, int32_t n_jobs, int64_t max_words, void* stream_) {
  GB_REQUIRE(dev_args != nullptr, "collate: args is NULL");
  GB_REQUIRE(n_jobs >= 0 && n_jobs <= GB_COLLATE_MAX_JOBS, "collate: at most %d jobs", GB_COLLATE_MAX_JOBS);
  if (n_jobs == 0) return GB_OK;
  long long blocks = (max_words + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 2 * gb::sm_count()) blocks = 2 * gb::sm_count();
  gb::collate_kernel<<<dim3((unsigned)blocks, (unsigned)n_jobs), 256, 0, (cudaStream_t)stream_>>>(dev_args);
  GB_CHECK_LAUNCH();
  return GB_OK;
}
