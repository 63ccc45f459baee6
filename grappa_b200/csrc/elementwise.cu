// Small fused kernels around the GEMMs: input featurisation, writer output maps (+ permutation sum),
// dropout, axpby, squared-norm, Adam with global-norm clipping, molecule-wise loss.
#include "common.cuh"

namespace gb {

// ------------------------------------------------------------------------------------------------
// featurisation: concat + sinusoidal charge encoding (reference models/graph_attention.py:157-164,428-444)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) featurize_kernel(gb_featurize_args a, float* __restrict__ out, int n, int ld,
                                                        int width_total) {
  pdl_trigger();
  const long long total = (long long)n * ld;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(i / ld);
    int c = (int)(i - (long long)row * ld);
    float v = 0.f;
    if (c < width_total) {
      for (int f = 0; f < a.n_feats; ++f) {
        if (c < a.width[f]) { v = __ldg(a.feats[f] + (size_t)row * a.width[f] + c); break; }
        c -= a.width[f];
      }
    } else if (a.charge && c < width_total + a.enc_dim) {
      const int j = c - width_total;
      float q = fminf(fmaxf(__ldg(a.charge + row), -2.f), 2.f);
      const float s = (q + 2.f) / 4.f;
      const float freq = expf((float)(j >> 1) * -logf(10000.f) / (float)(a.enc_dim >> 1));
      v = (j & 1) ? cosf(s * freq) : sinf(s * freq);
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// writer output maps
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float to_positive(float x, float mos, float sd, float mn) {
  return sd * (elu1(mos + x - 1.f) + 1.f) + mn;
}
__device__ __forceinline__ float to_positive_grad(float x, float mos, float sd) {
  const float z = mos + x - 1.f;
  return sd * (z > 0.f ? 1.f : expf(z));
}

// The statistics either come from the argument struct (buffers: host copies baked into the launch) or, with
// learnable_statistics, from the parameters themselves in device memory (a->stat[]: a captured step then sees every
// optimizer update).  kind 0: {k m/s, k std, eq m/s, eq std}; kind 1: {k m/s, k std, eq std/max}; kind 2: {k_std[n], k_mean[n]}.
struct HeadStats {
  float k_mos, k_std, eq_mos, eq_std, eq_som;
  float tk_std[6], tk_mean[6];
};
__device__ __forceinline__ HeadStats head_stats(const gb_head_out_args& a) {
  HeadStats h;
  h.k_mos = a.k_mean_over_std; h.k_std = a.k_std;
  h.eq_mos = a.eq_mean_over_std; h.eq_std = a.eq_std; h.eq_som = a.eq_std_over_max;
#pragma unroll
  for (int n = 0; n < 6; ++n) { h.tk_std[n] = a.tk_std[n]; h.tk_mean[n] = a.tk_mean[n]; }
  if (a.kind <= 1) {
    if (a.stat[0]) h.k_mos = __ldg(a.stat[0]);
    if (a.stat[1]) h.k_std = __ldg(a.stat[1]);
    if (a.kind == 0) {
      if (a.stat[2]) h.eq_mos = __ldg(a.stat[2]);
      if (a.stat[3]) h.eq_std = __ldg(a.stat[3]);
    } else if (a.stat[2]) {
      h.eq_som = __ldg(a.stat[2]);
    }
  } else {
#pragma unroll
    for (int n = 0; n < 6; ++n) {
      if (n < a.n_per && a.stat[0]) h.tk_std[n] = __ldg(a.stat[0] + n);
      if (n < a.n_per && a.stat[1]) h.tk_mean[n] = __ldg(a.stat[1] + n);
    }
  }
  return h;
}

__global__ void __launch_bounds__(256) head_output_fwd_kernel(gb_head_out_args a, const float* __restrict__ scores,
                                                              float* __restrict__ k, float* __restrict__ eq) {
  pdl_trigger();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.T) return;
  const HeadStats h = head_stats(a);
  float c[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    c[j] = 0.f;
    if (j < a.n_out)
      for (int p = 0; p < a.n_perm; ++p) c[j] += __ldg(scores + ((size_t)p * a.T + t) * a.n_out + j);
  }
  if (a.kind == 0) {
    eq[t] = to_positive(c[0], h.eq_mos, h.eq_std, a.eq_min);
    k[t] = to_positive(c[1], h.k_mos, h.k_std, a.k_min);
  } else if (a.kind == 1) {
    eq[t] = a.eq_max * sigmoidf_(h.eq_som * c[0]);
    k[t] = to_positive(c[1], h.k_mos, h.k_std, a.k_min);
  } else {
#pragma unroll
    for (int n = 0; n < 6; ++n) {
      if (n < a.n_per) {
        float v = a.gated ? c[n] * sigmoidf_(c[n + a.n_per]) * h.tk_std[n] : c[n] * h.tk_std[n] + h.tk_mean[n];
        if (a.cutoff > 0.f && !(fabsf(v) > a.cutoff)) v = 0.f;
        k[(size_t)t * a.n_per + n] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) head_output_bwd_kernel(gb_head_out_args a, const float* __restrict__ scores,
                                                              const float* __restrict__ dk, const float* __restrict__ deq,
                                                              float* __restrict__ dscores) {
  pdl_trigger();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.T) return;
  const HeadStats h = head_stats(a);
  float c[12], d[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    c[j] = 0.f;
    d[j] = 0.f;
    if (j < a.n_out)
      for (int p = 0; p < a.n_perm; ++p) c[j] += __ldg(scores + ((size_t)p * a.T + t) * a.n_out + j);
  }
  if (a.kind == 0) {
    if (deq) d[0] = __ldg(deq + t) * to_positive_grad(c[0], h.eq_mos, h.eq_std);
    if (dk) d[1] = __ldg(dk + t) * to_positive_grad(c[1], h.k_mos, h.k_std);
  } else if (a.kind == 1) {
    if (deq) {
      const float s = sigmoidf_(h.eq_som * c[0]);
      d[0] = __ldg(deq + t) * a.eq_max * h.eq_som * s * (1.f - s);
    }
    if (dk) d[1] = __ldg(dk + t) * to_positive_grad(c[1], h.k_mos, h.k_std);
  } else if (dk) {
#pragma unroll
    for (int n = 0; n < 6; ++n) {
      if (n < a.n_per) {
        const float g = __ldg(dk + (size_t)t * a.n_per + n);
        if (a.gated) {
          const float s = sigmoidf_(c[n + a.n_per]);
          const float v = c[n] * s * h.tk_std[n];
          const bool keep = !(a.cutoff > 0.f) || fabsf(v) > a.cutoff;
          d[n] = keep ? g * s * h.tk_std[n] : 0.f;
          d[n + a.n_per] = keep ? g * c[n] * s * (1.f - s) * h.tk_std[n] : 0.f;
        } else {
          const float v = c[n] * h.tk_std[n] + h.tk_mean[n];
          const bool keep = !(a.cutoff > 0.f) || fabsf(v) > a.cutoff;
          d[n] = keep ? g * h.tk_std[n] : 0.f;
        }
      }
    }
  }
  for (int p = 0; p < a.n_perm; ++p)
#pragma unroll
    for (int j = 0; j < 12; ++j)
      if (j < a.n_out) dscores[((size_t)p * a.T + t) * a.n_out + j] = d[j];
}

// Gradients of the learnable statistics: one CTA walks all tuples and reduces up to 12 sums in a fixed order (no atomics:
// the step stays bit-reproducible).  The slots follow a->stat[]: kinds 0/1 -> one value per pointer; kind 2 -> n_per values
// for k_std (slots 0..5) and n_per for k_mean (slots 6..11; zero when gated, where the mean does not enter).
__global__ void __launch_bounds__(1024) head_output_stats_bwd_kernel(gb_head_out_args a, const float* __restrict__ scores,
                                                                     const float* __restrict__ dk,
                                                                     const float* __restrict__ deq, gb_head_stat_grads o) {
  pdl_trigger();
  const HeadStats h = head_stats(a);
  float acc[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) acc[j] = 0.f;
  for (int t = threadIdx.x; t < a.T; t += blockDim.x) {
    float c[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) {
      c[j] = 0.f;
      if (j < a.n_out)
        for (int p = 0; p < a.n_perm; ++p) c[j] += __ldg(scores + ((size_t)p * a.T + t) * a.n_out + j);
    }
    if (a.kind <= 1) {
      if (dk) {
        const float g = __ldg(dk + t), z = h.k_mos + c[1] - 1.f;
        acc[0] += g * h.k_std * (z > 0.f ? 1.f : expf(z));      // d/d(mean_over_std)
        acc[1] += g * (elu1(z) + 1.f);                           // d/d(std)
      }
      if (deq) {
        const float g = __ldg(deq + t);
        if (a.kind == 0) {
          const float z = h.eq_mos + c[0] - 1.f;
          acc[2] += g * h.eq_std * (z > 0.f ? 1.f : expf(z));
          acc[3] += g * (elu1(z) + 1.f);
        } else {
          const float s = sigmoidf_(h.eq_som * c[0]);
          acc[2] += g * a.eq_max * s * (1.f - s) * c[0];         // d/d(std_over_max)
        }
      }
    } else if (dk) {
#pragma unroll
      for (int n = 0; n < 6; ++n) {
        if (n < a.n_per) {
          const float g = __ldg(dk + (size_t)t * a.n_per + n);
          if (a.gated) {
            const float u = c[n] * sigmoidf_(c[n + a.n_per]);
            const bool keep = !(a.cutoff > 0.f) || fabsf(u * h.tk_std[n]) > a.cutoff;
            if (keep) acc[n] += g * u;
          } else {
            const bool keep = !(a.cutoff > 0.f) || fabsf(c[n] * h.tk_std[n] + h.tk_mean[n]) > a.cutoff;
            if (keep) { acc[n] += g * c[n]; acc[6 + n] += g; }
          }
        }
      }
    }
  }
  __shared__ float red[32][12];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    float v = acc[j];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) red[w][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    float v = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += red[i][threadIdx.x];
    const int j = threadIdx.x;
    if (a.kind <= 1) {
      if (j < 4 && o.d[j]) o.d[j][0] = (o.accumulate ? o.d[j][0] : 0.f) + v;
    } else {
      float* dst = o.d[j / 6];
      if (dst && (j % 6) < a.n_per) dst[j % 6] = (o.accumulate ? dst[j % 6] : 0.f) + v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dropout / axpby / sumsq / adam
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dropout_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                      uint32_t thresh, float inv_keep, uint64_t seed0,
                                                      const uint64_t* __restrict__ seed_off) {
  pdl_trigger();
  const uint64_t seed = seed_with_offset(seed0, seed_off);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = x[i] * dropout_scale(seed, (uint64_t)i, thresh, inv_keep);
}

__global__ void __launch_bounds__(256) act_dropout_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ act_out,
                                                              float* __restrict__ dx, long long n, uint32_t thresh,
                                                              float inv_keep, uint64_t seed0,
                                                              const uint64_t* __restrict__ seed_off) {
  pdl_trigger();
  const uint64_t seed = seed_with_offset(seed0, seed_off);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = dy[i];
    if (thresh) v *= dropout_scale(seed, (uint64_t)i, thresh, inv_keep);
    if (act_out) v *= elu1_grad_from_out(act_out[i]);
    dx[i] = v;
  }
}

// out[r, c] = c < cols ? in[r, c] : 0   (row pitch ld_in -> ld_out): a weight whose rows are not 16-byte multiples
// (pre_dense: 85 input features) gets a TMA-legal padded copy so its GEMM can run on the tensor cores
__global__ void __launch_bounds__(256) pad_rows_kernel(const float* __restrict__ in, int rows, int cols, int ld_in,
                                                       float* __restrict__ out, int ld_out) {
  pdl_trigger();
  const long long total = (long long)rows * ld_out;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_out), c = (int)(i - (long long)r * ld_out);
    out[i] = c < cols ? __ldg(in + (size_t)r * ld_in + c) : 0.f;
  }
}

__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float a,
                                                    float b) {
  pdl_trigger();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = b == 0.f ? a * x[i] : a * x[i] + b * y[i];
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* out) {
  pdl_trigger();
  __shared__ float sh[8];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(out, t);
  }
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                                   float bc1, float bc2_sqrt, const float* __restrict__ gnorm_sq, float clip,
                                                   float grad_scale) {
  pdl_trigger();
  float coef = grad_scale;
  if (gnorm_sq) {
    const float ss = __ldg(gnorm_sq);
    // a non-finite gradient norm (one NaN / inf anywhere in the flat gradient) would turn EVERY parameter and both
    // moment buffers into NaN in this one step: the step is skipped instead, parameters and moments stay as they are
    if (!(ss >= 0.f && ss <= 3.0e38f)) return;
    if (clip > 0.f) {
      // the norm was accumulated on UNSCALED gradients: total norm of the scaled gradient = scale * sqrt(sumsq)
      const float total = grad_scale * sqrtf(ss);
      coef *= fminf(1.f, clip / (total + 1e-6f));
    }
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= (lr / bc1) * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// ------------------------------------------------------------------------------------------------
// molecule-wise loss: one CTA per molecule
// ------------------------------------------------------------------------------------------------
__device__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(256) molwise_loss_kernel(gb_loss_args a) {
  pdl_trigger();
  __shared__ float sh[8];
  const int b = blockIdx.x, tid = threadIdx.x, C = a.C;
  const float invB = (a.grad_scale ? __ldg(a.grad_scale) : 1.f) / a.B;
  const float invB_loss = 1.f / a.B;
  float term = 0.f;
  const int nv = a.n_valid ? min(max(a.n_valid[b], 0), C) : C;   // real (non-padding) conformations of this molecule
  // energies
  if (a.w_energy != 0.f && a.energy && nv > 0) {
    float s1 = 0.f, s2 = 0.f;
    for (int c = tid; c < nv; c += blockDim.x) {
      s1 += a.energy[(size_t)b * C + c];
      s2 += a.energy_ref[(size_t)b * C + c];
    }
    const float me = block_sum(s1, sh) / nv, mr = block_sum(s2, sh) / nv;
    float q = 0.f;
    for (int c = tid; c < C; c += blockDim.x) {
      float d = 0.f;
      if (c < nv) {
        d = (a.energy[(size_t)b * C + c] - me) - (a.energy_ref[(size_t)b * C + c] - mr);
        q += d * d;
      }
      if (a.g_energy) a.g_energy[(size_t)b * C + c] = a.w_energy * invB * 2.f * d / nv;
    }
    term += a.w_energy * block_sum(q, sh) / nv;
  } else if (a.g_energy) {
    for (int c = tid; c < C; c += blockDim.x) a.g_energy[(size_t)b * C + c] = 0.f;
  }
  // gradients
  const int a0 = a.atom_off[b], a1 = a.atom_off[b + 1];
  const long long n0 = (long long)a0 * C * 3, n1 = (long long)a1 * C * 3;
  if (a.w_grad != 0.f && a.grad && nv > 0 && a1 > a0) {
    float q = 0.f;
    const float n_el = (float)(a1 - a0) * (float)nv * 3.f;
    const float sc = a.w_grad * invB * 2.f / n_el;
    const int row = C * 3, valid = nv * 3;               // per atom: C conformations x 3, the first nv x 3 are real
    for (long long i = n0 + tid; i < n1; i += blockDim.x) {
      float d = 0.f;
      if (nv == C || (int)((i - n0) % row) < valid) {
        d = a.grad[i] - a.grad_ref[i];
        q += d * d;
      }
      if (a.g_grad) a.g_grad[i] = sc * d;
    }
    term += a.w_grad * block_sum(q, sh) / n_el;
  } else if (a.g_grad) {
    for (long long i = n0 + tid; i < n1; i += blockDim.x) a.g_grad[i] = 0.f;
  }
  // torsion L2 regularisers
  for (int which = 0; which < 2; ++which) {
    const float* k = which ? a.k_improper : a.k_proper;
    float* gk = which ? a.g_k_improper : a.g_k_proper;
    const int32_t* off = which ? a.improper_off : a.proper_off;
    const int nper = which ? a.n_per_i : a.n_per_p;
    const float w = which ? a.w_improper : a.w_proper;
    if (!k || !off) continue;
    const long long i0 = (long long)off[b] * nper, i1 = (long long)off[b + 1] * nper;
    if (i1 <= i0) continue;
    float q = 0.f;
    const float sc = w * invB * 2.f / (float)(i1 - i0);
    for (long long i = i0 + tid; i < i1; i += blockDim.x) {
      const float v = k[i];
      q += v * v;
      if (gk) gk[i] = sc * v;
    }
    term += w * block_sum(q, sh) / (float)(i1 - i0);
  }
  if (tid == 0) a.mol_loss[b] = (term + (a.extra_mol_loss ? a.extra_mol_loss[b] : 0.f)) * invB_loss;
}

// classical-parameter MSE, one block per molecule (see gb_param_loss_args)
__global__ void __launch_bounds__(256) param_loss_kernel(gb_param_loss_args a) {
  pdl_trigger();
  __shared__ float sh[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  long long total = 0;
  for (int i = 0; i < a.n_terms; ++i) total += (long long)(a.off[i][b + 1] - a.off[i][b]) * a.width[i];
  const float w = a.mol_weight[b];
  const float inv_total = total > 0 ? 1.f / (float)total : 0.f;
  const float gsc = (a.grad_scale ? __ldg(a.grad_scale) : 1.f) / a.B * w * 2.f * inv_total;
  float q = 0.f;
  for (int i = 0; i < a.n_terms; ++i) {
    const int wd = a.width[i], rw = a.ref_width[i];
    const long long e0 = (long long)a.off[i][b] * wd, e1 = (long long)a.off[i][b + 1] * wd;
    const float fac = a.fac[i];
    for (long long e = e0 + tid; e < e1; e += blockDim.x) {
      const long long t = e / wd;
      const int j = (int)(e - t * wd);
      const float r = j < rw ? a.ref[i][t * rw + j] : 0.f;
      const float d = (r != r) ? 0.f : fac * (a.pred[i][e] - r);
      q += d * d;
      if (a.g_pred[i]) a.g_pred[i][e] = gsc * fac * d;
    }
  }
  q = block_sum(q, sh);
  if (tid == 0) a.mol_loss[b] = w * q * inv_total;
}

__global__ void __launch_bounds__(256) loss_final_kernel(const float* __restrict__ mol_loss, int B, float* loss) {
  pdl_trigger();
  __shared__ float sh[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) s += mol_loss[i];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) loss[0] = s;
}

static int grid_for(long long n) {
  long long b = (n + 255) / 256;
  long long cap = (long long)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace gb

using namespace gb;

extern "C" int grappa_b200_featurize(const gb_featurize_args* a, float* out, int32_t n, int32_t ld, void* stream_) {
  GB_REQUIRE(a && out, "featurize: NULL pointer");
  GB_REQUIRE(a->n_feats >= 0 && a->n_feats <= 8, "featurize: at most 8 feature tensors");
  int w = 0;
  for (int f = 0; f < a->n_feats; ++f) {
    GB_REQUIRE(a->feats[f] != nullptr && a->width[f] > 0, "featurize: feature %d is NULL / empty", f);
    w += a->width[f];
  }
  const int enc = a->charge ? a->enc_dim : 0;
  GB_REQUIRE(enc % 2 == 0, "featurize: encoding dimension must be even");
  GB_REQUIRE(ld >= w + enc, "featurize: ld %d < %d features", ld, w + enc);
  if (n == 0) return GB_OK;
  featurize_kernel<<<grid_for((long long)n * ld), 256, 0, (cudaStream_t)stream_>>>(*a, out, n, ld, w);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

static int check_head(const gb_head_out_args* a) {
  GB_REQUIRE(a != nullptr, "head_output: args is NULL");
  GB_REQUIRE(a->kind >= 0 && a->kind <= 2, "head_output: kind must be 0, 1 or 2");
  GB_REQUIRE(a->n_perm >= 1 && a->n_perm <= 6 && a->n_out >= 1 && a->n_out <= 12, "head_output: bad n_perm / n_out");
  if (a->kind == 2) {
    GB_REQUIRE(a->n_per >= 1 && a->n_per <= 6, "head_output: n_periodicity must be 1..6");
    GB_REQUIRE(a->n_out == (a->gated ? 2 : 1) * a->n_per, "head_output: n_out does not match n_periodicity / gating");
  } else {
    GB_REQUIRE(a->n_out >= 2, "head_output: bonds / angles need >= 2 outputs");
  }
  return GB_OK;
}

extern "C" int grappa_b200_head_output_fwd(const gb_head_out_args* a, const float* scores, float* k, float* eq,
                                           void* stream_) {
  int rc = check_head(a);
  if (rc) return rc;
  if (a->T == 0) return GB_OK;
  GB_REQUIRE(scores && k && (a->kind == 2 || eq), "head_output_fwd: NULL pointer");
  head_output_fwd_kernel<<<(a->T + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(*a, scores, k, eq);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_head_output_stats_bwd(const gb_head_out_args* a, const float* scores, const float* dk,
                                                 const float* deq, const gb_head_stat_grads* out, void* stream_) {
  int rc = check_head(a);
  if (rc) return rc;
  GB_REQUIRE(out != nullptr && (a->T <= 0 || scores != nullptr), "head_output_stats_bwd: NULL pointer");
  if (a->T <= 0 && !out->accumulate) {   // no tuples: the gradients are zero
    for (int j = 0; j < 4; ++j)
      if (out->d[j]) GB_CHECK_CUDA(cudaMemsetAsync(out->d[j], 0, sizeof(float) * (a->kind == 2 ? a->n_per : 1), (cudaStream_t)stream_));
    return GB_OK;
  }
  if (a->T <= 0) return GB_OK;
  head_output_stats_bwd_kernel<<<1, 1024, 0, (cudaStream_t)stream_>>>(*a, scores, dk, deq, *out);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_head_output_bwd(const gb_head_out_args* a, const float* scores, const float* dk,
                                           const float* deq, float* dscores, void* stream_) {
  int rc = check_head(a);
  if (rc) return rc;
  if (a->T == 0) return GB_OK;
  GB_REQUIRE(scores && dscores, "head_output_bwd: NULL pointer");
  head_output_bwd_kernel<<<(a->T + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(*a, scores, dk, deq, dscores);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_dropout(const float* x, float* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_offset,
                                   void* stream_) {
  GB_REQUIRE(p >= 0.f && p < 1.f, "dropout: p must be in [0,1)");
  if (n == 0) return GB_OK;
  GB_REQUIRE(x && y, "dropout: NULL pointer");
  double t = (double)p * 4294967296.0;
  uint32_t thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
  if (p > 0.f && thresh == 0) thresh = 1;
  if (p == 0.f) thresh = 0;
  dropout_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(x, y, n, thresh, 1.f / (1.f - p), seed, seed_offset);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_act_dropout_bwd(const float* dy, const float* act_out, float* dx, int64_t n, float p,
                                           uint64_t seed, const uint64_t* seed_offset, void* stream_) {
  GB_REQUIRE(p >= 0.f && p < 1.f, "act_dropout_bwd: p must be in [0,1)");
  if (n == 0) return GB_OK;
  GB_REQUIRE(dy && dx, "act_dropout_bwd: NULL pointer");
  uint32_t thresh = 0;
  if (p > 0.f) {
    double t = (double)p * 4294967296.0;
    thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    if (thresh == 0) thresh = 1;
  }
  act_dropout_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(dy, act_out, dx, n, thresh, 1.f / (1.f - p), seed, seed_offset);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_pad_rows(const float* in, int32_t rows, int32_t cols, int32_t ld_in, float* out, int32_t ld_out,
                                    void* stream_) {
  if (rows == 0) return GB_OK;
  GB_REQUIRE(in && out, "pad_rows: NULL pointer");
  GB_REQUIRE(cols >= 0 && ld_in >= cols && ld_out >= cols, "pad_rows: pitch smaller than the row length");
  pad_rows_kernel<<<grid_for((long long)rows * ld_out), 256, 0, (cudaStream_t)stream_>>>(in, rows, cols, ld_in, out, ld_out);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_axpby(const float* x, float* y, int64_t n, float a, float b, void* stream_) {
  if (n == 0) return GB_OK;
  GB_REQUIRE(x && y, "axpby: NULL pointer");
  axpby_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(x, y, n, a, b);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

// deterministic variant: fixed grid, per-block partials, the last block (ticket) adds them in block order
__global__ void __launch_bounds__(256) sumsq_det_kernel(const float* __restrict__ x, long long n, float* __restrict__ out,
                                                        unsigned int* __restrict__ ticket, float* __restrict__ partial) {
  pdl_trigger();
  __shared__ float sh[8];
  __shared__ bool last;
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    s += x[i] * x[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    __stcg(partial + blockIdx.x, t);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float t = 0.f;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += 256) t += __ldcg(partial + i);   // fixed assignment -> fixed order
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int i = 0; i < 8; ++i) r += sh[i];
    *out = r;
    *ticket = 0u;
  }
}

extern "C" int grappa_b200_sumsq_det(const float* x, int64_t n, float* out, float* workspace, void* stream_) {
  GB_REQUIRE(x && out && workspace, "sumsq_det: NULL pointer");
  GB_REQUIRE(((uintptr_t)x & 15) == 0, "sumsq_det: x must be 16-byte aligned");
  int blocks = sm_count() * 4;
  if (blocks > 1000) blocks = 1000;
  long long need = (n / 4 + 255) / 256;
  if (need < 1) need = 1;
  if (blocks > need) blocks = (int)need;
  sumsq_det_kernel<<<blocks, 256, 0, (cudaStream_t)stream_>>>(x, n, out, reinterpret_cast<unsigned int*>(workspace), workspace + 4);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_sumsq(const float* x, int64_t n, float* out, void* stream_) {
  if (n == 0) return GB_OK;
  GB_REQUIRE(x && out, "sumsq: NULL pointer");
  sumsq_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(x, n, out);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                                     float beta2, float eps, int32_t step, const float* gnorm_sq, float clip,
                                     float grad_scale, void* stream_) {
  if (n == 0) return GB_OK;
  GB_REQUIRE(p && g && m && v, "adam_step: NULL pointer");
  GB_REQUIRE(step >= 1, "adam_step: step counts from 1");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream_>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, bc2s, gnorm_sq, clip, grad_scale);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                       float* __restrict__ v, long long n, const float* __restrict__ lr_dev,
                                                       float b1, float b2, float eps, const uint64_t* __restrict__ step_dev,
                                                       const float* __restrict__ gnorm_sq, float clip, float grad_scale) {
  pdl_trigger();
  const float step = (float)(__ldg(reinterpret_cast<const unsigned long long*>(step_dev)) + 1ull);
  const float lr = __ldg(lr_dev);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  float coef = grad_scale;
  if (gnorm_sq) {
    const float ss = __ldg(gnorm_sq);
    if (!(ss >= 0.f && ss <= 3.0e38f)) return;   // non-finite gradient: skip the step (see adam_kernel)
    if (clip > 0.f) {
      const float total = grad_scale * sqrtf(ss);
      coef *= fminf(1.f, clip / (total + 1e-6f));
    }
  }
  const long long n4 = n >> 2;   // flat buffers are 16-byte aligned (cudaMalloc / torch allocator)
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float a = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float* pf = &pp.x; float* gf = &gg.x; float* mf = &mm.x; float* vf = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = gf[j] * coef;
      mf[j] = b1 * mf[j] + (1.f - b1) * gi;
      vf[j] = b2 * vf[j] + (1.f - b2) * gi * gi;
      pf[j] -= a * mf[j] / (sqrtf(vf[j]) / bc2_sqrt + eps);
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= a * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

__global__ void tick_kernel(uint64_t* counters, int n) {
  pdl_trigger();
  if ((int)threadIdx.x < n) counters[threadIdx.x] += 1ull;
}

extern "C" int grappa_b200_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_dev,
                                         float beta1, float beta2, float eps, const uint64_t* step_dev,
                                         const float* gnorm_sq, float clip, float grad_scale, void* stream_) {
  if (n == 0) return GB_OK;
  GB_REQUIRE(p && g && m && v && lr_dev && step_dev, "adam_step_dev: NULL pointer");
  GB_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adam_step_dev: buffers must be 16-byte aligned");
  adam_dev_kernel<<<grid_for(n / 4 + 1), 256, 0, (cudaStream_t)stream_>>>(p, g, m, v, n, lr_dev, beta1, beta2, eps, step_dev, gnorm_sq, clip, grad_scale);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_tick(uint64_t* counters, int32_t n, void* stream_) {
  GB_REQUIRE(counters && n >= 1 && n <= 32, "tick: need 1..32 counters");
  tick_kernel<<<1, 32, 0, (cudaStream_t)stream_>>>(counters, n);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_param_loss(const gb_param_loss_args* a, void* stream_) {
  GB_REQUIRE(a != nullptr, "param_loss: args is NULL");
  GB_REQUIRE(a->B > 0, "param_loss: empty batch");
  GB_REQUIRE(a->n_terms >= 0 && a->n_terms <= GB_PARAM_LOSS_MAX_TERMS, "param_loss: at most %d terms", GB_PARAM_LOSS_MAX_TERMS);
  GB_REQUIRE(a->mol_weight && a->mol_loss, "param_loss: NULL pointer");
  for (int i = 0; i < a->n_terms; ++i) {
    GB_REQUIRE(a->pred[i] && a->ref[i] && a->off[i], "param_loss: term %d has a NULL pointer", i);
    GB_REQUIRE(a->width[i] > 0 && a->ref_width[i] >= 0, "param_loss: term %d has a bad width", i);
  }
  param_loss_kernel<<<a->B, 256, 0, (cudaStream_t)stream_>>>(*a);
  GB_CHECK_LAUNCH();
  return GB_OK;
}

extern "C" int grappa_b200_molwise_loss(const gb_loss_args* a, void* stream_) {
  GB_REQUIRE(a != nullptr, "molwise_loss: args is NULL");
  GB_REQUIRE(a->B > 0 && a->C > 0, "molwise_loss: empty batch");
  GB_REQUIRE(a->atom_off && a->loss && a->mol_loss, "molwise_loss: NULL pointer");
  GB_REQUIRE(!a->energy || a->energy_ref, "molwise_loss: energy without energy_ref");
  GB_REQUIRE(!a->grad || a->grad_ref, "molwise_loss: gradient without gradient_ref");
  cudaStream_t stream = (cudaStream_t)stream_;
  molwise_loss_kernel<<<a->B, 256, 0, stream>>>(*a);
  GB_CHECK_LAUNCH();
  loss_final_kernel<<<1, 256, 0, stream>>>(a->mol_loss, a->B, a->loss);
  GB_CHECK_LAUNCH();
  return GB_OK;
}
