// Fused GEMM epilogue shared by the FFMA and the tcgen05 kernels:
//   v = acc + bias[n]; v = ELU(v); v *= elu'(saved output); v *= dropout; v += residual; C (+)= v
#pragma once
#include "common.cuh"

namespace gb {

struct Epilogue {
  const float* bias;
  const float* mul_elu_out;
  const float* residual;
  float* C;
  float* act_out;
  float* colsum;     // [ceil(M/32), ldcs] column sums of the stored values per 32-row group, or NULL
  int ldcs;
  int ldc, ldm, ldr, ldact, N;
  int act, accumulate;
  uint32_t drop_thresh;   // 0 = dropout off
  float drop_inv_keep;
  uint64_t drop_seed;
  const uint64_t* drop_offset;   // device counter mixed into the seed (CUDA-graph replays), or NULL

  __device__ __forceinline__ uint64_t seed() const { return seed_with_offset(drop_seed, drop_offset); }
  // accumulate == 2: the old C joins the accumulator BEFORE bias / activation (two Linear layers under one activation)
  __device__ __forceinline__ float apply(float acc, int m, int n) const {
    float v = acc;
    if (accumulate == 2) v += C[(size_t)m * ldc + n];
    if (bias) v += __ldg(bias + n);
    if (act == 1) v = elu1(v);
    if (act_out) act_out[(size_t)m * ldact + n] = v;
    if (mul_elu_out) v *= elu1_grad_from_out(__ldg(mul_elu_out + (size_t)m * ldm + n));
    if (drop_thresh) v *= dropout_scale(seed(), (uint64_t)m * (uint64_t)N + (uint64_t)n, drop_thresh, drop_inv_keep);
    if (residual) v += __ldg(residual + (size_t)m * ldr + n);
    return v;
  }
  // 4 consecutive columns of one row (n % 4 == 0); every pointer / pitch is 16-byte aligned (vec_ok)
  __device__ __forceinline__ bool vec_ok() const {
    const uintptr_t ptrs = (uintptr_t)bias | (uintptr_t)mul_elu_out | (uintptr_t)residual | (uintptr_t)C | (uintptr_t)act_out;
    const int lds = ldc | (mul_elu_out ? ldm : 0) | (residual ? ldr : 0) | (act_out ? ldact : 0) | N;
    return (ptrs & 15) == 0 && (lds & 3) == 0;
  }
  // FAST: bias already added by the caller, ELU through ex2.approx (tensor-core path; |abs err| ~1e-7)
  template <bool FAST>
  __device__ __forceinline__ void store4(float4 acc, int m, int n) const {
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if (accumulate == 2) {
      const float4 o = *reinterpret_cast<const float4*>(C + (size_t)m * ldc + n);
      v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    if (!FAST && bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n));
      v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = FAST ? (v[i] > 0.f ? v[i] : __expf(v[i]) - 1.f) : elu1(v[i]);
    }
    if (act_out) *reinterpret_cast<float4*>(act_out + (size_t)m * ldact + n) = make_float4(v[0], v[1], v[2], v[3]);
    if (mul_elu_out) {
      const float4 y = __ldg(reinterpret_cast<const float4*>(mul_elu_out + (size_t)m * ldm + n));
      v[0] *= elu1_grad_from_out(y.x); v[1] *= elu1_grad_from_out(y.y);
      v[2] *= elu1_grad_from_out(y.z); v[3] *= elu1_grad_from_out(y.w);
    }
    if (drop_thresh) {
      const uint64_t base = (uint64_t)m * (uint64_t)N + (uint64_t)n;
      if ((N & 3) == 0) {   // aligned group of four: one key, two finalisers
        const float4 d = dropout_scale4(seed(), base, drop_thresh, drop_inv_keep);
        v[0] *= d.x; v[1] *= d.y; v[2] *= d.z; v[3] *= d.w;
      } else {
        const uint64_t sd = seed();
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] *= dropout_scale(sd, base + i, drop_thresh, drop_inv_keep);
      }
    }
    if (residual) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(residual + (size_t)m * ldr + n));
      v[0] += r.x; v[1] += r.y; v[2] += r.z; v[3] += r.w;
    }
    float4* c = reinterpret_cast<float4*>(C + (size_t)m * ldc + n);
    if (accumulate == 1) {
      const float4 o = *c;
      v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    *c = make_float4(v[0], v[1], v[2], v[3]);
  }
  // tensor-core path: bias already added, residual / saved-activation values already loaded by the caller
  __device__ __forceinline__ void store4_pre(float4 acc, int m, int n, const float4& res, const float4& y) const {
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if (accumulate == 2) {
      const float4 o = *reinterpret_cast<const float4*>(C + (size_t)m * ldc + n);
      v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    if (act == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = v[i] > 0.f ? v[i] : __expf(v[i]) - 1.f;
    }
    if (act_out) *reinterpret_cast<float4*>(act_out + (size_t)m * ldact + n) = make_float4(v[0], v[1], v[2], v[3]);
    if (mul_elu_out) {
      v[0] *= elu1_grad_from_out(y.x); v[1] *= elu1_grad_from_out(y.y);
      v[2] *= elu1_grad_from_out(y.z); v[3] *= elu1_grad_from_out(y.w);
    }
    if (drop_thresh) {
      const float4 d = dropout_scale4(seed(), (uint64_t)m * (uint64_t)N + (uint64_t)n, drop_thresh, drop_inv_keep);
      v[0] *= d.x; v[1] *= d.y; v[2] *= d.z; v[3] *= d.w;
    }
    if (residual) { v[0] += res.x; v[1] += res.y; v[2] += res.z; v[3] += res.w; }
    float4* c = reinterpret_cast<float4*>(C + (size_t)m * ldc + n);
    if (accumulate == 1) {
      const float4 o = *c;
      v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    *c = make_float4(v[0], v[1], v[2], v[3]);
  }
  // Feature-specialised form of store4_pre for the tensor-core epilogue.  MASK lists the features that MAY be active
  // (bit 0: activation / act_out, 1: saved-activation mask, 2: dropout, 3: residual, 4: accumulate); everything else is
  // compiled out.  The generic epilogue spent ~1300 warp instructions per 32x32 patch on warp-uniform feature tests and
  // 64-bit address arithmetic -- the GEMM kernel was issue-bound in its EPILOGUE (22 us for a 14848 x 512 output with
  // K = 32); the plain variant is ~10x shorter.
  template <int MASK>
  __device__ __forceinline__ float4 store4_masked(float4 acc, float* __restrict__ crow, float* __restrict__ act_row, uint64_t didx,
                                                  uint64_t sd, const float4& res, const float4& y) const {
    float v[4] = {acc.x, acc.y, acc.z, acc.w};
    if ((MASK & 16) && accumulate == 2) {
      const float4 o = *reinterpret_cast<const float4*>(crow);
      v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    if (MASK & 1) {
      if (act == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = v[i] > 0.f ? v[i] : __expf(v[i]) - 1.f;
      }
      if (act_row) *reinterpret_cast<float4*>(act_row) = make_float4(v[0], v[1], v[2], v[3]);
    }
    if ((MASK & 2) && mul_elu_out) {
      v[0] *= elu1_grad_from_out(y.x); v[1] *= elu1_grad_from_out(y.y);
      v[2] *= elu1_grad_from_out(y.z); v[3] *= elu1_grad_from_out(y.w);
    }
    if ((MASK & 4) && drop_thresh) {
      const float4 d = dropout_scale4(sd, didx, drop_thresh, drop_inv_keep);
      v[0] *= d.x; v[1] *= d.y; v[2] *= d.z; v[3] *= d.w;
    }
    if ((MASK & 8) && residual) { v[0] += res.x; v[1] += res.y; v[2] += res.z; v[3] += res.w; }
    if ((MASK & 16) && accumulate == 1) {
      const float4 o = *reinterpret_cast<const float4*>(crow);
      v[0] += o.x; v[1] += o.y; v[2] += o.z; v[3] += o.w;
    }
    const float4 out = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(crow) = out;
    return out;
  }
  // smallest compiled feature set that covers this problem's epilogue (0, 1, 12, 10, 31; +32 = column sums: 42 or 63)
  __device__ __forceinline__ int feature_mask() const {
    const int need = ((act != 0 || act_out) ? 1 : 0) | (mul_elu_out ? 2 : 0) | (drop_thresh ? 4 : 0) | (residual ? 8 : 0) |
                     (accumulate ? 16 : 0);
    if (colsum) return (need & ~10) == 0 ? 42 : 63;
    if (need == 0) return 0;
    if ((need & ~1) == 0) return 1;
    if ((need & ~12) == 0) return 12;
    if ((need & ~10) == 0) return 10;
    return 31;
  }
  __device__ __forceinline__ void store(float acc, int m, int n) const {
    float v = apply(acc, m, n);
    float* c = C + (size_t)m * ldc + n;
    if (accumulate == 1) v += *c;
    *c = v;
  }
};

inline Epilogue make_epilogue(const gb_gemm_args* a) {
  Epilogue e;
  e.bias = a->bias;
  e.mul_elu_out = a->mul_elu_out;
  e.residual = a->residual;
  e.C = a->C;
  e.act_out = a->act_out;
  e.ldact = a->ldact;
  e.colsum = a->colsum;
  e.ldcs = a->ld_colsum;
  e.ldc = a->ldc;
  e.ldm = a->ldm;
  e.ldr = a->ldr;
  e.N = a->N;
  e.act = a->act;
  e.accumulate = a->accumulate;
  if (a->dropout_p > 0.f) {
    double t = (double)a->dropout_p * 4294967296.0;
    e.drop_thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
    if (e.drop_thresh == 0) e.drop_thresh = 1;
    e.drop_inv_keep = 1.0f / (1.0f - a->dropout_p);
  } else {
    e.drop_thresh = 0;
    e.drop_inv_keep = 1.f;
  }
  e.drop_seed = a->dropout_seed;
  e.drop_offset = a->dropout_offset;
  return e;
}

}  // namespace gb
