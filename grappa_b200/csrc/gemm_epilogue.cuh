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
  int ldc, ldm, ldr, ldact, N;
  int act, accumulate;
  uint32_t drop_thresh;   // 0 = dropout off
  float drop_inv_keep;
  uint64_t drop_seed;

  __device__ __forceinline__ float apply(float acc, int m, int n) const {
    float v = acc;
    if (bias) v += __ldg(bias + n);
    if (act == 1) v = elu1(v);
    if (act_out) act_out[(size_t)m * ldact + n] = v;
    if (mul_elu_out) v *= elu1_grad_from_out(__ldg(mul_elu_out + (size_t)m * ldm + n));
    if (drop_thresh) v *= dropout_scale(drop_seed, (uint64_t)m * (uint64_t)N + (uint64_t)n, drop_thresh, drop_inv_keep);
    if (residual) v += __ldg(residual + (size_t)m * ldr + n);
    return v;
  }
  __device__ __forceinline__ void store(float acc, int m, int n) const {
    float v = apply(acc, m, n);
    float* c = C + (size_t)m * ldc + n;
    if (accumulate) v += *c;
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
  return e;
}

}  // namespace gb
