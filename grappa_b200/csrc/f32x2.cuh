// Packed pair of fp32 values in one 64-bit register and the Blackwell packed-fp32 arithmetic on it (PTX add / sub / mul /
// fma .f32x2 -> SASS FADD2 / FMUL2 / FFMA2: one issue slot for two operations).
#pragma once
#include <stdint.h>

namespace gb {

struct F2 {
  unsigned long long v;
};

__device__ __forceinline__ F2 f2(float a, float b) {
  F2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ F2 f2(float a) { return f2(a, a); }
__device__ __forceinline__ float lo(F2 a) { return __uint_as_float((unsigned)(a.v & 0xffffffffull)); }
__device__ __forceinline__ float hi(F2 a) { return __uint_as_float((unsigned)(a.v >> 32)); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) {
  F2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 operator-(F2 a, F2 b) {
  F2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 operator*(F2 a, F2 b) {
  F2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
  return r;
}
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) {
  F2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return r;
}

}  // namespace gb
