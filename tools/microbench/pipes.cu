// Instruction-throughput micro-benchmark for sm_100a: warp instructions per cycle per SM for the ops the converter warps
// of the bf16x3 GEMM and the packed-pair energy kernel are built from.  nvcc -arch=sm_100a -O3 pipes.cu -o pipes && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(1024) k(uint32_t* out, uint32_t seed, float fs) {
  uint32_t r[CHAINS];
  unsigned long long q[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) { r[i] = seed + threadIdx.x * 7 + i; q[i] = ((unsigned long long)(r[i] | 0x3f800000u) << 32) | (r[i] | 0x3f800000u); }
  unsigned long long cq = ((unsigned long long)__float_as_uint(fs) << 32) | __float_as_uint(fs);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*(float*)&r[i]) : "f"(fs));
      if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(q[i]) : "l"(cq));
      if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(cq));
      if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(cq));
      if (OP == 4) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(r[i]) : "r"(seed));
      if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(seed), "r"(it));
      if (OP == 6) asm volatile("shf.r.wrap.b32 %0, %0, %1, 7;" : "+r"(r[i]) : "r"(seed));
      if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(seed));
      if (OP == 8) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r[i]) : "f"(*(float*)&r[i]), "f"(fs));
      if (OP == 9) asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(*(float*)&r[i]));
      if (OP == 10) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(seed));
      if (OP == 11) { uint32_t t; asm volatile("shr.u32 %0, %1, 16;" : "=r"(t) : "r"(r[i])); asm volatile("add.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(t), "r"(seed)); }
      if (OP == 12) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float*)&r[i]) : "f"(fs));
      if (OP == 13) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(r[i]) : "r"(seed), "r"(it));
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= r[i] ^ (uint32_t)q[i] ^ (uint32_t)(q[i] >> 32);
  if (acc == 0x12345678u) out[0] = acc;
}

template <int OP>
void run(const char* name, int ops_per_iter = 1) {
  uint32_t* d;
  cudaMalloc(&d, 4);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int warps : {4, 16, 32}) {
    dim3 grid(sms), block(32 * warps);
    k<OP><<<grid, block>>>(d, 12345u, 1.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, block>>>(d, 12345u, 1.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * clk_khz * 1e3;
    double winstr = (double)ITERS * CHAINS * warps * ops_per_iter;
    printf("%-28s warps/SM=%2d  %.3f warp-instr/clk/SM (%.2f clk per warp-instr)  [%.3f ms]\n", name, warps, winstr / cycles, cycles / winstr, ms);
  }
  cudaFree(d);
}

int main() {
  run<0>("FFMA (fma.rn.f32)");
  run<12>("FADD (add.rn.f32)");
  run<1>("FFMA2 (fma.rn.f32x2)");
  run<2>("FADD2 (add.rn.f32x2)");
  run<3>("FMUL2 (mul.rn.f32x2)");
  run<4>("PRMT");
  run<5>("LOP3");
  run<6>("SHF");
  run<7>("IADD");
  run<10>("IMAD");
  run<11>("SHR+IADD (2 ops)", 2);
  run<13>("SETP+SELP (2 ops)", 2);
  run<8>("F2FP (cvt.rn.bf16x2.f32)");
  run<9>("MUFU.RSQ");
  return 0;
}
