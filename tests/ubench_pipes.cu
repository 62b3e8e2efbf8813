#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// Per-SM issue throughput of candidate epilogue / softmax instructions (sm_100a).
template <int OP>
__global__ void k(float* out, int iters, float seed) {
  float a[8];
  uint32_t u[8];
  unsigned long long d[4];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.01f + threadIdx.x * 1e-4f; u[i] = __float_as_uint(a[i]) ; }
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = ((unsigned long long)u[2*i] << 32) | u[2*i+1];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if constexpr (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if constexpr (OP == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u[i]));
      if constexpr (OP == 3) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if constexpr (OP == 4) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u[i]));
      if constexpr (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u[i]));
      if constexpr (OP == 6) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if constexpr (OP == 7) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if constexpr (OP == 8) asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(d[i & 3]));
      if constexpr (OP == 9) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i+1)&7]));
      if constexpr (OP == 10) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i+1)&7]), "f"(a[(i+2)&7]));
      if constexpr (OP == 11) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i+1)&7])); a[i] = __uint_as_float(u[i]); }
      if constexpr (OP == 15) { asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(u[i]) : "r"(u[i]), "r"(u[(i+1)&7])); }
      if constexpr (OP == 16) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(a[i]), "f"(a[(i+1)&7])); a[i] = __uint_as_float(u[i]); }
      if constexpr (OP == 17) { asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i])); }
      if constexpr (OP == 18) { asm volatile("cvt.rni.s32.f32 %0, %1;" : "=r"(u[i]) : "f"(a[i])); a[i] = __uint_as_float(u[i]); }
      if constexpr (OP == 12) asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(d[i & 3]));
      if constexpr (OP == 13) asm volatile("fma.rn.bf16x2 %0, %0, %0, %0;" : "+r"(u[i]));
      if constexpr (OP == 14) asm volatile("mul.rn.f32x2 %0, %0, %0;" : "+l"(d[i & 3]));
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) s += (float)d[i];
  if (s == 123.456f) out[0] = s;
}
template <int OP>
void run(const char* name, int per_it) {
  float* o; cudaMalloc(&o, 4);
  int iters = 4096, threads = 512, blocks = 148;
  k<OP><<<blocks, threads>>>(o, 16, 0.5f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<OP><<<blocks, threads>>>(o, iters, 0.5f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double ops = (double)iters * per_it * threads;  // thread-instructions per SM
  double cyc = ms * 1e-3 * clk * 1e3;
  printf("%-22s %8.2f thread-instr/clk/SM (assuming %d MHz; %.3f ms)\n", name, ops / cyc, clk / 1000, ms);
  cudaFree(o);
}
int main() {
  run<0>("ex2.f32", 8); run<1>("ex2.f16x2", 8); run<2>("ex2.bf16x2", 8);
  run<3>("tanh.f32", 8); run<4>("tanh.f16x2", 8); run<5>("tanh.bf16x2", 8);
  run<6>("rcp.f32", 8); run<7>("fma.f32", 8); run<8>("fma.f32x2", 8);
  run<9>("max.f32", 8); run<10>("max3.f32", 8); run<11>("cvt.bf16x2.f32", 8);
  run<12>("add.f32x2", 8); run<13>("fma.bf16x2", 8); run<14>("mul.f32x2", 8);
  run<15>("prmt", 8); run<16>("cvt.f16x2.f32", 8); run<17>("lg2", 8); run<18>("cvt.rni.s32.f32", 8);
  return 0;
}
