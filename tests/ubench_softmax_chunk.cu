// Micro-benchmark: cycles per 32-column softmax chunk (32 ex2 + row-sum + bf16 pack per thread) for a
// lone warp per SM sub-partition, for several code shapes.  Guides the schedule of attention_tc2.cu.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint64_t pack2(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint32_t packbf(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }

template <int SHAPE>
__global__ void k(const float* in, uint32_t* out, float* lout, int chunks, long long* cyc) {
  __shared__ float sm[32 * 33 * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* my = sm + warp * 32 * 33;
  for (int j = 0; j < 32; ++j) my[lane * 33 + j] = in[(threadIdx.x * 32 + j) & 1023];
  __syncthreads();
  uint64_t la = 0, lb = 0, lc = 0, ld = 0;
  const uint64_t cc = pack2(0.18f, 0.18f), nm = pack2(-3.f, -3.f);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int c = 0; c < chunks; ++c) {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = my[lane * 33 + ((j + c) & 31)];
    float t[32];
#pragma unroll
    for (int j = 0; j < 16; ++j) unpack2(ffma2(pack2(v[2 * j], v[2 * j + 1]), cc, nm), t[2 * j], t[2 * j + 1]);
    uint32_t pk[16];
    if constexpr (SHAPE == 0) {  // naive: per pair ex2, ex2, add, pack
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float e0 = ex2(t[2 * j]), e1 = ex2(t[2 * j + 1]);
        if (j & 1) lb = fadd2(lb, pack2(e0, e1)); else la = fadd2(la, pack2(e0, e1));
        pk[j] = packbf(e0, e1);
      }
    } else if constexpr (SHAPE == 1) {  // two phases split by a warp barrier
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = ex2(t[j]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (j & 1) lb = fadd2(lb, pack2(t[2 * j], t[2 * j + 1])); else la = fadd2(la, pack2(t[2 * j], t[2 * j + 1]));
        pk[j] = packbf(t[2 * j], t[2 * j + 1]);
      }
    } else if constexpr (SHAPE == 2) {  // three groups split by warp barriers (half-chunk skew)
#pragma unroll
      for (int j = 0; j < 16; ++j) t[j] = ex2(t[j]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        t[16 + 2 * j] = ex2(t[16 + 2 * j]); t[17 + 2 * j] = ex2(t[17 + 2 * j]);
        if (j & 1) lb = fadd2(lb, pack2(t[2 * j], t[2 * j + 1])); else la = fadd2(la, pack2(t[2 * j], t[2 * j + 1]));
        pk[j] = packbf(t[2 * j], t[2 * j + 1]);
      }
      __syncwarp();
#pragma unroll
      for (int j = 8; j < 16; ++j) {
        if (j & 1) lb = fadd2(lb, pack2(t[2 * j], t[2 * j + 1])); else la = fadd2(la, pack2(t[2 * j], t[2 * j + 1]));
        pk[j] = packbf(t[2 * j], t[2 * j + 1]);
      }
    } else if constexpr (SHAPE == 3) {  // naive, 4 sum chains
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float e0 = ex2(t[2 * j]), e1 = ex2(t[2 * j + 1]);
        uint64_t e = pack2(e0, e1);
        if ((j & 3) == 0) la = fadd2(la, e); else if ((j & 3) == 1) lb = fadd2(lb, e); else if ((j & 3) == 2) lc = fadd2(lc, e); else ld = fadd2(ld, e);
        pk[j] = packbf(e0, e1);
      }
    } else if constexpr (SHAPE == 4) {  // no row sum at all (sum from the MMA instead)
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = packbf(ex2(t[2 * j]), ex2(t[2 * j + 1]));
    } else if constexpr (SHAPE == 5) {  // ex2 only
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = ex2(t[j]);
#pragma unroll
      for (int j = 0; j < 16; ++j) pk[j] = __float_as_uint(t[2 * j]) ^ __float_as_uint(t[2 * j + 1]);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= pk[j];
  }
  long long t1 = clock64();
  float a, b, c2, d; unpack2(fadd2(fadd2(la, lb), fadd2(lc, ld)), a, b); c2 = a + b; d = c2;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  lout[blockIdx.x * blockDim.x + threadIdx.x] = d;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int SHAPE> void run(const char* name, int warps) {
  float* in; uint32_t* out; float* lout; long long* cyc;
  cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&lout, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int chunks = 2000;
  k<SHAPE><<<148, warps * 32>>>(in, out, lout, 10, cyc);
  k<SHAPE><<<148, warps * 32>>>(in, out, lout, chunks, cyc);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-28s warps/SM=%d (per SMSP %.1f): %7.1f cycles per chunk per warp -> %.1f cycles per MUFU per SMSP\n", name, warps, warps / 4.0, (double)h / chunks,
         (double)h / chunks / 32.0 / (warps / 4.0));
}
int main() {
  for (int w : {4, 8, 16}) {
    if (w == 4) { run<0>("naive", 4); run<1>("2 phases", 4); run<2>("3 groups", 4); run<3>("naive 4 chains", 4); run<4>("no rowsum", 4); run<5>("ex2 only", 4); }
    if (w == 8) { run<0>("naive", 8); run<1>("2 phases", 8); run<2>("3 groups", 8); run<3>("naive 4 chains", 8); run<4>("no rowsum", 8); run<5>("ex2 only", 8); }
    if (w == 16) { run<0>("naive", 16); run<1>("2 phases", 16); run<2>("3 groups", 16); run<3>("naive 4 chains", 16); run<4>("no rowsum", 16); run<5>("ex2 only", 16); }
  }
  return 0;
}
