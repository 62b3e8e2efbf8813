// Micro-benchmark: issue / execution cost of small tcgen05.mma instructions (one CTA per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I automated-recycling-sorter-with-vision-transformers_b200/csrc
// Variants: A from smem (SS) or TMEM (TS), N in {64, 128, 208, 256}, one or two issuing warps.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vitk::ptx;

struct Result { long long issue, total; };

template <bool TS, int N, int WARPS>
__global__ void __launch_bounds__(128, 1) k(Result* out, int count) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + 16384, bar0 = base + 16384 + 32768, slot = bar0 + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) { mbar_init(bar0, WARPS); mbar_init(bar0 + 8, 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  const bool issuer = (warp == 0) || (WARPS == 2 && warp == 1);
  long long t0 = 0, t1 = 0, t2 = 0;
  if (issuer) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, TS ? 1 : 0);
    const uint64_t a_desc = make_desc_sw128(sA, 16, 1024);
    const uint64_t b_desc = TS ? make_desc_sw128(sB, 32768, 1024) : make_desc_sw128(sB, 16, 1024);
    const uint32_t d = tmem_base + 256 + (warp == 1 ? 0 : 0);   // both issuers accumulate into the same tile
    __syncwarp();
    t0 = clock64();
    if (elect_one_sync()) {
      for (int i = 0; i < count / WARPS; ++i) {
        if (TS) mma_bf16_ts(d, tmem_base + (i & 7) * 8, b_desc + (uint64_t)(i & 3) * 128u, idesc, 1u);
        else mma_bf16_ss(d, a_desc + 2u * (i & 3), b_desc + 2u * (i & 3), idesc, 1u);
      }
      mma_commit(bar0);
    }
    __syncwarp();
    t1 = clock64();
    mbar_wait(bar0, 0);
    t2 = clock64();
    if (warp == 0 && lane == 0 && blockIdx.x == 0) { out->issue = t1 - t0; out->total = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

template <bool TS, int N, int WARPS> void run(const char* name) {
  Result* d; cudaMalloc(&d, sizeof(Result));
  auto kern = k<TS, N, WARPS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int count = 512;
  kern<<<148, 128, 60000>>>(d, 16);
  kern<<<148, 128, 60000>>>(d, count);
  Result h; cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  printf("%-28s %4d MMAs: issue %7.1f cycles/MMA, issue->retire %7.1f cycles/MMA (floor %d)  %s\n", name, count,
         (double)h.issue / count, (double)h.total / count, 128 * N / 256, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<false, 64, 1>("SS N=64  1 issuer");
  run<true, 64, 1>("TS N=64  1 issuer");
  run<true, 64, 2>("TS N=64  2 issuers");
  run<false, 128, 1>("SS N=128 1 issuer");
  run<false, 128, 2>("SS N=128 2 issuers");
  run<false, 208, 1>("SS N=208 1 issuer");
  run<false, 256, 1>("SS N=256 1 issuer");
  run<true, 128, 1>("TS N=128 1 issuer");
  return 0;
}
