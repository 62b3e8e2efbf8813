// fp32-parity mode (VitkConfig.precision = 1): the SAME tcgen05 bf16 GEMM kernel evaluates fp32
// contractions by operand splitting.  Every fp32 value is written as hi + mid + lo with three bf16
// terms (3 x 8 = 24 mantissa bits); a product keeps the six partial products of weight >= 2^-16:
//     x*w ~= xh*wh + xh*wm + xm*wh + xh*wl + xl*wh + xm*wm        (dropped terms <= 2^-24 |x w|)
// Laid out along K: activations [h|h|m|h|l|m], weights [h|m|h|l|h|m]  ->  one GEMM with K' = 6K,
// bf16 products are exact in fp32 and accumulate in the fp32 TMEM accumulator.  No second GEMM
// code path, no TF32.  Used for north_star's "logits within 1e-4 in fp32" bar; throughput is not
// a goal of this mode.  Attention runs in plain fp32 FMAs.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "fp32_mode.cuh"

namespace vitk {
using namespace ptx;

namespace {

__device__ __forceinline__ float gelu_exact(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

// in f32 [rows, K] (row pitch ld_in) -> out bf16 [rows, 6K]; is_weight selects the slot order.
__global__ void __launch_bounds__(256)
split3_kernel(const float* __restrict__ in, long long ld_in, __nv_bfloat16* __restrict__ out,
              long long rows, int K, int is_weight, int apply_gelu) {
  const long long total = rows * K;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = idx / K;
    const int k = static_cast<int>(idx - r * K);
    float x = in[r * ld_in + k];
    if (apply_gelu) x = gelu_exact(x);
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(m);
    const __nv_bfloat16 l = __float2bfloat16_rn(r2);
    __nv_bfloat16* o = out + r * (6ll * K) + k;
    if (is_weight) {  // [h | m | h | l | h | m]
      o[0] = h;
      o[1ll * K] = m;
      o[2ll * K] = h;
      o[3ll * K] = l;
      o[4ll * K] = h;
      o[5ll * K] = m;
    } else {          // [h | h | m | h | l | m]
      o[0] = h;
      o[1ll * K] = h;
      o[2ll * K] = m;
      o[3ll * K] = h;
      o[4ll * K] = l;
      o[5ll * K] = m;
    }
  }
}

// images f32 NCHW -> f32 patch rows [B*P, C*p*p] (same gather as patchify_kernel, no rounding)
__global__ void __launch_bounds__(256)
patchify_f32_kernel(const float* __restrict__ img, float* __restrict__ out, int B, int C, int S,
                    int p) {
  const long long total = static_cast<long long>(B) * C * S * S;
  const int gw = S / p, P = gw * gw, Kp = C * p * p;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int x = static_cast<int>(idx % S);
    long long t = idx / S;
    const int y = static_cast<int>(t % S);
    t /= S;
    const int c = static_cast<int>(t % C);
    const int b = static_cast<int>(t / C);
    const long long row = static_cast<long long>(b) * P + (y / p) * gw + (x / p);
    const int col = c * p * p + (y % p) * p + (x % p);
    out[row * Kp + col] = img[idx];
  }
}

// fp32 attention (reference train.py:543-549 evaluated in fp32): one thread per query row, keys
// streamed through shared memory in tiles of 32, online softmax.  qkv f32 [B*N, 3D] packed.
template <int HD>
__global__ void __launch_bounds__(128)
attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, int N, int H,
                     float scale) {
  __shared__ float sK[32][HD];
  __shared__ float sV[32][HD];
  const int b = blockIdx.y / H, h = blockIdx.y - b * H;
  const int D = H * HD;
  const long long D3 = 3ll * D;
  const float* base = qkv + static_cast<long long>(b) * N * D3 + h * HD;
  const int q = blockIdx.x * 128 + threadIdx.x;
  const bool active = q < N;
  float qr[HD], o[HD];
#pragma unroll
  for (int d = 0; d < HD; ++d) {
    qr[d] = active ? base[static_cast<long long>(q) * D3 + d] : 0.f;
    o[d] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k0 = 0; k0 < N; k0 += 32) {
    __syncthreads();
    for (int i = threadIdx.x; i < 32 * HD; i += 128) {
      const int kk = i / HD, d = i - kk * HD;
      const bool ok = k0 + kk < N;
      sK[kk][d] = ok ? base[static_cast<long long>(k0 + kk) * D3 + D + d] : 0.f;
      sV[kk][d] = ok ? base[static_cast<long long>(k0 + kk) * D3 + 2 * D + d] : 0.f;
    }
    __syncthreads();
    const int kmax = min(32, N - k0);
    for (int kk = 0; kk < kmax; ++kk) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s = fmaf(qr[d], sK[kk][d], s);
      s *= scale;
      const float mn = fmaxf(m, s);
      const float corr = expf(m - mn);   // exp(-inf) = 0 on the first key
      const float pj = expf(s - mn);
      l = l * corr + pj;
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = fmaf(pj, sV[kk][d], o[d] * corr);
      m = mn;
    }
  }
  if (active) {
    const float inv = 1.f / l;
    float* dst = ctx + (static_cast<long long>(b) * N + q) * D + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) dst[d] = o[d] * inv;
  }
}

int grid_for(long long work_items, int block, int max_blocks_per_sm = 8) {
  long long g = (work_items + block - 1) / block;
  const long long cap = static_cast<long long>(sm_count()) * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int split3(const float* in, long long ld_in, void* out_bf16, long long rows, int K, int is_weight,
           int apply_gelu, cudaStream_t stream) {
  VITK_REQUIRE(in && out_bf16 && rows > 0 && K > 0 && K % 8 == 0, "split3: bad argument");
  split3_kernel<<<grid_for(rows * K, 256, 16), 256, 0, stream>>>(
      in, ld_in, static_cast<__nv_bfloat16*>(out_bf16), rows, K, is_weight, apply_gelu);
  VITK_CHECK_LAUNCH("split3_kernel");
  return VITK_OK;
}

int patchify_f32(const float* img, float* out, int B, int C, int S, int p, cudaStream_t stream) {
  VITK_REQUIRE(img && out && S % p == 0, "patchify_f32: bad argument");
  patchify_f32_kernel<<<grid_for(static_cast<long long>(B) * C * S * S, 256, 16), 256, 0, stream>>>(
      img, out, B, C, S, p);
  VITK_CHECK_LAUNCH("patchify_f32_kernel");
  return VITK_OK;
}

int attention_f32(const float* qkv, float* ctx, int B, int N, int H, int hd, cudaStream_t stream) {
  VITK_REQUIRE(qkv && ctx && B > 0 && N > 0 && H > 0, "attention_f32: bad argument");
  VITK_REQUIRE(hd == 64, "attention_f32: head_dim %d unsupported (needs 64)", hd);
  const dim3 grid((N + 127) / 128, B * H);
  attention_f32_kernel<64><<<grid, 128, 0, stream>>>(qkv, ctx, N, H,
                                                     1.0f / sqrtf(static_cast<float>(hd)));
  VITK_CHECK_LAUNCH("attention_f32_kernel");
  return VITK_OK;
}

}  // namespace vitk
