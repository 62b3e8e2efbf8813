// fp32-parity mode helpers (fp32_mode.cu).
#pragma once
#include <cuda_runtime.h>

namespace vitk {

// f32 [rows, K] -> bf16 [rows, 6K] three-term split; is_weight picks the slot order that pairs with
// the activation order; apply_gelu applies the exact erf-GELU first (fc1 -> fc2 hand-over).
int split3(const float* in, long long ld_in, void* out_bf16, long long rows, int K, int is_weight,
           int apply_gelu, cudaStream_t stream);
int patchify_f32(const float* img, float* out, int B, int C, int S, int p, cudaStream_t stream);
int attention_f32(const float* qkv, float* ctx, int B, int N, int H, int hd, cudaStream_t stream);

}  // namespace vitk
