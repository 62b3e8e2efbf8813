// nn.Dropout(p) in training mode (reference train.py:528-530,545,553,563-572,650,681): every
// element is zeroed with probability p and the survivors are scaled by 1 / (1 - p).
//
// The masks are never stored: the keep decision of element `idx` of a dropout site is a pure
// function of (seed, site, layer, idx) - a counter-based hash - so the forward kernels and the
// backward kernels regenerate identical masks from the VitkConfig.seed they are both given, and
// vitk_dropout_keep_mask() materialises them for the oracle tests.  One 32-bit hash decides a
// PAIR of adjacent elements (16 bits each); p is therefore quantised to 1/65536.
// (torch's Philox stream cannot be reproduced - the reference consumes it in an implementation-
// defined order - so parity under dropout is checked with the masks injected into the oracle.)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vitk {

enum DropSite : int {
  DROP_EMBED = 0,  // tokens + position embedding          train.py:681 / evaluation.py:150
  DROP_ATTN = 1,   // attention probabilities              train.py:545
  DROP_PROJ = 2,   // attention output projection          train.py:553
  DROP_GELU = 3,   // MLP hidden activation                train.py:570
  DROP_FC2 = 4,    // MLP output                           train.py:572
  // nn.TransformerDecoderLayer(dropout=0.1) of the detection head (train.py:701-707): torch applies
  // it to the attention probabilities of both attentions, to each of the three branch outputs
  // before the residual add (dropout1 / dropout2 / dropout3) and to the ReLU output of the FFN
  DROP_DEC_SA_ATTN = 8,
  DROP_DEC_SA_OUT = 9,
  DROP_DEC_CA_ATTN = 10,
  DROP_DEC_CA_OUT = 11,
  DROP_DEC_FFN = 12,
  DROP_DEC_FF2 = 13,
};

struct DropParams {
  uint32_t key = 0;     // site key (drop_site_key)
  uint32_t thresh = 0;  // keep iff r16 >= thresh; 0 = dropout off
  float scale = 1.f;    // 1 / (1 - thresh / 65536)
};

__host__ __device__ __forceinline__ uint32_t drop_mix(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
__host__ __device__ __forceinline__ uint32_t drop_site_key(uint32_t seed, int site, int layer) {
  return drop_mix(seed * 0x9E3779B1u + static_cast<uint32_t>(site) * 0x7F4A7C15u +
                  static_cast<uint32_t>(layer) * 0x94D049BBu + 0x3C6EF372u);
}
// 2 x 16 random bits for the element pair (2 * pair_idx, 2 * pair_idx + 1)
__host__ __device__ __forceinline__ uint32_t drop_bits(uint32_t pair_idx, uint32_t key) {
  return drop_mix(pair_idx * 0x9E3779B1u + key);
}
__host__ __device__ __forceinline__ bool drop_keep_lo(uint32_t bits, uint32_t thresh) {
  return (bits & 0xFFFFu) >= thresh;
}
__host__ __device__ __forceinline__ bool drop_keep_hi(uint32_t bits, uint32_t thresh) {
  return (bits >> 16) >= thresh;
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t idx, uint32_t key, uint32_t thresh) {
  const uint32_t bits = drop_bits(idx >> 1, key);
  return (idx & 1u) ? drop_keep_hi(bits, thresh) : drop_keep_lo(bits, thresh);
}

inline DropParams make_drop_params(float p, uint32_t seed, int site, int layer) {
  DropParams d;
  if (p <= 0.f) return d;
  uint32_t t = static_cast<uint32_t>(p * 65536.f + 0.5f);
  if (t > 65535u) t = 65535u;
  if (t == 0u) return d;
  d.key = drop_site_key(seed, site, layer);
  d.thresh = t;
  d.scale = 65536.f / static_cast<float>(65536u - t);
  return d;
}

// x[j] (a run of adjacent elements starting at the EVEN element index idx0) <- dropout(x[j])
template <int COUNT>
__device__ __forceinline__ void drop_apply_run(float (&x)[COUNT], uint32_t idx0, const DropParams& d) {
#pragma unroll
  for (int j = 0; j < COUNT / 2; ++j) {
    const uint32_t bits = drop_bits((idx0 >> 1) + j, d.key);
    x[2 * j] = drop_keep_lo(bits, d.thresh) ? x[2 * j] * d.scale : 0.f;
    x[2 * j + 1] = drop_keep_hi(bits, d.thresh) ? x[2 * j + 1] * d.scale : 0.f;
  }
}

}  // namespace vitk
