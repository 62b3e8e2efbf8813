// Host-visible declarations for the tcgen05 GEMM (gemm_sm100.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "dropout.cuh"

namespace vitk {

// Epilogue selector (compile-time variants of one kernel).
enum GemmEpi : int {
  EPI_BF16 = 0,        // out_bf16 = acc + bias
  EPI_GELU_BF16 = 1,   // out_bf16 = gelu_erf(acc + bias); optional out2_bf16 = acc + bias
  EPI_RESID_F32 = 2,   // out_f32 = acc + bias + resid_f32   (optional row remap: patch embed)
  EPI_F32 = 3,         // out_f32 = alpha * acc + bias (+ beta * out_f32)
  EPI_DGELU_BF16 = 4,  // out_bf16 = (acc + bias) * gelu'(aux_bf16)            (fc2 dgrad)
  EPI_GELU_TANH_BF16 = 6,  // as EPI_GELU_BF16 with the tanh.approx form of the normal CDF (A/B)
  EPI_RELU_BF16 = 7,   // out_bf16 = max(acc + bias, 0)   (decoder feed-forward, evaluation.py:170-176)
  // x = x + acc + bias in place (out == resid; the tile of x is TMA-loaded into the staging slab,
  // updated there and TMA-stored back), out2_bf16 = bf16(x) and per-row partial sums (sum x,
  // sum x^2) of the updated values -> ln_part_out: what the LayerNorm that follows needs, without
  // a pass of its own over x (train.py:586-591).  TMA epilogue only.
  EPI_RESID_STATS_F32 = 8,
};

struct GemmEpilogue {
  const float* bias = nullptr;   // [N] or null
  const float* resid = nullptr;  // EPI_RESID_F32: fp32 [*, ldr]
  const void* aux = nullptr;     // EPI_DGELU_BF16: bf16 [M, ldo] pre-activation
  void* out = nullptr;           // [M(out rows), ldo]
  void* out2 = nullptr;          // EPI_GELU_BF16 optional pre-activation copy (bf16, ldo)
  int ldo = 0;                   // elements
  int ldr = 0;                   // elements
  // Row remap (token assembly, reference evaluation.py:145-148 / train.py:673-678):
  //   out_row   = (m / rows_per_group) * group_stride + group_offset + m % rows_per_group
  //   resid_row = group_offset + m % rows_per_group        (position-embedding row)
  // rows_per_group == 0 disables the remap (out_row = resid_row = m).
  int rows_per_group = 0;
  int group_stride = 0;
  int group_offset = 0;
  float alpha = 1.f;
  float beta = 0.f;
  // nn.Dropout on the epilogue's result, element index = row * N + column (dropout.cuh):
  //   EPI_RESID_F32: out += dropout(acc + bias)          (projection / linear2 output dropout)
  //   EPI_GELU_*   : out = dropout(gelu(acc + bias))     (out2 keeps the undropped pre-activation)
  //   EPI_DGELU    : out = dropout_mask(acc) * gelu'(aux) (the backward of the above)
  DropParams drop;
  // LayerNorm of the updated rows (EPI_RESID_F32; reference train.py:586-591: x = x + f(..);
  // LN(x) feeds the next Linear): ln_out = LN(out) * ln_gamma + ln_beta as bf16.  By default a
  // layernorm_fwd launch behind the GEMM.  With gemm_set_fused_layernorm(1), the in-place TMA
  // reduce-add form and ln_counters != nullptr, one launch: every epilogue warp counts its
  // finished column tile of a 128-row block in ln_counters[row / 128]; the CTA that completes the
  // count fetches those rows of `out` (still L2-resident) into shared memory by bulk copies and
  // its LayerNorm warps normalise them.  Same arithmetic, lane for lane, as layernorm_fwd_kernel:
  // both forms produce the same bits.  ln_out == nullptr: no LayerNorm.
  // ln_counters: zero-filled uint32[ceil(M / 128)], left zero-filled.
  void* ln_out = nullptr;         // bf16 [M, ln_ldo]
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float* ln_mean = nullptr;       // optional [M]
  float* ln_rstd = nullptr;       // optional [M]
  float* ln_xcopy = nullptr;      // optional fp32 [M, N] copy of the updated rows (saved for backward)
  unsigned int* ln_counters = nullptr;
  int ln_ldo = 0;
  float ln_eps = 1e-5f;
  // ---- LayerNorm folded into the two GEMMs around it (inference).
  // Producer (EPI_RESID_STATS_F32): every epilogue thread owns a row and 128 (BLOCK_N / 2) of its
  // columns, so the partial sums need no exchange: ln_part_out[(column group) * M + row] =
  // (sum x, sum x^2) over that group; gemm_stats_parts(N) groups per row, summed by the consumer
  // in a fixed order (deterministic, no atomics).
  float2* ln_part_out = nullptr;
  // Consumer (EPI_BF16 / EPI_GELU_*): A holds bf16(x) (NOT normalised), B holds the row-centred
  // bf16(W * gamma - mean_k(W * gamma)): a zero-sum weight row makes the contraction itself drop
  // the row mean of x, so with rstd of the row from ln_part[0 .. ln_nparts) (M entries each;
  // ln_dim features) out = rstd * acc + bias[n], bias[n] = b[n] + sum_k W[n,k] beta[k]
  // (ln_fold, rowops.cuh) == Linear(LayerNorm(x)).
  const float2* ln_part = nullptr;
  int ln_nparts = 0;
  int ln_dim = 0;
};

struct GemmProblem {
  const void* A = nullptr;  // bf16 [M, lda], K contiguous
  const void* B = nullptr;  // bf16 [N, ldb], K contiguous  (C = A * B^T)
  int M = 0, N = 0, K = 0;
  int lda = 0, ldb = 0;     // elements
  // mn_major: A is stored [K, lda] with M contiguous and B [K, ldb] with N contiguous
  // (C = A^T B, the weight-gradient contraction over tokens).  EPI_F32 only.
  bool mn_major = false;
  int split_k = 1;          // K range split across work items; needs EPI_F32 with beta == 1
  GemmEpi epi = EPI_BF16;
  GemmEpilogue e;
};

// Returns 0 on success, else a vitk error code (message via vitk_last_error()).
int gemm_bf16_tn(const GemmProblem& p, cudaStream_t stream);

// 0 = choose automatically (CTA pairs whenever M > 128), 1 / 2 = force the cta_group (tests, A/B).
void gemm_force_cta_group(int ctas);
// 1 = always use the register->global epilogue instead of the smem-staged TMA store/reduce.
void gemm_force_direct_epilogue(int on);
// Trace builds (-DVITK_GEMM_TRACE): copies [pair][{total, wait operands, wait accumulator stage,
// tiles}] SM-clock counts of the last GEMM launch into out (n entries); returns the entries
// written, 0 when the library is not a trace build.
int gemm_debug_trace(long long* out, int n);
// Column groups per row of EPI_RESID_STATS_F32's partial sums for an N-column output.
int gemm_stats_parts(int N);
// GemmEpilogue::ln_out: 0 (default) = residual GEMM, then a separate layernorm_fwd launch;
// 1 = the LayerNorm tail inside the GEMM kernel.  Same bits either way.
void gemm_set_fused_layernorm(int on);
}  // namespace vitk
