// Detection head: ObjectDetectionHead.forward, evaluation.py:160-200 ==
// train.py:691-731 - `nn.TransformerDecoder` of 6 post-LN `nn.TransformerDecoderLayer`s
// (d_model = D, nhead = 8, dim_feedforward = 2048, ReLU, batch_first) decoding `num_queries`
// learned object queries against the encoder's patch tokens, then `class_head` and
// `sigmoid(bbox_head)`.  Inference (eval-mode) forward.
//
// Per layer (torch's TransformerDecoderLayer with norm_first = False):
//     x = norm1(x + self_attn(x, x, x))
//     x = norm2(x + multihead_attn(x, memory, memory))
//     x = norm3(x + linear2(relu(linear1(x))))
//
// What is done differently from the reference's op-by-op execution:
//   * the K/V projections of the encoder tokens for ALL layers are one tcgen05 GEMM
//     ([B*N, D] x [D, L*2D]): the memory does not change between layers;
//   * layer 0's self-attention sees the same queries for every image, so it is evaluated once
//     (Q rows) and broadcast by the LayerNorm that follows;
//   * residual adds are the GEMMs' fp32 reduce-add epilogue, ReLU is fc1's epilogue, every
//     LayerNorm writes the fp32 stream and its bf16 GEMM operand in one pass;
//   * attention (head_dim = D / 8 = 96 for ViT-B) never materialises [B, 8, Q, keys] scores.
#include <cuda_bf16.h>

#include <mutex>

#include "common.h"
#include "gemm_sm100.cuh"
#include "ptx.cuh"
#include "rowops.cuh"

namespace vitk {
using namespace ptx;

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------
// Post-norm LayerNorm: y = LN(in[r % in_rows]) written twice, fp32 (the residual stream the next
// GEMM reduce-adds into; may alias `in`) and bf16 (the next GEMM's operand).  in_rows < rows
// broadcasts (layer 0: the same queries for every image).  One warp per row.
// ---------------------------------------------------------------------------------------------
constexpr int kLnVec = 8;  // D <= 1024

__global__ void __launch_bounds__(256)
ln_post_kernel(const float* in, int in_rows, const float* __restrict__ gamma,
               const float* __restrict__ beta, float* out_f32, __nv_bfloat16* __restrict__ out_bf16,
               int rows, int D, float eps) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = in + static_cast<size_t>(row % in_rows) * D;
  const int nvec = D >> 2;
  float4 v[kLnVec];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      v[j] = *reinterpret_cast<const float4*>(xr + 4 * i);
      s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / static_cast<float>(D) + eps);
  float* yf = out_f32 + static_cast<size_t>(row) * D;
  __nv_bfloat16* yb = out_bf16 + static_cast<size_t>(row) * D;
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    const int i = lane + 32 * j;
    if (i < nvec) {
      const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma) + i);
      const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + i);
      float4 y;
      y.x = (v[j].x - mean) * rstd * gm.x + bt.x;
      y.y = (v[j].y - mean) * rstd * gm.y + bt.y;
      y.z = (v[j].z - mean) * rstd * gm.z + bt.z;
      y.w = (v[j].w - mean) * rstd * gm.w + bt.w;
      *reinterpret_cast<float4*>(yf + 4 * i) = y;
      uint2 pk;
      pk.x = pack_bf16x2(y.x, y.y);
      pk.y = pack_bf16x2(y.z, y.w);
      *reinterpret_cast<uint2*>(yb + 4 * i) = pk;
    }
  }
}

int ln_post(const float* in, int in_rows, const float* gamma, const float* beta, float* out_f32,
            void* out_bf16, int rows, int D, float eps, cudaStream_t stream) {
  ProfileScope prof(PROF_LN, static_cast<double>(rows) * D * 10.0, stream);
  ln_post_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(in, in_rows, gamma, beta, out_f32,
                                                     static_cast<__nv_bfloat16*>(out_bf16), rows, D,
                                                     eps);
  VITK_CHECK_LAUNCH("ln_post_kernel");
  return VITK_OK;
}

// ---------------------------------------------------------------------------------------------
// class_head and sigmoid(bbox_head) on the decoder output (evaluation.py:191-194): one warp per
// query row, fp32 weights (n_cls + 4 rows of D, L1/L2 resident), fp32 accumulation.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
det_heads_kernel(const float* __restrict__ x, const float* __restrict__ cw,
                 const float* __restrict__ cb, const float* __restrict__ bw,
                 const float* __restrict__ bb, float* __restrict__ cls_out,
                 float* __restrict__ box_out, int rows, int D, int n_cls) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  const int nvec = D >> 2;
  float4 v[kLnVec];
#pragma unroll
  for (int j = 0; j < kLnVec; ++j) {
    const int i = lane + 32 * j;
    v[j] = (i < nvec) ? xr[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int o = 0; o < n_cls + 4; ++o) {
    const float4* wr = reinterpret_cast<const float4*>(
        (o < n_cls ? cw + static_cast<size_t>(o) * D : bw + static_cast<size_t>(o - n_cls) * D));
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kLnVec; ++j) {
      const int i = lane + 32 * j;
      if (i < nvec) {
        const float4 w = __ldg(wr + i);
        acc += (v[j].x * w.x + v[j].y * w.y) + (v[j].z * w.z + v[j].w * w.w);
      }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (lane == 0) {
      if (o < n_cls) {
        cls_out[static_cast<size_t>(row) * n_cls + o] = acc + cb[o];
      } else {
        const float z = acc + bb[o - n_cls];
        box_out[static_cast<size_t>(row) * 4 + (o - n_cls)] = 1.f / (1.f + __expf(-z));
      }
    }
  }
}

struct HeadDims {
  int B, Ntok, skip, P, D, H, hd, F, L, Q, C;
  long long M, Mm;  // query rows, encoder token rows
};

struct HeadWorkspace {
  float* x;    // decoder stream f32 [M, D]
  void* xb;    // bf16 [M, D]
  void* qkv;   // bf16 [M, 3D]   (cross-attention: the q projection lives in its first D columns)
  void* ctx;   // bf16 [M, D]
  void* h;     // bf16 [M, F]
  void* mem;   // encoder tokens bf16 [Mm, D]
  void* kv;    // their K/V projections for every layer, bf16 [Mm, L*2D]
  float* xq;   // layer-0 self-attention block on the bare queries, f32 [Q, D]
  void* xqb;   // bf16 [Q, D]
  size_t bytes;
};

HeadWorkspace carve(const HeadDims& d, void* base) {
  HeadWorkspace w;
  size_t off = 0;
  auto take = [&](size_t n) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(n, 1024);
    return p;
  };
  const size_t D = d.D;
  w.x = static_cast<float*>(take(d.M * D * 4));
  w.xb = take(d.M * D * 2);
  w.qkv = take(d.M * 3 * D * 2);
  w.ctx = take(d.M * D * 2);
  w.h = take(d.M * static_cast<size_t>(d.F) * 2);
  w.mem = take(d.Mm * D * 2);
  w.kv = take(d.Mm * static_cast<size_t>(d.L) * 2 * D * 2);
  w.xq = static_cast<float*>(take(static_cast<size_t>(d.Q) * D * 4));
  w.xqb = take(static_cast<size_t>(d.Q) * D * 2);
  w.bytes = off;
  return w;
}

int check(const VitkDetectionHeadConfig* cfg, int batch, int n_tokens, int skip, HeadDims* d) {
  VITK_REQUIRE(cfg != nullptr, "detection head: config is null");
  VITK_REQUIRE(batch > 0 && n_tokens > 0 && skip >= 0 && skip < n_tokens,
               "detection head: bad batch %d / tokens %d / skipped prefix %d", batch, n_tokens, skip);
  VITK_REQUIRE(cfg->embed_dim > 0 && cfg->num_heads > 0 && cfg->embed_dim % cfg->num_heads == 0,
               "detection head: embed_dim %d must be divisible by num_heads %d", cfg->embed_dim,
               cfg->num_heads);
  d->hd = cfg->embed_dim / cfg->num_heads;
  VITK_REQUIRE(d->hd == 32 || d->hd == 64 || d->hd == 96 || d->hd == 128,
               "detection head: head_dim %d unsupported (32, 64, 96 or 128)", d->hd);
  VITK_REQUIRE(cfg->embed_dim % 8 == 0 && cfg->embed_dim <= 128 * kLnVec && cfg->ffn_dim > 0 &&
                   cfg->ffn_dim % 8 == 0,
               "detection head: embed_dim (<= %d) and ffn_dim must be multiples of 8", 128 * kLnVec);
  VITK_REQUIRE(cfg->num_layers > 0 && cfg->num_queries > 0 && cfg->num_outputs > 0,
               "detection head: layers, queries and outputs must be positive");
  d->B = batch;
  d->Ntok = n_tokens;
  d->skip = skip;
  d->P = n_tokens - skip;
  d->D = cfg->embed_dim;
  d->H = cfg->num_heads;
  d->F = cfg->ffn_dim;
  d->L = cfg->num_layers;
  d->Q = cfg->num_queries;
  d->C = cfg->num_outputs;
  d->M = static_cast<long long>(batch) * d->Q;
  d->Mm = static_cast<long long>(batch) * n_tokens;
  VITK_REQUIRE(d->M < (1ll << 29) && d->Mm < (1ll << 29), "detection head: batch too large");
  return VITK_OK;
}

int linear(const void* A, int lda, const void* W, int M, int N, int K, GemmEpi epi,
           const float* bias, float* resid_out, void* out, int ldo, cudaStream_t stream) {
  GemmProblem p;
  p.A = A;
  p.lda = lda;
  p.B = W;
  p.ldb = K;
  p.M = M;
  p.N = N;
  p.K = K;
  p.epi = epi;
  p.e.bias = bias;
  p.e.resid = resid_out;
  p.e.ldr = N;
  p.e.out = (epi == EPI_RESID_F32) ? static_cast<void*>(resid_out) : out;
  p.e.ldo = ldo;
  return gemm_bf16_tn(p, stream);
}

int attn_x(const HeadDims& d, int B, const void* q, long long q_img, int ldq, const void* k,
           const void* v, long long kv_img, int ldkv, int Nk, void* ctx, cudaStream_t stream) {
  const long long ctx_img = static_cast<long long>(d.Q) * d.D;
  // tcgen05 kernel when the queries are one MMA tile and the keys one block (the reference's 100
  // queries against 100 / 196 keys); attention impl 1 forces the mma.sync kernel (tests, A/B)
  if (attention_impl() != 1 && attention_xtc_applicable(q_img, kv_img, B, d.Q, Nk, d.hd))
    return attention_xtc(q, q_img, ldq, k, v, kv_img, ldkv, ctx, ctx_img, d.D, B, d.Q, Nk, d.H, d.hd,
                         stream);
  return attention_x(q, q_img, ldq, k, v, kv_img, ldkv, ctx, ctx_img, d.D, B, d.Q, Nk, d.H, d.hd,
                     stream);
}

int head_forward(const VitkDetectionHeadConfig* cfg, const VitkDetectionHeadWeights* w,
                 const float* tokens, const HeadDims& d, float* class_logits, float* bbox,
                 const HeadWorkspace& ws, cudaStream_t stream) {
  const int D = d.D, M = static_cast<int>(d.M), Q = d.Q;
  const __nv_bfloat16* qkv = static_cast<const __nv_bfloat16*>(ws.qkv);
  const __nv_bfloat16* kv = static_cast<const __nv_bfloat16*>(ws.kv);
  const int ldkv = d.L * 2 * D;
  // -- memory = features[:, skip:, :] (evaluation.py:235): bf16 operand, then K and V of every
  //    layer's multihead_attn in one GEMM (rows of the skipped prefix tokens are computed and
  //    never read)
  VITK_TRY(cast_f32_to_bf16(tokens, ws.mem, d.Mm * D, stream));
  VITK_TRY(linear(ws.mem, D, w->ca_kv_w, static_cast<int>(d.Mm), ldkv, D, EPI_BF16, w->ca_kv_b,
                  nullptr, ws.kv, ldkv, stream));
  for (int l = 0; l < d.L; ++l) {
    const VitkDecoderLayerWeights& lw = w->layers[l];
    // ---- x = norm1(x + self_attn(x, x, x))
    if (l == 0) {
      // object_queries.unsqueeze(0).expand(B, -1, -1) (evaluation.py:186): identical for every
      // image, so the block runs on Q rows and norm1 broadcasts its result
      VITK_CHECK_CUDA(cudaMemcpyAsync(ws.xq, w->object_queries, static_cast<size_t>(Q) * D * 4,
                                      cudaMemcpyDeviceToDevice, stream));
      VITK_TRY(cast_f32_to_bf16(w->object_queries, ws.xqb, static_cast<long long>(Q) * D, stream));
      VITK_TRY(linear(ws.xqb, D, lw.sa_in_w, Q, 3 * D, D, EPI_BF16, lw.sa_in_b, nullptr, ws.qkv,
                      3 * D, stream));
      VITK_TRY(attn_x(d, 1, qkv, 0, 3 * D, qkv + D, qkv + 2 * D, 0, 3 * D, Q, ws.ctx, stream));
      VITK_TRY(linear(ws.ctx, D, lw.sa_out_w, Q, D, D, EPI_RESID_F32, lw.sa_out_b, ws.xq, nullptr,
                      D, stream));
      VITK_TRY(ln_post(ws.xq, Q, lw.norm1_w, lw.norm1_b, ws.x, ws.xb, M, D, cfg->ln_eps, stream));
    } else {
      VITK_TRY(linear(ws.xb, D, lw.sa_in_w, M, 3 * D, D, EPI_BF16, lw.sa_in_b, nullptr, ws.qkv,
                      3 * D, stream));
      const long long img = static_cast<long long>(Q) * 3 * D;
      VITK_TRY(attn_x(d, d.B, qkv, img, 3 * D, qkv + D, qkv + 2 * D, img, 3 * D, Q, ws.ctx, stream));
      VITK_TRY(linear(ws.ctx, D, lw.sa_out_w, M, D, D, EPI_RESID_F32, lw.sa_out_b, ws.x, nullptr, D,
                      stream));
      VITK_TRY(ln_post(ws.x, M, lw.norm1_w, lw.norm1_b, ws.x, ws.xb, M, D, cfg->ln_eps, stream));
    }
    // ---- x = norm2(x + multihead_attn(x, memory, memory))
    VITK_TRY(linear(ws.xb, D, lw.ca_q_w, M, D, D, EPI_BF16, lw.ca_q_b, nullptr, ws.qkv, D, stream));
    {
      const __nv_bfloat16* kl = kv + static_cast<size_t>(d.skip) * ldkv + static_cast<size_t>(l) * 2 * D;
      VITK_TRY(attn_x(d, d.B, qkv, static_cast<long long>(Q) * D, D, kl, kl + D,
                      static_cast<long long>(d.Ntok) * ldkv, ldkv, d.P, ws.ctx, stream));
    }
    VITK_TRY(linear(ws.ctx, D, lw.ca_out_w, M, D, D, EPI_RESID_F32, lw.ca_out_b, ws.x, nullptr, D,
                    stream));
    VITK_TRY(ln_post(ws.x, M, lw.norm2_w, lw.norm2_b, ws.x, ws.xb, M, D, cfg->ln_eps, stream));
    // ---- x = norm3(x + linear2(relu(linear1(x))))
    VITK_TRY(linear(ws.xb, D, lw.ff1_w, M, d.F, D, EPI_RELU_BF16, lw.ff1_b, nullptr, ws.h, d.F,
                    stream));
    VITK_TRY(linear(ws.h, d.F, lw.ff2_w, M, D, d.F, EPI_RESID_F32, lw.ff2_b, ws.x, nullptr, D,
                    stream));
    VITK_TRY(ln_post(ws.x, M, lw.norm3_w, lw.norm3_b, ws.x, ws.xb, M, D, cfg->ln_eps, stream));
  }
  {
    ProfileScope prof(PROF_OTHER, static_cast<double>(M) * D * 4.0, stream);
    det_heads_kernel<<<(M + 7) / 8, 256, 0, stream>>>(ws.x, w->class_w, w->class_b, w->bbox_w,
                                                      w->bbox_b, class_logits, bbox, M, D, d.C);
    VITK_CHECK_LAUNCH("det_heads_kernel");
  }
  return VITK_OK;
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" {

int vitk_detection_head_workspace_bytes(const VitkDetectionHeadConfig* cfg, int batch, int n_tokens,
                                        int skip_tokens, size_t* out_bytes) {
  HeadDims d;
  VITK_TRY(check(cfg, batch, n_tokens, skip_tokens, &d));
  VITK_REQUIRE(out_bytes != nullptr, "detection head: out_bytes is null");
  *out_bytes = carve(d, nullptr).bytes;
  return VITK_OK;
}

int vitk_detection_head_forward(const VitkDetectionHeadConfig* cfg,
                                const VitkDetectionHeadWeights* w, const float* tokens, int batch,
                                int n_tokens, int skip_tokens, float* class_logits_out,
                                float* bbox_out, void* workspace, size_t workspace_bytes,
                                vitk_stream_t stream) {
  HeadDims d;
  VITK_TRY(check(cfg, batch, n_tokens, skip_tokens, &d));
  VITK_REQUIRE(w != nullptr && w->layers != nullptr && w->object_queries != nullptr &&
                   w->ca_kv_w != nullptr && w->class_w != nullptr && w->bbox_w != nullptr,
               "detection head: weights are null");
  VITK_REQUIRE(tokens && class_logits_out && bbox_out && workspace,
               "detection head: null tokens / outputs / workspace");
  VITK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0,
               "detection head: workspace must be 1024-byte aligned");
  VITK_REQUIRE(device_cc() >= 100, "detection head: requires an sm_100 device (found sm_%d)",
               device_cc());
  const HeadWorkspace ws = carve(d, workspace);
  if (workspace_bytes < ws.bytes)
    return set_error(VITK_ERR_WORKSPACE, "detection head: workspace of %zu bytes < %zu required",
                     workspace_bytes, ws.bytes);
  return head_forward(cfg, w, tokens, d, class_logits_out, bbox_out, ws,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"
