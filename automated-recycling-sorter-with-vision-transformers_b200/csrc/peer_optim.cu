// Data-parallel optimizer step over peer memory (NVLink 5 / NVSwitch): the gradient exchange of the
// reference-equivalent training step (losses.backward(); optimizer.step(), train.py:1455-1460, run
// data-parallel) fused with AdamW instead of an NCCL all-reduce followed by a replicated update.
//
// Every rank's gradient, parameter and bf16-shadow arenas live in symmetric memory (the same
// allocation on every GPU, mapped into every process; with NVSwitch also behind one multicast
// address).  The flat arena is cut into `world` contiguous shards; rank r owns shard r:
//   pass A  vitk_peer_reduce_scan      g[shard] = sum over ranks of grad_rank[shard]
//                                      (multimem.ld_reduce: the NVSwitch adds the eight copies and
//                                      returns one; without multicast, eight peer loads in rank
//                                      order), written to the local arena, and a non-finite value
//                                      raises the skip flag on EVERY rank (train.py:1456 semantics)
//   pass B  vitk_peer_adamw_broadcast  AdamW on the shard only (Adam moments are sharded: 1/world of
//                                      the optimizer arithmetic and state traffic per GPU), the
//                                      new fp32 parameters and their bf16 shadows stored to all
//                                      ranks at once (multimem.st, or a store per peer)
// with a cross-rank barrier before A (all backward passes done), between A and B (flags final) and
// after B (parameters in place) - the host enqueues them (symmetric-memory signal pads).  Each
// element is reduced and updated exactly once, by its owner, so the replicas stay bitwise identical
// by construction.  Per GPU and step the links carry 1/world of the gradients in and 6 bytes per
// owned parameter out (through the switch's multicast: once, not once per peer), against
// 2 * (world - 1) / world of the gradients both ways for a ring all-reduce.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"
#include "train_ops.cuh"

namespace vitk {
using namespace ptx;

namespace {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const float* grad[kMaxPeers];
  float* param[kMaxPeers];
  void* shadow[kMaxPeers];
  int* guard[kMaxPeers];
  int world;
};

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// g_local[i] = sum_r grad_r[i] for the float4s [lo4, hi4); flag = any non-finite value.
__global__ void __launch_bounds__(512)
peer_reduce_scan_kernel(const PeerPtrs a, const float* __restrict__ grad_mc,
                        float* __restrict__ grad_local, long long lo4, long long hi4, int scan) {
  bool bad = false;
  // four independent 16-byte reductions in flight per thread: a round trip through the switch is
  // several microseconds, and the links only fill with enough requests outstanding
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i0 = lo4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i0 < hi4;
       i0 += 4 * stride) {
    float4 s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      s[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < hi4) {
        if (grad_mc != nullptr) {
          s[u] = multimem_ld_reduce_add(grad_mc + 4 * i);
        } else {
          for (int r = 0; r < a.world; ++r) {   // fixed order
            const float4 v = __ldcg(reinterpret_cast<const float4*>(a.grad[r]) + i);
            s[u].x += v.x;
            s[u].y += v.y;
            s[u].z += v.z;
            s[u].w += v.w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      if (i < hi4) {
        reinterpret_cast<float4*>(grad_local)[i] = s[u];
        bad = bad || !(isfinite(s[u].x) && isfinite(s[u].y) && isfinite(s[u].z) && isfinite(s[u].w));
      }
    }
  }
  if (scan && __syncthreads_or(bad ? 1 : 0)) {
    if (threadIdx.x == 0)
      for (int r = 0; r < a.world; ++r) atomicExch_system(a.guard[r], 1);
    __threadfence_system();
  }
}

// AdamW on the owned shard (same arithmetic as adamw_kernel, train_ops.cu), results to every rank.
__global__ void __launch_bounds__(512)
peer_adamw_bcast_kernel(const PeerPtrs a, float* __restrict__ param_mc, void* __restrict__ shadow_mc,
                        const float* __restrict__ p_local, const float* __restrict__ g_local,
                        float* __restrict__ m, float* __restrict__ v, long long lo4, long long hi4,
                        float decay, float b1, float b2, float omb1, float omb2, float step_size,
                        float inv_sqrt_bc2, float eps, float grad_scale,
                        const int* __restrict__ guard, double lr, double beta1, double beta2,
                        int step) {
  if (guard != nullptr) {
    if (guard[0] != 0) return;   // a non-finite gradient somewhere: every rank skips the step
    const int skipped = guard[1];
    if (skipped != 0) {
      __shared__ float s_fix[2];
      if (threadIdx.x == 0) {
        const int t = step - skipped;
        s_fix[0] = static_cast<float>(lr / (1.0 - pow(beta1, static_cast<double>(t))));
        s_fix[1] = static_cast<float>(1.0 / sqrt(1.0 - pow(beta2, static_cast<double>(t))));
      }
      __syncthreads();
      step_size = s_fix[0];
      inv_sqrt_bc2 = s_fix[1];
    }
  }
  for (long long i = lo4 + blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < hi4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<const float4*>(p_local)[i];
    const float4 gv = reinterpret_cast<const float4*>(g_local)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gp[k] * grad_scale;
      pp[k] *= decay;
      mp[k] = b1 * mp[k] + omb1 * gr;
      vp[k] = b2 * vp[k] + omb2 * gr * gr;
      const float denom = sqrtf(vp[k]) * inv_sqrt_bc2 + eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    uint2 pk;
    pk.x = pack_bf16x2(pv.x, pv.y);
    pk.y = pack_bf16x2(pv.z, pv.w);
    if (param_mc != nullptr) {
      multimem_st(param_mc + 4 * i, pv);
      // 8 bytes of bf16 through the same multicast mapping
      asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(
                       static_cast<char*>(shadow_mc) + 8 * i),
                   "f"(__uint_as_float(pk.x)), "f"(__uint_as_float(pk.y))
                   : "memory");
    } else {
      for (int r = 0; r < a.world; ++r) {
        reinterpret_cast<float4*>(a.param[r])[i] = pv;
        reinterpret_cast<uint2*>(a.shadow[r])[i] = pk;
      }
    }
  }
  __threadfence_system();
}

int fill(const VitkPeerBuffers* pb, PeerPtrs* a) {
  VITK_REQUIRE(pb != nullptr, "peer optimizer: null buffers");
  VITK_REQUIRE(pb->world >= 1 && pb->world <= kMaxPeers && pb->rank >= 0 && pb->rank < pb->world,
               "peer optimizer: world %d / rank %d out of range (at most %d ranks)", pb->world,
               pb->rank, kMaxPeers);
  a->world = pb->world;
  for (int r = 0; r < kMaxPeers; ++r) {
    a->grad[r] = r < pb->world ? pb->grad[r] : nullptr;
    a->param[r] = r < pb->world ? pb->param[r] : nullptr;
    a->shadow[r] = r < pb->world ? pb->shadow[r] : nullptr;
    a->guard[r] = r < pb->world ? pb->guard[r] : nullptr;
  }
  for (int r = 0; r < pb->world; ++r)
    VITK_REQUIRE(a->grad[r] && a->param[r] && a->shadow[r], "peer optimizer: rank %d has null arenas", r);
  return VITK_OK;
}

}  // namespace
}  // namespace vitk

using namespace vitk;

extern "C" {

int vitk_peer_reduce_scan(const VitkPeerBuffers* pb, long long lo, long long hi,
                          vitk_stream_t stream) {
  PeerPtrs a;
  VITK_TRY(fill(pb, &a));
  VITK_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0,
               "peer_reduce_scan: the range must be multiples of 4 elements");
  if (hi == lo) return VITK_OK;
  const int scan = pb->guard[pb->rank] != nullptr;
  if (scan)
    for (int r = 0; r < pb->world; ++r)
      VITK_REQUIRE(pb->guard[r] != nullptr, "peer_reduce_scan: rank %d has no guard buffer", r);
  const long long n4 = (hi - lo) / 4;
  long long blocks = (n4 + 511) / 512;
  const long long cap = static_cast<long long>(sm_count()) * 4;
  if (blocks > cap) blocks = cap;
  ProfileScope prof(PROF_OPT, static_cast<double>(hi - lo) * 4.0 * (pb->world + 1), static_cast<cudaStream_t>(stream));
  peer_reduce_scan_kernel<<<static_cast<unsigned>(blocks), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      a, pb->grad_mc, const_cast<float*>(pb->grad[pb->rank]), lo / 4, hi / 4, scan);
  VITK_CHECK_LAUNCH("peer_reduce_scan_kernel");
  return VITK_OK;
}

int vitk_peer_adamw_broadcast(const VitkPeerBuffers* pb, float* exp_avg, float* exp_avg_sq,
                              long long lo, long long hi, double lr, double beta1, double beta2,
                              double eps, double weight_decay, int step, float grad_scale,
                              vitk_stream_t stream) {
  PeerPtrs a;
  VITK_TRY(fill(pb, &a));
  VITK_REQUIRE(exp_avg && exp_avg_sq, "peer_adamw_broadcast: null moments");
  VITK_REQUIRE(lo >= 0 && hi >= lo && lo % 4 == 0 && hi % 4 == 0,
               "peer_adamw_broadcast: the range must be multiples of 4 elements");
  VITK_REQUIRE(step >= 1, "adamw: step must be >= 1 (got %d)", step);
  VITK_REQUIRE((pb->param_mc == nullptr) == (pb->shadow_mc == nullptr),
               "peer_adamw_broadcast: multicast addresses for both parameters and shadows, or neither");
  if (hi == lo) return VITK_OK;
  const double bc1 = 1.0 - pow(beta1, step);
  const double bc2 = 1.0 - pow(beta2, step);
  const long long n4 = (hi - lo) / 4;
  long long blocks = (n4 + 511) / 512;
  const long long cap = static_cast<long long>(sm_count()) * 4;
  if (blocks > cap) blocks = cap;
  ProfileScope prof(PROF_OPT, static_cast<double>(hi - lo) * (24.0 + 6.0 * pb->world), static_cast<cudaStream_t>(stream));
  peer_adamw_bcast_kernel<<<static_cast<unsigned>(blocks), 512, 0, static_cast<cudaStream_t>(stream)>>>(
      a, pb->param_mc, pb->shadow_mc, pb->param[pb->rank], pb->grad[pb->rank], exp_avg, exp_avg_sq,
      lo / 4, hi / 4, static_cast<float>(1.0 - lr * weight_decay), static_cast<float>(beta1),
      static_cast<float>(beta2), static_cast<float>(1.0 - beta1), static_cast<float>(1.0 - beta2),
      static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)), static_cast<float>(eps),
      grad_scale, pb->guard[pb->rank], lr, beta1, beta2, step);
  VITK_CHECK_LAUNCH("peer_adamw_bcast_kernel");
  return VITK_OK;
}

}  // extern "C"
