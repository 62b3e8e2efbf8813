"""A/B on one box: residual GEMM + stand-alone LayerNorm (two launches) against the same GEMM with
the LayerNorm tail fused (one launch).  ViT-B/16 batch 256 shapes: projection (K = 768) and
linear2 (K = 3072).  CUDA events, L2 flushed by the 310 MB operands of the neighbouring shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

M, D = 197 * 256, 768
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(M, D, generator=g, device="cuda")
gamma = torch.ones(D, device="cuda")
beta = torch.zeros(D, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
counters = torch.zeros((M + 127) // 128, dtype=torch.int32, device="cuda")
for K in (768, 3072):
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(D, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.zeros(D, device="cuda")
    for unfused in (True, False, True, False):
        vitk._lib.set_gemm_fused_layernorm(not unfused)
        ts = []
        for it in range(12):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            vitk.ops.gemm_resid_layernorm(a, w, x, gamma, beta, bias=bias, counters=counters)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts = sorted(ts[2:])
        print(f"K={K:5d} {'gemm + layernorm (2 launches)' if unfused else 'fused LayerNorm tail (1 launch) '}: "
              f"median {ts[len(ts) // 2]:7.1f} us  best {ts[0]:7.1f} us")
vitk._lib.set_gemm_fused_layernorm(False)
