"""Same-box timing of the LayerNorm folding (ViT-B/16 batch 256 shapes): residual GEMM with the
TMA reduce-add + stand-alone LayerNorm + consumer GEMM, against residual GEMM with statistics +
consumer GEMM that normalises in its epilogue.  CUDA events, L2 flushed between samples."""
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
EPI = vitk._lib


def timed(fn, n=12):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    return ts[len(ts) // 2], ts[0]


for K, N2, epi, name in ((768, 3072, EPI.EPI_GELU_TANH_BF16, "proj -> LN2 -> fc1"),
                         (3072, 2304, EPI.EPI_BF16, "fc2 -> LN1 -> qkv")):
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(D, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.zeros(D, device="cuda")
    W2 = torch.randn(N2, D, generator=g, device="cuda") / D ** 0.5
    b2 = torch.zeros(N2, device="cuda")
    w2 = W2.bfloat16()
    w_ln, b_ln, _ = vitk.ops.fold_layernorm(W2, gamma, beta, b2)
    out2 = torch.empty(M, N2, dtype=torch.bfloat16, device="cuda")
    for rnd in range(2):
        r = {}
        r["resid (reduce-add)"] = timed(lambda: vitk.ops.gemm(a, w, EPI.EPI_RESID_F32, bias=bias, resid=x, out=x))
        r["layernorm"] = timed(lambda: vitk.ops.layernorm(x, gamma, beta))
        xn = vitk.ops.layernorm(x, gamma, beta)
        r["consumer"] = timed(lambda: vitk.ops.gemm(xn, w2, epi, bias=b2, out=out2))
        r["resid + stats"] = timed(lambda: vitk.ops.gemm_resid_stats(a, w, x, bias=bias))
        xb, st = vitk.ops.gemm_resid_stats(a, w, x, bias=bias)
        r["consumer (folded LN)"] = timed(lambda: vitk.ops.gemm_layernorm_folded(xb, w_ln, b_ln, st, epilogue=epi))
        plain = r["resid (reduce-add)"][0] + r["layernorm"][0] + r["consumer"][0]
        fold = r["resid + stats"][0] + r["consumer (folded LN)"][0]
        print(f"{name} round {rnd}: " + ", ".join(f"{k} {v[0]:.1f} (best {v[1]:.1f})" for k, v in r.items()))
        print(f"    three launches {plain:.1f} us -> two launches {fold:.1f} us")
        x.normal_(generator=g)
