"""Not a pytest file: interleaved A/B of the two GELU epilogues on the fc1 shape."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M, N, K = B * 197, 3072, 768
a = torch.randn(M, K, device="cuda").bfloat16()
b = torch.randn(N, K, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
def t(epi, iters=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        vitk.ops.gemm(a, b, epi, bias=bias, out=out)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for e in (0, 1, 6): t(e)
res = {0: [], 1: [], 6: []}
for rep in range(8):
    for e in (0, 1, 6):
        res[e].append(t(e))
for e, name in ((0, "no gelu"), (1, "logistic (2 MUFU)"), (6, "tanh (1 MUFU)")):
    v = sorted(res[e]); print(f"{name:18s} median {v[len(v)//2]:7.1f} us  min {v[0]:7.1f}  max {v[-1]:7.1f}")
