"""Wave quantisation of the N = 768 GEMMs (not a pytest file): train-step and detection-head row
counts against the inference one, CTA pairs (256-row tiles) vs single CTAs (128-row tiles).
    python tests/bench_gemm_quant.py
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from tests.bench_gemm import timeit  # noqa: E402

CASES = [(25216, 768, 768, vitk._lib.EPI_RESID_F32), (25216, 768, 3072, vitk._lib.EPI_RESID_F32),
         (25216, 768, 2304, vitk._lib.EPI_BF16), (25216, 768, 3072, vitk._lib.EPI_BF16),
         (25600, 768, 768, vitk._lib.EPI_RESID_F32), (25600, 768, 2048, vitk._lib.EPI_RESID_F32),
         (50432, 768, 768, vitk._lib.EPI_RESID_F32), (50432, 768, 3072, vitk._lib.EPI_RESID_F32),
         (25216, 2304, 768, vitk._lib.EPI_BF16), (25216, 3072, 768, vitk._lib.EPI_BF16)]
for m, n, k, epi in CASES:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    f32 = epi == vitk._lib.EPI_RESID_F32
    out = torch.zeros(m, n, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    kw = dict(out=out, bias=torch.randn(n, device="cuda"))
    if f32:
        kw["resid"] = out
    res = []
    for mode in (2, 1):
        vitk._lib.set_gemm_cta_group(mode)
        ms = timeit(lambda: vitk.ops.gemm(a, b, epi, **kw))
        res.append(f"cta{mode} {ms*1e3:7.1f} us {2*m*n*k/ms/1e9:7.1f} TF")
    vitk._lib.set_gemm_cta_group(0)
    print(f"M={m} N={n} K={k} epi={epi}: " + " | ".join(res))
