"""Not a pytest file: device-timed TFLOP/s of the tcgen05 GEMM for the encoder's shapes.
    python tests/bench_gemm.py [batch]   (both cta_group modes)
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M = B * 197
SHAPES = [("qkv", M, 2304, 768, vitk._lib.EPI_BF16), ("proj", M, 768, 768, vitk._lib.EPI_RESID_F32),
          ("fc1", M, 3072, 768, vitk._lib.EPI_GELU_BF16), ("fc2", M, 768, 3072, vitk._lib.EPI_RESID_F32),
          ("fc1_nogelu", M, 3072, 768, vitk._lib.EPI_BF16), ("fc1_tanh", M, 3072, 768, 6)]


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    for mode in (1, 2):
        vitk._lib.set_gemm_cta_group(mode)
        for name, m, n, k, epi in SHAPES:
            a = torch.randn(m, k, device="cuda").bfloat16()
            b = torch.randn(n, k, device="cuda").bfloat16()
            bias = torch.randn(n, device="cuda")
            f32 = epi == vitk._lib.EPI_RESID_F32
            out = torch.zeros(m, n, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
            kw = dict(bias=bias, out=out)
            if f32:
                kw["resid"] = out
            ms = timeit(lambda: vitk.ops.gemm(a, b, epi, **kw))
            print(f"cta{mode} {name:11s} M={m} N={n} K={k}: {ms*1e3:8.1f} us  "
                  f"{2*m*n*k/ms/1e9:8.1f} TFLOP/s")
    vitk._lib.set_gemm_cta_group(0)
    # cuBLAS reference point for the same shape (library baseline, not part of the product)
    for name, m, n, k, epi in SHAPES[:4]:
        a = torch.randn(m, k, device="cuda").bfloat16()
        b = torch.randn(n, k, device="cuda").bfloat16()
        ms = timeit(lambda: torch.matmul(a, b.t()))
        print(f"cublas {name:11s}: {ms*1e3:8.1f} us  {2*m*n*k/ms/1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
