"""Not a pytest file: prints structured diagnostics of the tcgen05 GEMM for remote debugging.
    python tests/diag_gemm.py > gpurun_out/diag_gemm.log
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

torch.manual_seed(0)


def report(name, out, ref):
    err = (out.float() - ref.float()).abs()
    bad = err > 1e-2 * max(1.0, ref.abs().max().item())
    print(f"[{name}] max_err={err.max().item():.4g} bad={int(bad.sum())}/{bad.numel()} "
          f"out_absmax={out.float().abs().max().item():.4g} ref_absmax={ref.abs().max().item():.4g}")
    if bad.any():
        M, N = bad.shape
        rows = bad.any(dim=1).nonzero().flatten().tolist()
        cols = bad.any(dim=0).nonzero().flatten().tolist()
        print(f"   bad rows ({len(rows)}): {rows[:16]}...{rows[-4:]}")
        print(f"   bad cols ({len(cols)}): {cols[:16]}...{cols[-4:]}")
        r, c = rows[0], cols[0]
        print(f"   sample out[{r},{c}:{c+8}]={out[r, c:c+8].float().tolist()}")
        print(f"   sample ref[{r},{c}:{c+8}]={ref[r, c:c+8].float().tolist()}")
    return not bad.any()


def main():
    dev = "cuda"
    ok = True
    mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    vitk._lib.set_gemm_cta_group(mode)
    print("cta_group mode", mode)
    # 1. identity-like: A = one-hot rows selecting k; B random -> out[m, n] = B[n, k(m)]
    for K in (64, 128, 256):
        M, N = 128, 256
        a = torch.zeros(M, K, device=dev)
        a[torch.arange(M), torch.arange(M) % K] = 1.0
        b = torch.randn(N, K, device=dev)
        out = vitk.ops.gemm(a.bfloat16(), b.bfloat16(), vitk._lib.EPI_F32)
        torch.cuda.synchronize()
        ok &= report(f"onehot K={K}", out, a.bfloat16().float() @ b.bfloat16().float().t())
    # 2. k-slice isolation: only one 16-wide k slice non-zero
    for ks in range(4):
        M, N, K = 128, 256, 64
        a = torch.zeros(M, K, device=dev)
        a[:, ks * 16:(ks + 1) * 16] = torch.randn(M, 16, device=dev)
        b = torch.randn(N, K, device=dev)
        out = vitk.ops.gemm(a.bfloat16(), b.bfloat16(), vitk._lib.EPI_F32)
        torch.cuda.synchronize()
        ok &= report(f"kslice {ks}", out, a.bfloat16().float() @ b.bfloat16().float().t())
    # 3. random, growing sizes
    for (M, N, K) in [(256, 256, 64), (128, 256, 64), (128, 128, 64), (256, 512, 512), (1024, 768, 768),
                      (6304, 2304, 768), (300, 400, 400)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        b = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        out = vitk.ops.gemm(a, b, vitk._lib.EPI_F32)
        torch.cuda.synchronize()
        ok &= report(f"rand {M}x{N}x{K}", out, a.float() @ b.float().t())
    print("DIAG_GEMM", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
