"""Not a pytest file: where the MMA-issuing warp of the GEMM waits (needs a trace build:
    VITK_NVCC_EXTRA=-DVITK_GEMM_TRACE python automated-.../build.py --force).
ViT-B/16 batch 256 shapes; per launch: share of the issuing warp's time spent waiting for operands
(TMA -> full barrier) and for a free accumulator stage (the epilogue)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

M, D = 197 * 256, 768
E = vitk._lib
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def trace(name, fn):
    for _ in range(2):
        flush.zero_()
        fn()
    torch.cuda.synchronize()
    buf = (C.c_longlong * 1280)()
    n = E.lib().vitk_debug_gemm_trace(buf, 1280)
    if n <= 0:
        print("not a trace build")
        sys.exit(0)
    npairs = (148 - (int(sys.argv[1]) if len(sys.argv) > 1 else 0)) // 2
    rows = [buf[8 * i:8 * i + 8] for i in range(npairs)]
    tot = sum(r[0] for r in rows) / npairs
    full = sum(r[1] for r in rows) / npairs
    empty = sum(r[2] for r in rows) / npairs
    tiles = sum(r[3] for r in rows) / npairs
    print(f"{name:28s} issuing warp: {tot:9.0f} clk total ({tot / tiles:6.0f} per tile), waits for "
          f"operands {100 * full / tot:5.1f} %, for an accumulator stage {100 * empty / tot:5.1f} %")
    ep = [sum(r[k] for r in rows) / npairs / tiles for k in (4, 5, 6, 7)]
    if any(ep):
        print(f"{'':28s} epilogue warp 0, clk per tile: waits for the accumulator {ep[0]:6.0f}, TMEM "
              f"loads {ep[1]:6.0f}, bias / activation / pack {ep[2]:6.0f}, staging + TMA store {ep[3]:6.0f}")


x = torch.randn(M, D, generator=g, device="cuda")
a768 = torch.randn(M, 768, generator=g, device="cuda").bfloat16()
a3072 = torch.randn(M, 3072, generator=g, device="cuda").bfloat16()
w = lambda n, k: (torch.randn(n, k, generator=g, device="cuda") / k ** 0.5).bfloat16()
w_qkv, w_proj, w_fc1, w_fc2 = w(2304, 768), w(768, 768), w(3072, 768), w(768, 3072)
b = lambda n: torch.zeros(n, device="cuda")
o_qkv = torch.empty(M, 2304, dtype=torch.bfloat16, device="cuda")
o_fc1 = torch.empty(M, 3072, dtype=torch.bfloat16, device="cuda")
if len(sys.argv) > 1:
    vitk._lib.check(E.lib().vitk_reserve_sms(int(sys.argv[1])))
    print("SMs kept out of the grids:", sys.argv[1])
trace("qkv (bf16 out)", lambda: vitk.ops.gemm(a768, w_qkv, E.EPI_BF16, bias=b(2304), out=o_qkv))
trace("fc1 (gelu, bf16 out)", lambda: vitk.ops.gemm(a768, w_fc1, E.EPI_GELU_TANH_BF16, bias=b(3072), out=o_fc1))
trace("fc2 (reduce-add)", lambda: vitk.ops.gemm(a3072, w_fc2, E.EPI_RESID_F32, bias=b(768), resid=x, out=x))
trace("proj (reduce-add)", lambda: vitk.ops.gemm(a768, w_proj, E.EPI_RESID_F32, bias=b(768), resid=x, out=x))
trace("fc2 (statistics)", lambda: vitk.ops.gemm_resid_stats(a3072, w_fc2, x, bias=b(768)))
trace("proj (statistics)", lambda: vitk.ops.gemm_resid_stats(a768, w_proj, x, bias=b(768)))
