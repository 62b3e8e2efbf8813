"""Not a pytest file: device-timed attention forward per implementation.
    python tests/bench_attn.py [batch] [tokens] [heads]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H = int(sys.argv[3]) if len(sys.argv) > 3 else 12
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for e0, e1 in e:
        flush.zero_()          # evict qkv from L2 between launches
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in e)
    return t[len(t) // 2]


for impl, name in ((1, "flash mma.sync"), (3, "tcgen05 unpipelined"), (2, "tcgen05 pipelined"),
                   (4, "tcgen05 long-seq")):
    if (impl == 3 and N > 256) or (impl == 4 and N > 640):
        continue
    vitk._lib.set_attention_impl(impl)
    ms = timeit(lambda: vitk.ops.attention(qkv, B, N, H))
    fl = 4.0 * B * H * N * N * 64
    by = B * N * H * 64 * 2 * 4
    print(f"{name:22s} B={B} N={N} H={H}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  {by/ms/1e6:7.1f} GB/s")
vitk._lib.set_attention_impl(0)

# ---- backward (tcgen05 kernel vs the mma.sync one)
if N <= 256:
    ctx, lse = vitk.ops.attention(qkv, B, N, H, return_lse=True)
    dctx = (torch.randn_like(ctx.float()) * 0.1).bfloat16()
    for impl, name in ((1, "bwd mma.sync"), (2, "bwd tcgen05")):
        vitk._lib.set_attention_impl(impl)
        ms = timeit(lambda: vitk.ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H))
        fl = 10.0 * B * H * N * N * 64
        print(f"{name:22s} B={B} N={N} H={H}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s")
    vitk._lib.set_attention_impl(0)
