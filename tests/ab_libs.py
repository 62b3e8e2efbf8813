"""Not a pytest file: times the long-sequence attention (577 tokens, batch 64, 12 heads) and the
224 px attention with whatever library VITK_LIB selects; run alternately with two builds on one box:
    for i in 1 2 3; do VITK_LIB=$PWD/libA.so python tests/ab_libs.py; VITK_LIB=$PWD/libB.so python tests/ab_libs.py; done"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402


def timed(B, N, H, iters=30):
    qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
    for _ in range(5):
        vitk.ops.attention(qkv, B, N, H)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        vitk.ops.attention(qkv, B, N, H)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


print(os.path.basename(os.environ.get("VITK_LIB", "libvitk.so")),
      "N=577 B=64: %.1f us   N=197 B=256: %.1f us" % (timed(64, 577, 12), timed(256, 197, 12)))
