"""Profiling target (ncu) for the long-sequence attention kernel: ViT-B/16 at 384 px geometry,
577 tokens, 12 heads, batch 64.  `python tests/prof_long_attn.py [reps]`"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B, N, H = 64, 577, 12
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
for _ in range(reps):
    ctx = vitk.ops.attention(qkv, B, N, H)
torch.cuda.synchronize()
print("ok", float(ctx.float().abs().mean()))
