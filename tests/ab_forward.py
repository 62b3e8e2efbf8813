"""Not a pytest file: times the ViT-B/16 224 px batch-256 forward and the LayerNorm kernel alone
with whatever library VITK_LIB selects (run alternately with several builds on one box)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402


def timed(fn, iters):
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


torch.manual_seed(0)
model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=224, patch_size=16, embed_dim=768,
                           num_layers=12, num_heads=12, mlp_dim=3072).cuda().eval()
x = O.synthetic_images(256, 224).cuda()
rows = torch.randn(50432, 768, device="cuda")
w, b = torch.ones(768, device="cuda"), torch.zeros(768, device="cuda")
with torch.no_grad():
    ln = timed(lambda: vitk.ops.layernorm(rows, w, b, 1e-5, torch.bfloat16), 50) * 1e3
    fwd = timed(lambda: model(x), 20)
print(os.path.basename(os.environ.get("VITK_LIB", "libvitk.so")),
      "layernorm %.1f us   forward %.3f ms" % (ln, fwd))
