"""Not a pytest file: interleaved A/B of the full classifier forward against the CLS-only tail
(vitk_forward_cls) on ViT-B/16 224 px, batch 256."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

kw = dict(image_size=224, patch_size=16, embed_dim=768, num_layers=12, num_heads=12, mlp_dim=3072)
torch.manual_seed(0)
model = vitk.ViTClassifier(num_classes=6, dropout=0.0, **kw).cuda().eval()
x = O.synthetic_images(256, 224).cuda()


def run(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(n):
            fn(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(model, 10)
run(model.classify_pruned, 5)
res = {"full": [], "cls-only tail": []}
for rep in range(6):
    res["cls-only tail"].append(run(model.classify_pruned, 15))
    res["full"].append(run(model, 15))
for k, v in res.items():
    print(f"{k:14s} ms/step", [round(t, 3) for t in v], "median", round(sorted(v)[len(v) // 2], 3))
