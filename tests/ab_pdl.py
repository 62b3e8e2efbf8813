"""Not a pytest file: interleaved A/B of programmatic dependent launch on the ViT-B/16 forward / step."""
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


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    with torch.no_grad():
        for _ in range(n):
            model(x)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(10)
res = {0: [], 1: []}
for rep in range(6):
    for on in (1, 0):
        vitk._lib.lib().vitk_set_pdl(on)
        res[on].append(run(15))
for on in (1, 0):
    print("pdl", on, "ms/step", [round(v, 3) for v in res[on]], "median", sorted(res[on])[len(res[on]) // 2])
