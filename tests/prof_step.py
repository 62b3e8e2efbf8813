"""Not a pytest file: a short, fixed workload for ncu (launch list / --set full captures).
    python tests/prof_step.py [infer|train|both] [batch] [dropout]
2 forward passes of ViT-B/16 (batch 256) and/or 2 fine-tune steps (batch 128) after 1 warm-up."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "both"
kw = dict(image_size=224, patch_size=16, embed_dim=768, num_layers=12, num_heads=12, mlp_dim=3072)
torch.manual_seed(0)
dropout = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
model = vitk.ViTClassifier(num_classes=6, dropout=dropout, **kw).cuda()
if mode in ("infer", "both"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    x = O.synthetic_images(B, 224).cuda()
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            out = model(x)
    torch.cuda.synchronize()
    print("infer ok", out.shape)
if mode in ("train", "both"):
    B = 128
    x, y = O.synthetic_images(B, 224, seed=9).cuda(), O.synthetic_labels(B).cuda()
    tuner = vitk.FineTuner(model)
    for _ in range(3):
        loss, _ = tuner.step(x, y)
    torch.cuda.synchronize()
    print("train ok", loss.item())
