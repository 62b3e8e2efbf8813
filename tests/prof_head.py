"""Profiling target for the detection head alone (ncu): `python tests/prof_head.py [batch] [reps]`
runs vitk_detection_head_forward on random encoder tokens (ViT-B/16 geometry: 197 tokens, D 768)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vitk  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
head = vitk.ObjectDetectionHead(embed_dim=768, num_classes=6, num_queries=100).cuda().eval()
toks = torch.randn(B, 197, 768, device="cuda")
with torch.no_grad():
    for _ in range(reps):
        out = head.decode(toks, 1)
torch.cuda.synchronize()
print("ok", float(out["class_logits"].abs().mean()))
