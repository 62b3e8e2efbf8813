"""Not a pytest file: per-kernel-class device time of one detector training step (forward + loss +
backward through the detection head and the encoder), vitk's own per-launch CUDA events.
    python tests/prof_detector_train.py [batch] [head_dropout]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p_head = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
torch.manual_seed(0)
det = vitk.ViTObjectDetector(num_classes=6, num_queries=100, dropout=0.0, image_size=224,
                             patch_size=16, embed_dim=768, num_layers=12, num_heads=12,
                             mlp_dim=3072).cuda().train()
for m in det.modules():
    if isinstance(m, torch.nn.Dropout):
        m.p = p_head if "decoder" in str(type(m)) else m.p
for ly in det.detection_head.decoder.layers:
    for d in (ly.dropout, ly.dropout1, ly.dropout2, ly.dropout3):
        d.p = p_head
    ly.self_attn.dropout = ly.multihead_attn.dropout = p_head
x = torch.randn(B, 3, 224, 224, device="cuda")
tgt = torch.randint(0, 7, (B, 100), device="cuda")
box = torch.rand(B, 100, 4, device="cuda")
w = torch.ones(7, device="cuda")
w[-1] = 0.1


def step():
    for p in det.parameters():
        p.grad = None
    out = det(x)
    loss = vitk.weighted_cross_entropy(out["class_logits"], tgt, w) + (out["bbox_coords"] - box).abs().mean()
    loss.backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    step()
e1.record()
torch.cuda.synchronize()
print(f"batch {B}, head dropout {p_head}: {e0.elapsed_time(e1) / 3:.2f} ms per step")
vitk._lib.profile_enable(True)
step()
prof = vitk._lib.profile_collect()
vitk._lib.profile_enable(False)
for k, v in prof.items():
    if v["launches"]:
        print(f"  {k:10s} {v['ms']:8.3f} ms  {v['launches']:4d} launches")
# the head alone: forward + backward on fixed encoder features
tokens = torch.randn(B, 197, 768, device="cuda", requires_grad=True)
head = det.detection_head


def head_step():
    out = head.decode(tokens, 1)
    (out["class_logits"].sum() + out["bbox_coords"].sum()).backward()


for _ in range(2):
    head_step()
vitk._lib.profile_enable(True)
head_step()
prof = vitk._lib.profile_collect()
vitk._lib.profile_enable(False)
print("  head alone (forward + backward):")
for k, v in prof.items():
    if v["launches"]:
        print(f"  {k:10s} {v['ms']:8.3f} ms  {v['launches']:4d} launches")
