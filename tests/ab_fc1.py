"""Not a pytest file: the linear1 + GELU GEMM of ViT-B/16 batch 256 (M 50 432, N 3072, K 768) and the
whole forward with whatever library VITK_LIB selects (alternate two builds on one box)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

M = 197 * 256
g = torch.Generator(device="cuda").manual_seed(0)
a = torch.randn(M, 768, generator=g, device="cuda").bfloat16()
w = (torch.randn(3072, 768, generator=g, device="cuda") / 768 ** 0.5).bfloat16()
b = torch.zeros(3072, device="cuda")
out = torch.empty(M, 3072, dtype=torch.bfloat16, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(14):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    vitk.ops.gemm(a, w, vitk._lib.EPI_GELU_TANH_BF16, bias=b, out=out)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
ts = sorted(ts[2:])
ref = torch.nn.functional.gelu(a[:4096].float() @ w.float().t())
err = (out[:4096].float() - ref).abs().max().item()
torch.manual_seed(0)
model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=224, patch_size=16, embed_dim=768,
                           num_layers=12, num_heads=12, mlp_dim=3072).cuda().eval()
x = O.synthetic_images(256, 224).cuda()
with torch.no_grad():
    for _ in range(3):
        model(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        model(x)
    e1.record()
    torch.cuda.synchronize()
print(os.path.basename(os.environ.get("VITK_LIB", "libvitk.so")),
      f"fc1+gelu median {ts[len(ts) // 2]:.1f} us best {ts[0]:.1f} us, max |gelu err| {err:.2e}; "
      f"forward {e0.elapsed_time(e1) / 20:.3f} ms")
