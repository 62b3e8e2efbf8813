"""Not a pytest file: does processing a batch of 256 as smaller micro-batches (intermediates closer
to the 126 MB L2) beat one pass?  ViT-B/16 224 px classifier, device-timed.
    python tests/bench_microbatch.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

torch.manual_seed(0)
model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=224, patch_size=16, in_channels=3,
                           embed_dim=768, num_layers=12, num_heads=12, mlp_dim=3072).cuda().eval()
x = O.synthetic_images(256, 224).cuda()


def timed(fn, iters=10):
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


with torch.no_grad():
    for mb in (256, 128, 64, 32):
        chunks = [x[i:i + mb].contiguous() for i in range(0, 256, mb)]
        ms = timed(lambda: [model(c) for c in chunks])
        print(f"256 images as {256 // mb} x {mb}: {ms:7.2f} ms  {256 / ms * 1e3:8.0f} images/s")
