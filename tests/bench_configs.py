"""Not a pytest file: device-timed inference of the other BASELINE.json configurations.
    python tests/bench_configs.py
configs[3] ViT-L/16 224 px (batch 256 and the strong-scaling shards 128/64/32), configs[4] ViT-B/16
384 px (577 tokens, batch 64)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402


def flops_per_image(S, D, L, M, n_classes=6, p=16):
    P = (S // p) ** 2
    N = P + 1
    return 2 * P * 3 * p * p * D + L * (2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 4 * N * D * M) + 2 * D * n_classes


def run(name, B, S, D, L, H, M, iters=10):
    torch.manual_seed(0)
    model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=S, patch_size=16, in_channels=3,
                               embed_dim=D, num_layers=L, num_heads=H, mlp_dim=M).cuda().eval()
    x = O.synthetic_images(B, S).cuda()
    with torch.no_grad():
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            model(x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        vitk._lib.profile_enable(True)
        for _ in range(3):
            model(x)
        torch.cuda.synchronize()
        prof = vitk._lib.profile_collect()
        vitk._lib.profile_enable(False)
    ips = B / ms * 1e3
    tf = ips * flops_per_image(S, D, L, M) / 1e12
    kinds = {k: round(v["ms"] / 3, 2) for k, v in prof.items() if v["launches"]}
    print(f"{name:28s} B={B:4d}: {ms:8.2f} ms/step {ips:9.0f} img/s {tf:7.0f} TFLOP/s  {kinds}")
    del model
    torch.cuda.empty_cache()


run("ViT-L/16 224px", 256, 224, 1024, 24, 16, 4096)
run("ViT-L/16 224px", 128, 224, 1024, 24, 16, 4096)
run("ViT-L/16 224px", 32, 224, 1024, 24, 16, 4096)
run("ViT-B/16 384px (577 tok)", 64, 384, 768, 12, 12, 3072)
run("ViT-B/16 224px", 256, 224, 768, 12, 12, 3072)
