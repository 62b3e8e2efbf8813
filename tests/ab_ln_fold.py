"""Not a pytest file: interleaved A/B of the ViT forward with the LayerNorm folding on and off
(vitk_set_layernorm_folding), one process, one box.
    python tests/ab_ln_fold.py [rounds] [batch]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
torch.manual_seed(0)
model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=224, patch_size=16, embed_dim=768,
                           num_layers=12, num_heads=12, mlp_dim=3072).cuda().eval()
x = O.synthetic_images(B, 224).cuda()


def timed(iters=20):
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
    return e0.elapsed_time(e1) / iters


res = {True: [], False: []}
for r in range(rounds):
    for on in (False, True):
        vitk._lib.set_layernorm_folding(on)
        res[on].append(timed())
vitk._lib.set_layernorm_folding(True)
for on in (False, True):
    v = sorted(res[on])
    print(f"LayerNorm {'folded into the GEMMs' if on else 'as separate launches'}: "
          f"median {v[len(v) // 2]:.3f} ms  best {v[0]:.3f} ms  all {[round(t, 3) for t in res[on]]}")
