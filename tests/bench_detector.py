"""Device timing of the detector (ViT-B/16 backbone + detection head, evaluation.py:203-241):

    python tests/bench_detector.py [batch] [image_size]

CUDA events around `iters` back-to-back calls after warm-up; per-kernel-class breakdown of the
head from the library's own event profiler."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import vitk  # noqa: E402


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 224
    torch.manual_seed(0)
    det = vitk.ViTObjectDetector(image_size=S, num_classes=6, num_queries=100).cuda().eval()
    x = torch.randn(B, 3, S, S, device="cuda")
    P = (S // 16) ** 2
    D, Q, F, L = 768, 100, 2048, 6
    head_flops = B * (2 * (P + 1) * D * L * 2 * D          # K/V projection of the memory
                      + L * (2 * Q * D * 3 * D + 4 * Q * Q * D + 2 * Q * D * D      # self-attention
                             + 2 * Q * D * D + 4 * Q * P * D + 2 * Q * D * D        # cross-attention
                             + 4 * Q * D * F)                                       # feed-forward
                      + 2 * Q * D * 11)
    with torch.no_grad():
        toks = det.backbone(x)
        t_b = timed(lambda: det.backbone(x))
        t_h = timed(lambda: det.detection_head.decode(toks, 1))
        t_d = timed(lambda: det(x))
        vitk._lib.profile_enable(True)
        det.detection_head.decode(toks, 1)
        torch.cuda.synchronize()
        prof = vitk._lib.profile_collect()
        vitk._lib.profile_enable(False)
    print(f"batch {B} image {S}: backbone {t_b:.2f} ms, head {t_h:.2f} ms "
          f"({head_flops / t_h / 1e9:.0f} TFLOP/s), detector {t_d:.2f} ms = {B / t_d * 1e3:.0f} images/s")
    for k, v in prof.items():
        if v["launches"]:
            print(f"  head {k:10s} {v['ms']:.3f} ms in {v['launches']} launches")


if __name__ == "__main__":
    main()
