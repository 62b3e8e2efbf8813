"""Not a pytest file: where the end-to-end loop (HostBatchRunner) loses time against the
device-timed loop - wall clock per step, device time of each forward inside the runner, and the
same loop with the copies switched off one at a time."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

torch.manual_seed(0)
B, steps = 256, 20
model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=224, patch_size=16, embed_dim=768,
                           num_layers=12, num_heads=12, mlp_dim=3072).cuda().eval()
x_host = O.synthetic_images(B, 224).pin_memory()
x_dev = x_host.cuda()


def wall(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


with torch.no_grad():
    for _ in range(3):
        model(x_dev)

    def dev_loop():
        for _ in range(steps):
            model(x_dev)
    print(f"device loop, wall clock:            {wall(dev_loop):7.3f} ms/step")

    host = torch.empty(B, 6).pin_memory()

    def dev_loop_d2h():
        for _ in range(steps):
            host.copy_(model(x_dev), non_blocking=True)
    print(f"device loop + D2H of the logits:    {wall(dev_loop_d2h):7.3f} ms/step")

    runner = vitk.HostBatchRunner(model, B, "cuda")
    for _ in runner.run([x_host] * 3):
        pass

    def run_loop():
        for _ in runner.run([x_host] * steps):
            pass
    print(f"HostBatchRunner f32:                {wall(run_loop):7.3f} ms/step")

    # the same with the forwards timed on the device inside the runner
    ev = []
    orig = model.forward

    def timed_forward(x):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y = orig(x)
        e1.record()
        ev.append((e0, e1))
        return y
    model.forward = timed_forward
    w = wall(run_loop)
    model.forward = orig
    fw = [a.elapsed_time(b) for a, b in ev]
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(len(ev) - 1)]
    print(f"  with events: {w:7.3f} ms/step; forward on the device {sum(fw) / len(fw):7.3f} ms "
          f"(min {min(fw):.3f}, max {max(fw):.3f}); gap between forwards {sum(gaps) / len(gaps):7.3f} ms "
          f"(max {max(gaps):.3f})")
