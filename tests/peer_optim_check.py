"""Not a pytest file (needs >= 2 GPUs):
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/peer_optim_check.py
The peer-memory optimizer step (csrc/peer_optim.cu) against the NCCL all-reduce + replicated AdamW
on the same model, inputs and steps: parameters agree, replicas bitwise identical, the skip-on-
non-finite flag reaches every rank; then the ViT-B/16 batch-128 step timed with both transports."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dev = torch.device("cuda", torch.cuda.current_device())
dist.init_process_group("nccl", device_id=dev)
say = (lambda *a: print(*a, flush=True)) if rank == 0 else (lambda *a: None)

KW = dict(image_size=64, patch_size=16, embed_dim=256, num_layers=2, num_heads=4, mlp_dim=512)


def upd_err(t, ref, init):
    """Relative L2 error of the parameter UPDATE (Adam turns a last-bit difference of a near-zero
    gradient into a difference of the order of lr in that one element; the update as a whole must
    agree)."""
    du = (t.state.flat - ref.state.flat).double().norm()
    return float(du / ((ref.state.flat - init).double().norm() + 1e-30))


def run(sync, steps=4, poison_at=None, multicast=True):
    torch.manual_seed(0)
    model = vitk.ViTClassifier(num_classes=6, dropout=0.0, **KW).to(dev).train()
    global INIT
    tuner = vitk.FineTuner(model, lr=1e-3, weight_decay=1e-2, grad_sync=sync)
    INIT = tuner.state.flat.clone()
    if sync == "peer" and not multicast:
        tuner._peer_buffers = tuner.peer.buffers(True, multicast=False)
    x = O.synthetic_images(8, 64, seed=10 + rank).to(dev)
    y = O.synthetic_labels(8, 6, seed=20 + rank).to(dev)
    for k in range(steps):
        xx = x
        if poison_at == k and rank == world - 1:
            xx = x.clone()
            xx[0, 0, 0, 0] = float("inf")        # one rank's gradients go non-finite
        tuner.step(xx, y)
    torch.cuda.synchronize()
    return tuner


a = run("nccl")
b = run("peer")
say("transport:", b.grad_sync, "multicast:", b.peer.multicast, "buckets:", len(b.peer.buckets), "overlap:", b.overlap)
d = (a.state.flat - b.state.flat).abs().max().item()
ds = (a.state.shadow.float() - b.state.shadow.float()).abs().max().item()
lo, hi = b.state.flat.clone(), b.state.flat.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
say(f"peer vs nccl after 4 steps: max |d param| {d:.3e}, max |d shadow| {ds:.3e}, relative error of the "
    f"update {upd_err(b, a, INIT):.3e}, replicas identical: {bool(torch.equal(lo, hi))}")
assert upd_err(b, a, INIT) < 1e-2 and torch.equal(lo, hi)
c = run("peer", multicast=False)
d2 = (b.state.flat - c.state.flat).abs().max().item()
say(f"multicast vs peer loads/stores: max |d param| {d2:.3e}, relative error of the update "
    f"{upd_err(c, b, INIT):.3e}")
assert upd_err(c, b, INIT) < 1e-2
# skip-on-non-finite: poisoned on the LAST rank only; every rank must skip that step
p0 = run("peer", steps=3, poison_at=1)
n0 = run("nccl", steps=3, poison_at=1)
say("skipped steps (peer, nccl):", p0.skipped_steps, n0.skipped_steps,
    "relative error of the update", upd_err(p0, n0, INIT))
assert p0.skipped_steps == 1 and n0.skipped_steps == 1
assert upd_err(p0, n0, INIT) < 1e-2
# checkpoint: moments gathered from their owners
sd_p, sd_n = p0.optimizer_state_dict(), n0.optimizer_state_dict()
worst = max((sd_p["state"][i]["exp_avg"] - sd_n["state"][i]["exp_avg"]).abs().max().item()
            for i in sd_n["state"])
say("optimizer_state_dict: max |d exp_avg| vs nccl", worst)
assert worst < 1e-4
del a, b, c, p0, n0
torch.cuda.empty_cache()

if os.environ.get("VITK_PEER_CHECK_QUICK") == "1":   # the pytest wrapper: correctness only
    dist.destroy_process_group()
    sys.exit(0)

# ---- timing at the benchmark's geometry
VITB = dict(image_size=224, patch_size=16, embed_dim=768, num_layers=12, num_heads=12, mlp_dim=3072)
for sync in ("nccl", "peer", "nccl", "peer"):
    torch.manual_seed(0)
    model = vitk.ViTClassifier(num_classes=6, dropout=0.1, **VITB).to(dev).train()
    tuner = vitk.FineTuner(model, lr=1e-4, weight_decay=1e-4, grad_sync=sync, seed=1000 * rank)
    x = O.synthetic_images(128, 224, seed=99 + rank).to(dev)
    y = O.synthetic_labels(128, 6, seed=5 + rank).to(dev)
    for _ in range(3):
        tuner.step(x, y)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        tuner.step(x, y)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 8], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    vitk._lib.profile_enable(True)
    for _ in range(2):
        tuner.step(x, y)
    torch.cuda.synchronize()
    opt_ms = vitk._lib.profile_collect()["optimizer"]["ms"] / 2
    vitk._lib.profile_enable(False)
    say(f"ViT-B/16 batch 128 per GPU x {world} GPUs, grad_sync={tuner.grad_sync}: {t.item():.3f} ms per step "
        f"= {world * 128 / t.item() * 1e3:.0f} images/s (optimizer-class kernels {opt_ms:.3f} ms per step)")
    del tuner, model
    torch.cuda.empty_cache()
dist.destroy_process_group()
