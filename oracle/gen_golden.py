"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CLASSES (build container only).

    python oracle/gen_golden.py

Imports /root/reference/evaluation.py (VisionTransformer) and /root/reference/train.py
(DataEfficientImageTransformer) through oracle/ref_loader.py, builds them under fixed seeds,
feeds seeded synthetic images and stores inputs, weights (small configs) or weight checksums
(ViT-B/16) and the reference outputs in fp32 and fp64.  The fixtures pin the oracle
(tests/test_oracle.py) and, through it, the CUDA path (tests/test_golden_gpu.py).

Fixture kinds
  *_full.npz   : tiny configs - every parameter is stored, nothing depends on RNG reproduction
  vitb16_*.npz : ViT-B/16 - only seeds, a SHA-256 of the state_dict and the outputs are stored;
                 the tests rebuild the weights from the seed (our modules reproduce the
                 reference's construction order) and verify the checksum before comparing
  det_head_*   : the detection head (evaluation.py:160-200) on seeded encoder tokens; parameters are
                 a pure function of a seed (vit_oracle.randomize_head_state), outputs fp32 / fp64
  trainstep_*  : one fine-tune step (cross-entropy on the 6-class CLS head, AdamW of
                 train.py:1598-1602, dropout 0): loss, per-parameter gradient and updated weights
"""
from __future__ import annotations

import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref_loader  # noqa: E402
from oracle import vit_oracle as O  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"

TINY = dict(image_size=32, patch_size=16, in_channels=3, embed_dim=64, num_layers=2, num_heads=1,
            mlp_dim=128, dropout=0.0)
SMALL = dict(image_size=64, patch_size=16, in_channels=3, embed_dim=128, num_layers=3, num_heads=2,
             mlp_dim=512, dropout=0.0)
TRAIN_SMALL = dict(image_size=32, patch_size=16, in_channels=3, embed_dim=128, num_layers=2,
                   num_heads=2, mlp_dim=256, dropout=0.0)
VITB = dict(image_size=224, patch_size=16, in_channels=3, embed_dim=768, num_layers=12,
            num_heads=12, mlp_dim=3072, dropout=0.0)


def state_sha256(sd: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_reference(kind: str, kw: dict, seed: int, n_classes: int = 6):
    """Reference backbone + the north-star head, constructed exactly as SURVEY.md 8c prescribes:
    torch.manual_seed(seed); reference class; nn.Linear(D, C) immediately after."""
    torch.manual_seed(seed)
    if kind == "vit":
        backbone = ref_loader.load("evaluation").VisionTransformer(**kw)
    else:
        backbone = ref_loader.load("train").DataEfficientImageTransformer(**kw)
    head = torch.nn.Linear(kw["embed_dim"], n_classes)
    return backbone, head


def run_reference(backbone, head, x, dtype):
    backbone = backbone.to(dtype).eval()
    head = head.to(dtype)
    with torch.no_grad():
        tokens = backbone(x.to(dtype))          # evaluation.py:231 / train.py:831
        logits = head(tokens[:, 0])
    return tokens, logits


def full_state(backbone, head) -> dict:
    sd = {"backbone." + k: v.detach().clone() for k, v in backbone.state_dict().items()}
    sd["head.weight"] = head.weight.detach().clone()
    sd["head.bias"] = head.bias.detach().clone()
    return sd


def gen_forward(name, kind, kw, batch, seed, store_weights, image_seed=1234):
    backbone, head = build_reference(kind, kw, seed)
    sd = full_state(backbone, head)
    x = O.synthetic_images(batch, kw["image_size"], seed=image_seed)
    t32, l32 = run_reference(backbone, head, x, torch.float32)
    import copy
    t64, l64 = run_reference(copy.deepcopy(backbone), copy.deepcopy(head), x, torch.float64)
    out = dict(kind=kind, seed=seed, image_seed=image_seed, batch=batch,
               config=np.array(sorted(kw.items()), dtype=object), state_sha256=state_sha256(sd),
               logits_f32=l32.numpy(), logits_f64=l64.numpy())
    if store_weights:
        out["images"] = x.numpy()
        out["tokens_f32"] = t32.numpy()
        out["tokens_f64"] = t64.numpy()
        for k, v in sd.items():
            out["w:" + k] = v.numpy()
    else:
        out["tokens_f64_head"] = t64[:, :4, :32].numpy()     # a slice is enough to pin the tokens
        out["tokens_f64_rowsum"] = t64.sum(dim=-1).numpy()
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    print(f"{name}: logits32-64 max diff {float((l32.double() - l64).abs().max()):.2e}, "
          f"sha {out['state_sha256'][:12]}")


def gen_trainstep(name, kind, kw, batch, seed):
    """train.py:1441-1460 with cross-entropy on the CLS head in place of the detection loss."""
    backbone, head = build_reference(kind, kw, seed)
    sd0 = full_state(backbone, head)
    x = O.synthetic_images(batch, kw["image_size"], seed=4321)
    y = O.synthetic_labels(batch, 6, seed=1)
    backbone = backbone.double().train()   # dropout p = 0.0 -> deterministic
    head = head.double()
    params = list(backbone.parameters()) + list(head.parameters())
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)   # train.py:1598-1602
    opt.zero_grad()
    logits = head(backbone(x.double())[:, 0])
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    grads = {"backbone." + k: p.grad.detach().clone() for k, p in backbone.named_parameters()}
    grads["head.weight"] = head.weight.grad.detach().clone()
    grads["head.bias"] = head.bias.grad.detach().clone()
    opt.step()
    sd1 = full_state(backbone, head)
    out = dict(kind=kind, seed=seed, batch=batch, config=np.array(sorted(kw.items()), dtype=object),
               state_sha256=state_sha256(sd0), loss_f64=float(loss), labels=y.numpy(),
               images=x.numpy(), logits_f64=logits.detach().numpy())
    for k in sd0:   # fp32 storage keeps the fixtures small; fp64 math produced them
        out["w:" + k] = sd0[k].float().numpy()
        out["g:" + k] = grads[k].float().numpy()
        out["n:" + k] = sd1[k].float().numpy()
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    print(f"{name}: loss {float(loss):.6f}")


def gen_head(name, embed_dim, num_queries, n_tokens, batch, seed, token_seed, store_tokens):
    """ObjectDetectionHead of evaluation.py:160-200 run as ViTObjectDetector.forward runs it
    (evaluation.py:231-238): memory = tokens[:, 1:, :].  Parameters come from
    O.randomize_head_state (a pure function of the seed and the reference's own state_dict keys /
    shapes), so the tests rebuild them without the reference; the SHA-256 guards that."""
    ev = ref_loader.load("evaluation")
    torch.manual_seed(seed)
    head = ev.ObjectDetectionHead(embed_dim=embed_dim, num_classes=6, num_queries=num_queries).eval()
    sd = O.randomize_head_state(head.state_dict(), seed)
    head.load_state_dict(sd)
    tokens = torch.randn(batch, n_tokens, embed_dim, generator=torch.Generator().manual_seed(token_seed))
    import copy
    with torch.no_grad():
        o32 = head(tokens[:, 1:, :])
        o64 = copy.deepcopy(head).double()(tokens[:, 1:, :].double())
    out = dict(seed=seed, token_seed=token_seed, batch=batch, n_tokens=n_tokens, embed_dim=embed_dim,
               num_queries=num_queries, num_classes=6, state_sha256=state_sha256(sd),
               state_keys=np.array(sorted(sd)),
               class_logits_f32=o32["class_logits"].numpy(), bbox_f32=o32["bbox_coords"].numpy(),
               class_logits_f64=o64["class_logits"].numpy(), bbox_f64=o64["bbox_coords"].numpy())
    if store_tokens:
        out["tokens"] = tokens.numpy()
    np.savez_compressed(GOLDEN / f"{name}.npz", **out)
    d = float((o32["class_logits"].double() - o64["class_logits"]).abs().max())
    print(f"{name}: logits32-64 max diff {d:.2e}, sha {out['state_sha256'][:12]}")


def main():
    if not ref_loader.reference_available():
        raise SystemExit("/root/reference is not present - golden vectors can only be generated "
                         "in the build container")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    gen_forward("tiny_vit_full", "vit", TINY, 3, seed=0, store_weights=True)
    gen_forward("tiny_deit_full", "deit", TINY, 2, seed=1, store_weights=True)
    gen_forward("small_deit_full", "deit", SMALL, 5, seed=2, store_weights=True)
    gen_forward("vitb16_vit", "vit", VITB, 8, seed=0, store_weights=False)
    gen_forward("vitb16_deit", "deit", VITB, 8, seed=3, store_weights=False)
    gen_trainstep("trainstep_tiny_vit", "vit", TINY, 4, seed=5)
    gen_trainstep("trainstep_small_deit", "deit", TRAIN_SMALL, 6, seed=6)
    gen_heads()


def gen_heads():
    gen_head("det_head_small", 256, 10, 17, 3, seed=11, token_seed=12, store_tokens=True)
    gen_head("det_head_vitb", 768, 100, 197, 2, seed=13, token_seed=14, store_tokens=False)


if __name__ == "__main__":
    if "--heads-only" in sys.argv:      # the detection-head fixtures were added after the others
        GOLDEN.mkdir(parents=True, exist_ok=True)
        gen_heads()
    else:
        main()
