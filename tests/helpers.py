"""Shared fixture loading for the golden-vector tests."""
import hashlib
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=True)
    cfg = {k: (v if not isinstance(v, np.generic) else v.item()) for k, v in z["config"]}
    cfg = {k: (float(v) if k == "dropout" else int(v)) for k, v in cfg.items()}
    return z, cfg


def weights(z, prefix="w:"):
    return {k[len(prefix):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(prefix)}


def state_sha256(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def build_classifier(vitk, z, cfg, seed_rebuild=False):
    """ViTClassifier carrying the fixture's weights (stored, or rebuilt from the seed and checked
    against the stored SHA-256)."""
    kind = str(z["kind"])
    kw = {k: v for k, v in cfg.items()}
    if seed_rebuild:
        torch.manual_seed(int(z["seed"]))
        model = vitk.ViTClassifier(num_classes=6, deit=(kind == "deit"), **kw)
        got = state_sha256(model.state_dict())
        assert got == str(z["state_sha256"]), "seeded weights differ from the reference's"
    else:
        model = vitk.ViTClassifier(num_classes=6, deit=(kind == "deit"), **kw)
        missing = model.load_state_dict(weights(z), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
    return model


def build_head(vitk, z):
    """ObjectDetectionHead mirror carrying a det_head_* fixture's parameters: rebuilt from the seed
    with oracle.randomize_head_state over the mirror's own state_dict keys / shapes (which must be
    the reference's) and checked against the stored SHA-256."""
    from oracle import vit_oracle as O
    head = vitk.ObjectDetectionHead(embed_dim=int(z["embed_dim"]), num_classes=int(z["num_classes"]),
                                    num_queries=int(z["num_queries"]))
    assert sorted(head.state_dict()) == [str(k) for k in z["state_keys"]]
    sd = O.randomize_head_state(head.state_dict(), int(z["seed"]))
    assert state_sha256(sd) == str(z["state_sha256"]), "rebuilt head weights differ from the fixture"
    head.load_state_dict(sd)
    return head.eval(), sd


def head_tokens(z):
    if "tokens" in z.files:
        return torch.from_numpy(z["tokens"])
    return torch.randn(int(z["batch"]), int(z["n_tokens"]), int(z["embed_dim"]),
                       generator=torch.Generator().manual_seed(int(z["token_seed"])))
