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
