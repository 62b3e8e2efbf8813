"""Loads the reference's two scripts as modules: from /root/reference in the build container, from
the byte-compiled copies oracle/build_ref.py leaves in oracle/_ref/ on the GPU box (where
/root/reference does not exist).  Third-party imports the scripts make at module top but only use for data
loading / visualisation (albumentations, pycocotools, matplotlib - absent here) are stubbed in
sys.modules.  Test/fixture infrastructure only."""
from __future__ import annotations

import importlib.machinery
import importlib.util
import sys
import types
from pathlib import Path

REFERENCE_DIR = Path("/root/reference")
COMPILED_DIR = Path(__file__).resolve().parent / "_ref"


def _compiled_ok() -> bool:
    ver = COMPILED_DIR / "PYTHON_VERSION"
    return (COMPILED_DIR / "evaluation.pyc.bin").exists() and ver.exists() and \
        ver.read_text().strip() == "%d.%d" % sys.version_info[:2]


def reference_available() -> bool:
    return (REFERENCE_DIR / "evaluation.py").exists() or _compiled_ok()


def reference_origin() -> str:
    if (REFERENCE_DIR / "evaluation.py").exists():
        return "source (/root/reference)"
    return "byte-compiled (oracle/_ref)" if _compiled_ok() else "absent"


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _install_stubs():
    try:
        import albumentations  # noqa: F401
    except Exception:
        _stub("albumentations")
        _stub("albumentations.pytorch", ToTensorV2=object)
    try:
        import pycocotools  # noqa: F401
    except Exception:
        _stub("pycocotools")
        _stub("pycocotools.coco", COCO=object)
        _stub("pycocotools.cocoeval", COCOeval=object)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub("matplotlib")
        _stub("matplotlib.pyplot")
        _stub("matplotlib.patches")
    try:
        import wandb  # noqa: F401
    except Exception:
        _stub("wandb")


def load(script: str = "evaluation"):
    """Returns the reference script as a module. `evaluation` has no import side effects;
    `train` sets the fork start method and TF32 matmul precision (train.py:17-19) - the latter is
    reset to 'highest' here."""
    name = f"_reference_{script}"
    if name in sys.modules:
        return sys.modules[name]
    _install_stubs()
    if (REFERENCE_DIR / f"{script}.py").exists():
        spec = importlib.util.spec_from_file_location(name, REFERENCE_DIR / f"{script}.py")
    elif _compiled_ok():
        path = str(COMPILED_DIR / f"{script}.pyc.bin")
        spec = importlib.util.spec_from_loader(
            name, importlib.machinery.SourcelessFileLoader(name, path), origin=path)
    else:
        raise FileNotFoundError("the reference is neither at /root/reference nor in oracle/_ref "
                                "(python oracle/build_ref.py in the build container)")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    # train.py:17 calls mp.set_start_method('fork') at import, which raises when a start method is
    # already fixed (pytest-xdist, torch.distributed tests); neutralise it for the import only.
    import multiprocessing as _mp
    orig = _mp.set_start_method
    _mp.set_start_method = lambda *a, **k: None
    try:
        spec.loader.exec_module(mod)
    except BaseException:
        sys.modules.pop(name, None)
        raise
    finally:
        _mp.set_start_method = orig
    if script == "train":
        import torch
        torch.set_float32_matmul_precision("highest")
    return mod
