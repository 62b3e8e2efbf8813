"""Recipe for oracle/_ref: the reference's own two scripts, byte-compiled where they lie.

    python oracle/build_ref.py

The reference is Python (no native code, no setup.py / pyproject.toml, so nothing to pip-install
into baseline/_ref): its "build" is CPython's byte-compiler run on /root/reference/evaluation.py and
/root/reference/train.py, with the outputs written to oracle/_ref/ only (as *.pyc.bin: the gpurun
snapshot leaves *.pyc files behind).  That directory is
git-ignored (no reference source or binary enters the history) but not gpurun-ignored, so the
compiled modules travel to the GPU box, where /root/reference does not exist: there
`oracle.ref_loader` imports them with a sourceless loader and bench.py's reference arms
(`--impl reference`, `cpu_baseline`, `gpu_eager_baseline`) time the reference's OWN classes
(kind "reference") instead of the oracle's restatement (kind "port").
Test / measurement infrastructure only; the product never imports it.
"""
from __future__ import annotations

import py_compile
import sys
from pathlib import Path

REFERENCE_DIR = Path("/root/reference")
OUT_DIR = Path(__file__).resolve().parent / "_ref"
SCRIPTS = ("evaluation", "train")


def build() -> list[Path]:
    if not REFERENCE_DIR.exists():
        return []
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    out = []
    for name in SCRIPTS:
        src = REFERENCE_DIR / f"{name}.py"
        dst = OUT_DIR / f"{name}.pyc.bin"
        if not dst.exists() or dst.stat().st_mtime < src.stat().st_mtime:
            py_compile.compile(str(src), cfile=str(dst), dfile=f"reference/{name}.py", doraise=True)
        out.append(dst)
    (OUT_DIR / "PYTHON_VERSION").write_text("%d.%d" % sys.version_info[:2])
    return out


if __name__ == "__main__":
    for p in build():
        print(p)
