"""Builds libvitk.so (hand-written sm_100a CUDA + the C ABI of include/vitk.h) in-tree.

    python automated-recycling-sorter-with-vision-transformers_b200/build.py [--force]

nvcc cross-compiles for sm_100a without a GPU.  Objects are cached by source mtime; the shared
library lands next to this file so that it travels with a gpurun snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OBJ_DIR = CSRC / "obj"
LIB_PATH = PKG_DIR / "libvitk.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
    *os.environ.get("VITK_NVCC_EXTRA", "").split(),   # e.g. -DVITK_ATTN_TRACE (debug builds)
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _headers_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG_DIR.parent / "include" / "vitk.h"]
    return max(h.stat().st_mtime for h in hdrs if h.exists())


def _compile(src: Path, force: bool) -> tuple[Path, str]:
    obj = OBJ_DIR / (src.stem + ".o")
    newest = max(src.stat().st_mtime, _headers_mtime())
    if not force and obj.exists() and obj.stat().st_mtime >= newest:
        return obj, ""
    cmd = [NVCC, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ_DIR.mkdir(parents=True, exist_ok=True)
    srcs = sources()
    if not srcs:
        raise RuntimeError(f"no CUDA sources under {CSRC}")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, force), srcs))
    objs = [o for o, _ in results]
    log = "".join(l for _, l in results)
    if verbose and log:
        print(log)
    (OBJ_DIR / "ptxas.log").write_text(log) if log else None
    newest_obj = max(o.stat().st_mtime for o in objs)
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < newest_obj:
        cmd = [NVCC, "-shared", "-o", str(LIB_PATH), *map(str, objs), "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
