"""CPU tier: the C-ABI shared library loads, exports every symbol include/vitk.h declares, and its
argument checking / error reporting work without a GPU (no compute calls here)."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "vitk.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vitk_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(vitk):
    lib = vitk._lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 10
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vitk.h but not exported"
    assert sorted(vitk._lib.exported_symbols()) == declared, "ctypes binding out of sync with header"
    assert lib.vitk_abi_version() == vitk._lib.ABI_VERSION


def test_workspace_query_and_config_validation(vitk):
    lib = vitk._lib.lib()
    cfg = vitk._lib.VitkConfig(image_size=224, patch_size=16, in_channels=3, embed_dim=768,
                               num_layers=12, num_heads=12, mlp_dim=3072, n_prefix_tokens=1,
                               n_classes=6, precision=0, ln_eps=1e-5, dropout_p=0.0, seed=0)
    need = C.c_size_t(0)
    assert lib.vitk_workspace_bytes(C.byref(cfg), 256, C.byref(need)) == 0
    M = 256 * 197
    expect = M * 768 * 4 + M * 768 * 2 + M * 2304 * 2 + M * 768 * 2 + M * 3072 * 2 + 256 * 196 * 768 * 2
    expect += 256 * (768 * 4 + 768 * 2 + 3072 * 2)     # CLS-only tail of vitk_forward_cls
    expect += lib.vitk_stats_parts(768) * M * 8         # row statistics of the folded LayerNorms
    assert lib.vitk_stats_parts(768) == 6 and lib.vitk_stats_parts(1024) == 8
    assert expect <= need.value <= expect + 16 * 1024
    cfg.image_size = 225   # not a multiple of the patch size
    rc = lib.vitk_workspace_bytes(C.byref(cfg), 256, C.byref(need))
    assert rc == 1 and b"image_size" in lib.vitk_last_error()
    cfg.image_size = 224
    cfg.n_prefix_tokens = 3
    assert lib.vitk_workspace_bytes(C.byref(cfg), 1, C.byref(need)) == 1
    cfg.n_prefix_tokens = 1
    assert lib.vitk_workspace_bytes(C.byref(cfg), 0, C.byref(need)) == 1   # empty batch


def test_detection_head_workspace_and_validation(vitk):
    lib = vitk._lib.lib()
    cfg = vitk._lib.VitkDetectionHeadConfig(embed_dim=768, num_heads=8, ffn_dim=2048, num_layers=6,
                                            num_queries=100, num_outputs=7, ln_eps=1e-5)
    need = C.c_size_t(0)
    assert lib.vitk_detection_head_workspace_bytes(C.byref(cfg), 256, 197, 1, C.byref(need)) == 0
    M, Mm = 256 * 100, 256 * 197
    expect = (M * 768 * (4 + 2 + 6 + 2) + M * 2048 * 2 + Mm * 768 * 2 + Mm * 6 * 2 * 768 * 2
              + 100 * 768 * 6)
    assert expect <= need.value <= expect + 16 * 1024
    assert lib.vitk_detection_head_workspace_bytes(C.byref(cfg), 256, 197, 197, C.byref(need)) == 1
    assert b"detection head" in lib.vitk_last_error()          # no memory tokens left
    cfg.num_heads = 96                                          # head_dim 8
    assert lib.vitk_detection_head_workspace_bytes(C.byref(cfg), 1, 197, 1, C.byref(need)) == 1
    assert b"head_dim" in lib.vitk_last_error()
    cfg.num_heads = 8
    assert lib.vitk_detection_head_forward(C.byref(cfg), None, None, 1, 197, 1, None, None, None, 0,
                                           None) != 0           # null weights: reported, no crash


def test_null_arguments_are_reported_not_crashed(vitk):
    lib = vitk._lib.lib()
    assert lib.vitk_gemm(None, 0, None, 0, 0, 0, 0, 0, None, None, 0, None, None, None, 0,
                         1.0, 0.0, None) != 0
    assert b"gemm" in lib.vitk_last_error()
    assert lib.vitk_layernorm(None, 0, None, None, None, 0, 0, None, None, 0, 0, 1e-5, None) != 0
    assert lib.vitk_attention(None, None, None, 1, 1, 1, 64, None) != 0
    assert lib.vitk_patchify(None, None, 1, 3, 224, 16, None) != 0


def test_no_cpu_fallback(vitk):
    """CPU tensors are refused by every layer of the product path."""
    with pytest.raises(vitk.VitkError):
        vitk.ops.gemm(torch.zeros(8, 8).bfloat16(), torch.zeros(8, 8).bfloat16())
    m = vitk.ViTClassifier(image_size=32, embed_dim=64, num_layers=1, num_heads=1, mlp_dim=64)
    with pytest.raises(vitk.VitkError), torch.no_grad():
        m(torch.zeros(1, 3, 32, 32))


def test_product_never_imports_the_oracle():
    pkg = ROOT / "automated-recycling-sorter-with-vision-transformers_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        src = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, flags=re.M), \
            f"{f} imports the oracle"
        assert "#include \"../oracle" not in src and "oracle/" not in src, f"{f} uses the oracle"


def test_ctypes_structs_match_the_header(vitk, tmp_path):
    """Every struct of include/vitk.h, compiled by gcc, has the size and the field offsets of its
    ctypes mirror in _lib.py (a drifted mirror would hand the library shifted pointers)."""
    import shutil
    import subprocess
    L = vitk._lib
    mirrors = {name: getattr(L, name) for name in (
        "VitkConfig", "VitkBlockWeights", "VitkWeights", "VitkBlockWeightsT", "VitkWeightsT",
        "VitkBlockGrads", "VitkGrads", "VitkDetectionHeadConfig", "VitkDecoderLayerWeights",
        "VitkDetectionHeadWeights", "VitkDecoderLayerWeightsT", "VitkDetectionHeadWeightsT",
        "VitkDecoderLayerGrads", "VitkDetectionHeadGrads", "VitkPeerBuffers")}
    header = (ROOT / "include" / "vitk.h").read_text()
    for name in mirrors:
        assert re.search(r"typedef struct %s\b" % name, header), f"{name} is not in include/vitk.h"
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "vitk.h"', "int main(void) {"]
    for name, cls in mirrors.items():
        lines.append(f'  printf("{name} size %zu\\n", sizeof({name}));')
        for field, _ in cls._fields_:
            lines.append(f'  printf("{name} {field} %zu\\n", offsetof({name}, {field}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.splitlines():
        name, field, value = line.split()
        cls = mirrors[name]
        if field == "size":
            assert C.sizeof(cls) == int(value), (name, C.sizeof(cls), value)
        else:
            assert getattr(cls, field).offset == int(value), (name, field)
