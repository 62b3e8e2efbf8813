"""fp32-parity mode: logits within 1e-4 of the fp32/fp64 reference (north_star's fp32 bar)."""
import pytest
import torch

from oracle import vit_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_split_gemm_is_fp32_accurate(vitk):
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(300, 768, generator=g, device="cuda")
    w = torch.randn(512, 768, generator=g, device="cuda") * 0.03
    out = vitk.ops.gemm(vitk.ops.split3(a, False), vitk.ops.split3(w, True), vitk._lib.EPI_F32)
    ref = (a.double() @ w.double().t())
    # fp32 accumulation in the tensor core (K = 768 terms): ~4e-5; plain bf16 operands give ~3e-2
    assert (out.double() - ref).abs().max() < 1e-4
    bf = vitk.ops.gemm(a.bfloat16(), w.bfloat16(), vitk._lib.EPI_F32)
    assert (bf.double() - ref).abs().max() > 1e-3


@pytest.mark.parametrize("name", ["tiny_vit_full", "small_deit_full"])
def test_fp32_mode_small_fixtures(vitk, name):
    z, cfg = H.load(name)
    model = H.build_classifier(vitk, z, cfg).cuda().eval()
    model.backbone.set_precision("fp32")
    x = torch.from_numpy(z["images"]).cuda()
    with torch.no_grad():
        tokens = model.backbone(x).cpu().double()
        logits = model(x).cpu().double()
    assert (logits - torch.from_numpy(z["logits_f64"])).abs().max() < 1e-4
    assert (tokens - torch.from_numpy(z["tokens_f64"])).abs().max() < 5e-4


def test_fp32_mode_vit_b16(vitk):
    z, cfg = H.load("vitb16_vit")
    model = H.build_classifier(vitk, z, cfg, seed_rebuild=True).cuda().eval()
    model.backbone.set_precision("fp32")
    x = O.synthetic_images(int(z["batch"]), 224, seed=int(z["image_seed"])).cuda()
    with torch.no_grad():
        logits = model(x).cpu().double()
    l_ref = torch.from_numpy(z["logits_f64"])
    err = (logits - l_ref).abs().max().item()
    print("ViT-B/16 fp32-mode max |logit - reference fp64| =", err)
    assert err < 1e-4
    assert torch.equal(logits.argmax(-1), l_ref.argmax(-1))
    # and back to the fast mode on the same module
    model.backbone.set_precision("bf16")
    with torch.no_grad():
        assert (model(x).cpu().double() - l_ref).abs().max() < 2e-2
