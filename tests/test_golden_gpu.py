"""CUDA path (through vitk_forward) against the committed golden vectors that the REFERENCE's own
classes produced (oracle/gen_golden.py).  Tolerances are north_star's bf16 bar."""
import pytest
import torch

from oracle import vit_oracle as O
from tests import helpers as H
from tests.test_model_gpu import assert_top1

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["tiny_vit_full", "tiny_deit_full", "small_deit_full"])
def test_full_fixture(vitk, name):
    z, cfg = H.load(name)
    model = H.build_classifier(vitk, z, cfg).cuda().eval()
    x = torch.from_numpy(z["images"]).cuda()
    with torch.no_grad():
        tokens = model.backbone(x).cpu().double()
        logits = model(x).cpu().double()
    l_ref = torch.from_numpy(z["logits_f64"])
    assert (logits - l_ref).abs().max() < 2e-2
    assert (tokens - torch.from_numpy(z["tokens_f64"])).abs().max() < 6e-2
    assert_top1(logits, l_ref)


@pytest.mark.parametrize("name", ["vitb16_vit", "vitb16_deit"])
def test_vitb16_fixture(vitk, name):
    z, cfg = H.load(name)
    model = H.build_classifier(vitk, z, cfg, seed_rebuild=True).cuda().eval()
    x = O.synthetic_images(int(z["batch"]), cfg["image_size"], seed=int(z["image_seed"])).cuda()
    with torch.no_grad():
        tokens = model.backbone(x).cpu().double()
        logits = model(x).cpu().double()
    l_ref = torch.from_numpy(z["logits_f64"])
    err = (logits - l_ref).abs().max().item()
    print(name, "max |logit - reference fp64| =", err)
    assert err < 2e-2
    assert (tokens[:, :4, :32] - torch.from_numpy(z["tokens_f64_head"])).abs().max() < 1e-1
    assert (tokens.sum(-1) - torch.from_numpy(z["tokens_f64_rowsum"])).abs().max() < 2.0
    assert_top1(logits, l_ref)
