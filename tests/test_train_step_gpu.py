"""One fused fine-tune step (forward -> CE -> backward -> AdamW) through the C ABI against the
golden vectors produced by the reference's own classes in fp64 (tests/golden/trainstep_*.npz)
and against the oracle on other seeded configs."""
import pytest
import torch

from oracle import vit_oracle as O
from tests import helpers as H

pytestmark = pytest.mark.gpu

# Relative L2 error allowed per parameter gradient (bf16 operands, fp32 accumulation), set at about
# twice the worst value observed on B200 (printed by the tests): 0.0094 over the golden / oracle
# small configurations, 0.0050 for ViT-B/16 at full width (with and without dropout).
GRAD_REL_TOL = 0.02
GRAD_REL_TOL_FULL_WIDTH = 0.012


def _check_step(vitk, model, x, y, loss_ref, g_ref, n_ref, lr=1e-4):
    model = model.cuda()
    tuner = vitk.FineTuner(model, lr=lr, weight_decay=1e-4)
    loss, logits = tuner.step(x.cuda(), y.cuda())
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref) < 2e-2, (loss.item(), loss_ref)
    grads = dict(zip(tuner.state.names, [p.grad for p in tuner.state.params]))
    worst = 0.0
    for k, gr in g_ref.items():
        got = grads[k].detach().cpu().double().reshape(-1)
        want = gr.double().reshape(-1)
        rel = (got - want).norm() / (want.norm() + 1e-12)
        cos = torch.nn.functional.cosine_similarity(got, want, dim=0)
        worst = max(worst, rel.item())
        assert rel < GRAD_REL_TOL and cos > 0.999, (k, rel.item(), cos.item())
    print("worst relative gradient error:", worst)
    sd = model.state_dict()
    for k, want in n_ref.items():
        err = (sd[k].cpu().double() - want.double()).abs()
        # first AdamW step moves every weight by ~lr * sign(g): a sign flip of a near-zero gradient
        # costs 2 * lr, everything else must agree closely
        assert err.max() < 2.2 * lr, (k, err.max().item())
        g = g_ref[k].double().abs()
        solid = g > 1e-2 * g.max()          # e.g. the key bias has an exactly-zero gradient
        if solid.any():
            assert err[solid].mean() < 0.1 * lr, (k, err[solid].mean().item())
    return tuner


@pytest.mark.parametrize("name", ["trainstep_tiny_vit", "trainstep_small_deit"])
def test_train_step_matches_reference_golden(vitk, name):
    z, cfg = H.load(name)
    model = H.build_classifier(vitk, z, cfg)
    x, y = torch.from_numpy(z["images"]), torch.from_numpy(z["labels"])
    _check_step(vitk, model, x, y, float(z["loss_f64"]), H.weights(z, "g:"), H.weights(z, "n:"))


def test_train_step_matches_oracle_multi_tile(vitk):
    """More tokens than one 256-row GEMM tile and a ragged K split in the weight gradients."""
    kw = dict(image_size=96, patch_size=16, embed_dim=128, num_layers=2, num_heads=2, mlp_dim=256,
              dropout=0.0)
    torch.manual_seed(21)
    model = vitk.ViTClassifier(num_classes=6, **kw)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x, y = O.synthetic_images(9, 96, seed=7), O.synthetic_labels(9, 6, seed=3)
    loss, grads, new = O.train_step(sd, x, y, kw["num_heads"], dtype=torch.float64)
    _check_step(vitk, model, x, y, float(loss), grads, new)


def test_second_step_uses_updated_weights(vitk):
    kw = dict(image_size=32, patch_size=16, embed_dim=64, num_layers=2, num_heads=1, mlp_dim=128,
              dropout=0.0)
    torch.manual_seed(2)
    model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    tuner = vitk.FineTuner(model, lr=1e-3)
    x, y = O.synthetic_images(8, 32).cuda(), O.synthetic_labels(8).cuda()
    losses = [tuner.step(x, y)[0].item() for _ in range(8)]
    assert losses[-1] < losses[0] - 0.05, losses      # the same batch is being fitted
    # inference through vitk_forward sees the trained weights (shadow repack on version change)
    model.eval()
    with torch.no_grad():
        logits = model(x)
    _, ref = O.classifier_forward({k: v.cpu() for k, v in model.state_dict().items()}, x.cpu(), 1,
                                  dtype=torch.float64)
    assert (logits.cpu().double() - ref).abs().max() < 2e-2


def _library_masks(vitk, p, seed, B, N, D, H, Mlp, L):
    """The keep masks the kernels regenerate from (seed, site, layer, index), as multipliers
    (0 or 1 / (1 - p_quantised)) shaped for the oracle."""
    import ctypes as C
    thresh = int(p * 65536 + 0.5)
    scale = 65536.0 / (65536 - thresh)
    Nk = (N + 15) // 16 * 16
    st = torch.cuda.current_stream().cuda_stream

    def keep(site, layer, n):
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        vitk._lib.check(vitk._lib.lib().vitk_dropout_keep_mask(
            C.c_float(p), seed, site, layer, n, out.data_ptr(), st))
        return out.float() * scale

    masks = {("embed", 0): keep(0, 0, B * N * D).view(B, N, D).cpu()}
    for l in range(L):
        masks[("attn", l)] = keep(1, l, B * H * N * Nk).view(B, H, N, Nk)[..., :N].cpu()
        masks[("proj", l)] = keep(2, l, B * N * D).view(B, N, D).cpu()
        masks[("gelu", l)] = keep(3, l, B * N * Mlp).view(B, N, Mlp).cpu()
        masks[("fc2", l)] = keep(4, l, B * N * D).view(B, N, D).cpu()
    return masks


@pytest.mark.parametrize("p,deit,size", [(0.1, False, 96), (0.25, True, 96),
                                         # 170 tokens: two key tiles, three 64-query sub-tiles
                                         (0.1, False, 208)])
def test_train_step_with_dropout_matches_oracle_given_the_same_masks(vitk, p, deit, size):
    """nn.Dropout at all five sites of train.py (embedding, attention probabilities, projection,
    GELU, linear2): the kernels regenerate their masks from (seed, site, layer, index); injecting
    the same masks into the oracle must reproduce loss, every gradient and the AdamW update."""
    kw = dict(image_size=size, patch_size=16, embed_dim=128, num_layers=2, num_heads=2, mlp_dim=256,
              dropout=p)
    torch.manual_seed(33)
    model = vitk.ViTClassifier(num_classes=6, deit=deit, **kw)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    B, N = 5, (size // 16) ** 2 + (2 if deit else 1)
    x, y = O.synthetic_images(B, size, seed=17), O.synthetic_labels(B, 6, seed=4)
    masks = _library_masks(vitk, p, 77, B, N, 128, 2, 256, 2)
    keep_rate = masks[("gelu", 0)].ne(0).float().mean().item()
    assert abs(keep_rate - (1 - p)) < 0.02, keep_rate
    loss, grads, new = O.train_step(sd, x, y, 2, dtype=torch.float64, masks=masks)
    loss0, _, _ = O.train_step(sd, x, y, 2, dtype=torch.float64)
    assert abs(float(loss) - float(loss0)) > 1e-3          # the masks really change the step
    model = model.cuda().train()
    tuner = vitk.FineTuner(model, lr=1e-4, weight_decay=1e-4, seed=77)
    got_loss, _ = tuner.step(x.cuda(), y.cuda())
    assert abs(got_loss.item() - float(loss)) < 2e-2, (got_loss.item(), float(loss))
    got = dict(zip(tuner.state.names, [q.grad for q in tuner.state.params]))
    worst = 0.0
    for k, gr in grads.items():
        a, b = got[k].detach().cpu().double().reshape(-1), gr.double().reshape(-1)
        rel = (a - b).norm() / (b.norm() + 1e-12)
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0)
        worst = max(worst, rel.item())
        assert rel < GRAD_REL_TOL and cos > 0.999, (k, rel.item(), cos.item())
    print(f"dropout {p}: worst relative gradient error {worst:.4f}")


def test_dropout_follows_train_eval_mode_and_the_seed(vitk):
    kw = dict(image_size=32, patch_size=16, embed_dim=64, num_layers=2, num_heads=1, mlp_dim=128,
              dropout=0.2)
    x, y = O.synthetic_images(8, 32).cuda(), O.synthetic_labels(8).cuda()

    def first_loss(mode_train, seed):
        torch.manual_seed(2)
        model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
        model.train(mode_train)
        return vitk.FineTuner(model, seed=seed).step(x, y)[0].item()

    same = lambda a, b: abs(a - b) < 1e-5     # the loss is summed with float atomics
    assert same(first_loss(True, 1), first_loss(True, 1))          # deterministic given the seed
    assert not same(first_loss(True, 1), first_loss(True, 2))      # the seed selects the masks
    assert same(first_loss(False, 1), first_loss(False, 2))        # eval(): dropout is the identity
    assert not same(first_loss(False, 1), first_loss(True, 1))


def test_uncovered_shapes_fail_loudly(vitk):
    """226 tokens with dropout: beyond the 208-token tcgen05 softmax, served by the CUDA-core
    attention kernels.  What nothing covers raises instead of falling back: a head dimension that
    is not a multiple of 8."""
    kw = dict(image_size=240, patch_size=16, embed_dim=64, num_layers=1, num_heads=1, mlp_dim=64)
    x, y = O.synthetic_images(2, 240).cuda(), O.synthetic_labels(2).cuda()
    for p in (0.0, 0.1):
        loss, _ = vitk.FineTuner(vitk.ViTClassifier(num_classes=6, dropout=p, **kw).cuda().train()).step(x, y)
        assert bool(torch.isfinite(loss).all())
    bad = dict(kw, embed_dim=48, num_heads=4)          # head_dim 12
    tuner = vitk.FineTuner(vitk.ViTClassifier(num_classes=6, dropout=0.0, **bad).cuda().train())
    with pytest.raises(vitk.VitkError):
        tuner.step(x, y)


def test_vit_b16_three_steps_track_the_oracle(vitk):
    """ViT-B/16 at full width: 3 consecutive fused steps (batch 8) follow the oracle's fp32 loss
    trajectory (the oracle's tensor arithmetic is evaluated on the GPU here only to keep the test
    fast; it is the same restatement that is pinned against the reference's goldens on CPU)."""
    kw = dict(image_size=224, patch_size=16, embed_dim=768, num_layers=12, num_heads=12,
              mlp_dim=3072, dropout=0.0)
    torch.manual_seed(4)
    model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    x, y = O.synthetic_images(8, 224, seed=9).cuda(), O.synthetic_labels(8).cuda()
    m = {k: torch.zeros_like(v) for k, v in sd.items()}
    v2 = {k: torch.zeros_like(v) for k, v in sd.items()}
    ref_losses = []
    for step in range(1, 4):
        params = {k: t.clone().requires_grad_(True) for k, t in sd.items()}
        _, logits = O.classifier_forward(params, x, 12)
        loss = O.cross_entropy(logits, y)
        grads = dict(zip(params, torch.autograd.grad(loss, list(params.values()))))
        ref_losses.append(loss.item())
        O.adamw_step(sd, grads, m, v2, step)
    tuner = vitk.FineTuner(model)
    got = [tuner.step(x, y)[0].item() for _ in range(3)]
    print("oracle losses", ref_losses, "vitk losses", got)
    for a, b in zip(got, ref_losses):
        assert abs(a - b) < 0.05 * max(1.0, abs(b)), (got, ref_losses)


def test_checkpoints_round_trip_with_torch_adamw(vitk, tmp_path):
    """train.py:1647-1654 layout: model_state_dict keys of the reference, optimizer_state_dict in
    torch.optim.AdamW's format.  (a) a FineTuner resumed from its own checkpoint continues exactly;
    (b) torch.optim.AdamW accepts the optimizer state and its next step lands on the same weights
    as the fused kernel's; (c) a checkpoint written with torch.optim.AdamW resumes in FineTuner."""
    kw = dict(image_size=32, patch_size=16, embed_dim=64, num_layers=2, num_heads=1, mlp_dim=128,
              dropout=0.0)
    x, y = O.synthetic_images(8, 32).cuda(), O.synthetic_labels(8).cuda()
    torch.manual_seed(3)
    model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    tuner = vitk.FineTuner(model, lr=1e-3)
    for _ in range(2):
        tuner.step(x, y)
    path = tmp_path / "ck.pth"
    tuner.save_checkpoint(path, epoch=1, val_loss=0.5)
    ck = torch.load(path, weights_only=False)
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "val_loss", "config"}
    assert "backbone.transformer_blocks.0.attention.qkv.weight" in ck["model_state_dict"]
    # (a)
    model2 = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    tuner2 = vitk.FineTuner(model2, lr=1e-3)
    tuner2.load_checkpoint(path)
    assert tuner2.state.step_count == 2
    tuner.step(x, y)
    tuner2.step(x, y)
    for (k, a), (_, b) in zip(model.state_dict().items(), model2.state_dict().items()):
        assert (a - b).abs().max() < 1e-6, k
    # (b) torch's AdamW continues from the same state with the gradients of the fused backward
    model3 = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    model3.load_state_dict(ck["model_state_dict"])
    opt = torch.optim.AdamW(model3.parameters(), lr=1e-3, weight_decay=1e-4)
    opt.load_state_dict(ck["optimizer_state_dict"])
    model4 = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    tuner4 = vitk.FineTuner(model4, lr=1e-3)
    tuner4.load_checkpoint(path)
    tuner4.step(x, y)
    grads = [p.grad.clone() for p in tuner4.state.params]
    for p, g in zip(model3.parameters(), grads):
        p.grad = g
    opt.step()
    for (k, a), (_, b) in zip(model3.state_dict().items(), model4.state_dict().items()):
        assert (a - b).abs().max() < 2e-6, k
    # (c) torch-written optimizer state resumes in the fused trainer
    path2 = tmp_path / "ref.pth"
    torch.save({"epoch": 7, "model_state_dict": model3.state_dict(),
                "optimizer_state_dict": opt.state_dict(), "val_loss": 0.1, "config": {}}, path2)
    model5 = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    tuner5 = vitk.FineTuner(model5, lr=5e-4)
    assert tuner5.load_checkpoint(path2)["epoch"] == 7
    assert tuner5.state.step_count == 3 and tuner5.lr == 1e-3
    tuner5.step(x, y)


def test_postprocess_scores_matches_the_reference_loop(vitk):
    """evaluation.py:403-404 on detector-shaped logits [B, Q, C+1] and on classifier logits."""
    g = torch.Generator(device="cuda").manual_seed(0)
    logits = torch.randn(3, 100, 7, generator=g, device="cuda") * 3
    scores, labels, probs = vitk.ops.postprocess_scores(logits, exclude_last=True, return_probs=True)
    ref = torch.softmax(logits, dim=-1)
    want_s, want_l = torch.max(ref[..., :-1], dim=-1)
    torch.testing.assert_close(probs, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(scores, want_s, rtol=1e-5, atol=1e-6)
    assert torch.equal(labels, want_l)
    s2, l2 = vitk.ops.postprocess_scores(logits[:, 0, :6].contiguous())
    assert torch.equal(l2, logits[:, 0, :6].argmax(-1))


def test_eval_between_steps_sees_every_update(vitk):
    """train -> validate -> train -> validate (train.py:1619-1631): the inference engine packs its
    bf16 weight copies at the first validate; the fused optimizer updates the arena through raw
    pointers, so every later validate must still see the CURRENT weights."""
    kw = dict(image_size=32, patch_size=16, embed_dim=64, num_layers=2, num_heads=1, mlp_dim=128,
              dropout=0.0)
    torch.manual_seed(5)
    model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    tuner = vitk.FineTuner(model, lr=3e-3)
    x, y = O.synthetic_images(8, 32).cuda(), O.synthetic_labels(8).cuda()
    for epoch in range(3):
        model.train()
        for _ in range(3):
            tuner.step(x, y)
        model.eval()
        with torch.no_grad():
            logits = model(x)                  # epoch 0 packs; later epochs must repack
        _, ref = O.classifier_forward({k: v.cpu() for k, v in model.state_dict().items()},
                                      x.cpu(), 1, dtype=torch.float64)
        assert (logits.cpu().double() - ref).abs().max() < 2e-2, epoch


def test_frozen_parameters_are_left_alone(vitk):
    """torch.optim.AdamW(model.parameters()) skips parameters without a gradient: with the
    backbone frozen only the head may move (no Adam update, no weight decay on the rest)."""
    kw = dict(image_size=32, patch_size=16, embed_dim=64, num_layers=2, num_heads=1, mlp_dim=128,
              dropout=0.0)
    torch.manual_seed(6)
    model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
    for p in model.backbone.parameters():
        p.requires_grad_(False)
    model.backbone.layer_norm.weight.requires_grad_(True)   # a trainable island inside the arena
    before = {k: v.clone() for k, v in model.state_dict().items()}
    tuner = vitk.FineTuner(model, lr=1e-2, weight_decay=1e-1)
    x, y = O.synthetic_images(8, 32).cuda(), O.synthetic_labels(8).cuda()
    for _ in range(2):
        tuner.step(x, y)
    after = model.state_dict()
    for k in before:
        moved = not torch.equal(before[k], after[k])
        trainable = k.startswith("head.") or k == "backbone.layer_norm.weight"
        assert moved == trainable, k


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_vit_b16_full_width_gradients_match_oracle(vitk, dropout):
    """Every parameter gradient of ViT-B/16 (12 layers, 197 tokens, batch 8) against the oracle's
    autograd in fp32 (evaluated on the GPU to keep the test fast - the same restatement the CPU
    tests pin to the reference), without dropout and with the reference's 0.1 (the library's masks
    injected into the oracle)."""
    kw = dict(image_size=224, patch_size=16, embed_dim=768, num_layers=12, num_heads=12,
              mlp_dim=3072, dropout=dropout)
    torch.manual_seed(14)
    model = vitk.ViTClassifier(num_classes=6, **kw).cuda().train()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    B = 8
    x, y = O.synthetic_images(B, 224, seed=19).cuda(), O.synthetic_labels(B, 6, seed=6).cuda()
    masks = None
    if dropout > 0:
        masks = {k: v.cuda() for k, v in
                 _library_masks(vitk, dropout, 31, B, 197, 768, 12, 3072, 12).items()}
    params = {k: t.clone().requires_grad_(True) for k, t in sd.items()}
    _, logits = O.classifier_forward(params, x, 12, masks=masks)
    loss = O.cross_entropy(logits, y)
    g_ref = dict(zip(params, torch.autograd.grad(loss, list(params.values()))))
    tuner = vitk.FineTuner(model, seed=31)
    got_loss, _ = tuner.step(x, y)
    assert abs(got_loss.item() - loss.item()) < 2e-2
    got = dict(zip(tuner.state.names, [q.grad for q in tuner.state.params]))
    worst, worst_k = 0.0, None
    for k, gr in g_ref.items():
        a, b = got[k].detach().double().reshape(-1), gr.double().reshape(-1)
        rel = ((a - b).norm() / (b.norm() + 1e-30)).item()
        cos = torch.nn.functional.cosine_similarity(a, b, dim=0).item()
        if rel > worst:
            worst, worst_k = rel, k
        assert rel < GRAD_REL_TOL_FULL_WIDTH and cos > 0.999, (k, rel, cos)
    print(f"ViT-B/16 dropout {dropout}: worst relative gradient error {worst:.4f} ({worst_k})")


def test_nonfinite_gradients_skip_the_step(vitk):
    """scaler.step(optimizer) of the reference (train.py:1456) skips the update on inf / nan
    gradients: parameters and Adam moments stay put, the step counter does not advance - the
    next good step equals the first step of a run that never saw the bad batch."""
    kw = dict(image_size=32, patch_size=16, embed_dim=64, num_layers=2, num_heads=1, mlp_dim=128,
              dropout=0.0)
    x, y = O.synthetic_images(8, 32).cuda(), O.synthetic_labels(8).cuda()
    bad = x.clone()
    bad[3, 1, 5, 7] = float("inf")

    def fresh():
        torch.manual_seed(8)
        model = vitk.ViTClassifier(num_classes=6, **kw).cuda()
        return model, vitk.FineTuner(model, lr=1e-3)

    model_a, tuner_a = fresh()
    before = {k: v.clone() for k, v in model_a.state_dict().items()}
    tuner_a.step(bad, y)
    assert tuner_a.skipped_steps == 1
    assert all(torch.equal(before[k], v) for k, v in model_a.state_dict().items())
    assert float(tuner_a.state.exp_avg.abs().sum()) == 0.0
    tuner_a.step(x, y)                              # step counter 2, one skipped -> Adam step 1
    assert tuner_a.skipped_steps == 1
    model_b, tuner_b = fresh()
    tuner_b.step(x, y)
    for (k, a), b in zip(model_a.state_dict().items(), model_b.state_dict().values()):
        torch.testing.assert_close(a, b, rtol=0, atol=1e-7, msg=k)
    # without the guard the same batch poisons the weights
    model_c, _ = fresh()
    vitk.FineTuner(model_c, lr=1e-3, skip_nonfinite=False).step(bad, y)
    assert not all(bool(torch.isfinite(v).all()) for v in model_c.state_dict().values())


@pytest.mark.parametrize("dropout", [0.0, 0.1])
def test_reference_config_trains_one_step(vitk, dropout):
    """The configuration train.py actually ships (train.py:1345-1356): DeiT, D = 400, 25 heads of
    dimension 16, MLP 1600, 198 tokens - here with 3 of its 12 layers - one fused step against the
    oracle, with and without the reference's dropout."""
    kw = dict(image_size=224, patch_size=16, embed_dim=400, num_layers=3, num_heads=25,
              mlp_dim=1600, dropout=dropout)
    torch.manual_seed(41)
    model = vitk.ViTClassifier(num_classes=6, deit=True, **kw)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    B = 4
    x, y = O.synthetic_images(B, 224, seed=23), O.synthetic_labels(B, 6, seed=9)
    masks = _library_masks(vitk, dropout, 13, B, 198, 400, 25, 1600, 3) if dropout > 0 else None
    loss, grads, new = O.train_step(sd, x, y, 25, dtype=torch.float64, masks=masks)
    model = model.cuda().train()
    tuner = vitk.FineTuner(model, lr=1e-4, weight_decay=1e-4, seed=13)
    got_loss, _ = tuner.step(x.cuda(), y.cuda())
    assert abs(got_loss.item() - float(loss)) < 2e-2, (got_loss.item(), float(loss))
    got = dict(zip(tuner.state.names, [q.grad for q in tuner.state.params]))
    worst = 0.0
    for k, gr in grads.items():
        a, b = got[k].detach().cpu().double().reshape(-1), gr.double().reshape(-1)
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        worst = max(worst, rel)
        assert rel < GRAD_REL_TOL, (k, rel)
    print(f"train.py Config dims, dropout {dropout}: worst relative gradient error {worst:.4f}")
    # inference at head_dim 16 goes through the same CUDA-core attention
    model.eval()
    with torch.no_grad():
        logits = model(x.cuda()).cpu().double()
    _, ref = O.classifier_forward({k: v.cpu() for k, v in model.state_dict().items()}, x, 25,
                                  dtype=torch.float64)
    assert (logits - ref).abs().max() < 2e-2


def test_384px_fine_tune_step(vitk):
    """577 tokens (image_size is a constructor argument of the reference, train.py:638): a training
    step beyond the 256-token tcgen05 backward, with dropout beyond the 208-token softmax kernel."""
    kw = dict(image_size=384, patch_size=16, embed_dim=128, num_layers=2, num_heads=2, mlp_dim=256,
              dropout=0.1)
    torch.manual_seed(43)
    model = vitk.ViTClassifier(num_classes=6, **kw)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    B = 2
    x, y = O.synthetic_images(B, 384, seed=29), O.synthetic_labels(B, 6, seed=2)
    masks = _library_masks(vitk, 0.1, 17, B, 577, 128, 2, 256, 2)
    loss, grads, _ = O.train_step(sd, x, y, 2, dtype=torch.float64, masks=masks)
    tuner = vitk.FineTuner(model.cuda().train(), seed=17)
    got_loss, _ = tuner.step(x.cuda(), y.cuda())
    assert abs(got_loss.item() - float(loss)) < 2e-2
    got = dict(zip(tuner.state.names, [q.grad for q in tuner.state.params]))
    for k, gr in grads.items():
        a, b = got[k].detach().cpu().double().reshape(-1), gr.double().reshape(-1)
        rel = ((a - b).norm() / (b.norm() + 1e-12)).item()
        assert rel < GRAD_REL_TOL, (k, rel)
