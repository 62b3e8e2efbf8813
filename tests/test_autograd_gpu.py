"""The autograd bridge: backbone(images) is differentiable, so the reference's own training loop
(any loss on top of the tokens + a torch optimizer) runs unchanged (train.py:831,1455-1460)."""
import pytest
import torch

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu

KW = dict(image_size=64, patch_size=16, embed_dim=128, num_layers=2, num_heads=2, mlp_dim=256,
          dropout=0.0)


def test_backbone_gradients_match_oracle(vitk):
    torch.manual_seed(0)
    bb = vitk.VisionTransformer(**KW)
    sd = {k: v.clone() for k, v in bb.state_dict().items()}
    x = O.synthetic_images(3, 64)
    w = torch.randn(17, 128, dtype=torch.float64)        # an arbitrary loss on every token
    params = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    loss_ref = (O.backbone_forward(params, x, 2, dtype=torch.float64) * w).sum()
    g_ref = dict(zip(params, torch.autograd.grad(loss_ref, list(params.values()))))
    bb = bb.cuda().train()
    tokens = bb(x.cuda())
    assert tokens.requires_grad
    (tokens * w.cuda().float()).sum().backward()
    assert abs((tokens.detach().cpu().double() * w).sum().item() - loss_ref.item()) < 0.05 * (1 + abs(loss_ref.item()))
    worst = 0.0
    for k, p in bb.named_parameters():
        got, want = p.grad.cpu().double().flatten(), g_ref[k].flatten()
        rel = (got - want).norm() / (want.norm() + 1e-12)
        cos = torch.nn.functional.cosine_similarity(got, want, dim=0)
        worst = max(worst, rel.item())
        assert rel < 0.012 and cos > 0.999, (k, rel.item(), cos.item())
    print(f"autograd bridge: worst relative gradient error {worst:.4f}")


def test_reference_style_loop_with_torch_optimizer(vitk):
    """zero_grad -> forward -> loss.backward -> optimizer.step with torch.optim.AdamW."""
    torch.manual_seed(1)
    model = vitk.ViTClassifier(num_classes=6, **KW).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    x, y = O.synthetic_images(8, 64).cuda(), O.synthetic_labels(8).cuda()
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < losses[0] - 0.05, losses
    model.eval()
    with torch.no_grad():
        logits = model(x)                      # inference path repacks the updated weights
    _, ref = O.classifier_forward({k: v.cpu() for k, v in model.state_dict().items()}, x.cpu(), 2,
                                  dtype=torch.float64)
    assert (logits.cpu().double() - ref).abs().max() < 2e-2


def test_two_forwards_before_backward(vitk):
    """Siamese / triplet use (and micro-batches whose losses are summed): each forward keeps its
    own saved activations, also when the batch sizes differ."""
    torch.manual_seed(3)
    bb = vitk.VisionTransformer(**KW).cuda().train()
    xa, xb = O.synthetic_images(3, 64, seed=1).cuda(), O.synthetic_images(5, 64, seed=2).cuda()
    wa = torch.randn(3, 17, 128, device="cuda")
    wb = torch.randn(5, 17, 128, device="cuda")

    def grads(losses_fn):
        for p in bb.parameters():
            p.grad = None
        losses_fn().backward()
        return [p.grad.clone() for p in bb.parameters()]

    g_a = grads(lambda: (bb(xa) * wa).sum())
    g_b = grads(lambda: (bb(xb) * wb).sum())
    g_ab = grads(lambda: (bb(xa) * wa).sum() + (bb(xb) * wb).sum())   # both forwards, then backward
    for ga, gb, gab in zip(g_a, g_b, g_ab):
        torch.testing.assert_close(gab, ga + gb, rtol=1e-4, atol=1e-5)


def test_fp32_mode_refuses_the_grad_path(vitk):
    bb = vitk.VisionTransformer(**KW).cuda().eval().set_precision("fp32")
    with pytest.raises(vitk.VitkError):
        bb(O.synthetic_images(2, 64).cuda())
    with torch.no_grad():
        assert bb(O.synthetic_images(2, 64).cuda()).shape == (2, 17, 128)
