"""Training through the detection head: vitk_detection_head_forward_train
+ vitk_detection_head_backward behind autograd against the oracle's restatement of
ObjectDetectionHead.forward (train.py:691-731) differentiated by torch autograd in fp64, and the
device-side SetCriterion.loss_labels (train.py:1220-1239) against F.cross_entropy."""
import pytest
import torch
import torch.nn.functional as F

from oracle import vit_oracle as O

pytestmark = pytest.mark.gpu


def _head(vitk, D, Q, n_classes, seed=1, layers=None):
    torch.manual_seed(0)
    head = vitk.ObjectDetectionHead(embed_dim=D, num_classes=n_classes, num_queries=Q)
    if layers is not None:
        head.decoder.layers = torch.nn.ModuleList(list(head.decoder.layers)[:layers])
        head.decoder.num_layers = layers
    sd = O.randomize_head_state(head.state_dict(), seed)
    head.load_state_dict(sd)
    return head, sd


def _no_dropout(module):
    for m in module.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0


def _oracle_grads(sd, tokens, skip, r_logits, r_boxes):
    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    tok = tokens.double().clone().requires_grad_(True)
    out = O.detection_head_forward(sd64, tok[:, skip:, :], dtype=torch.float64)
    loss = (out["class_logits"] * r_logits.double()).sum() + (out["bbox_coords"] * r_boxes.double()).sum()
    loss.backward()
    return out, {k: v.grad for k, v in sd64.items()}, tok.grad


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), \
        (torch.dot(a, b) / (a.norm() * b.norm()).clamp_min(1e-30)).item()


@pytest.mark.parametrize("D,Q,P,B,skip,L,impl", [
    (256, 20, 30, 3, 1, 2, 0),      # head_dim 32, ragged sizes
    (256, 20, 30, 3, 1, 2, 1),      # the CUDA-core forward attention
    (768, 100, 196, 2, 1, 6, 0),    # the reference's geometry: 100 queries, 196 patch tokens, 6 layers
    (768, 100, 196, 2, 2, 2, 0),    # DeiT: CLS and DIST rows skipped
    (512, 130, 40, 2, 0, 1, 0),     # more queries than one tensor-core tile -> generic forward
])
def test_head_backward_matches_oracle_autograd(vitk, D, Q, P, B, skip, L, impl):
    head, sd = _head(vitk, D, Q, 6, layers=L)
    _no_dropout(head)
    head = head.cuda().train()
    g = torch.Generator().manual_seed(3)
    tokens = torch.randn(B, P + skip, D, generator=g)
    r_logits = torch.randn(B, Q, 7, generator=g)
    r_boxes = torch.randn(B, Q, 4, generator=g)
    ref_out, ref_g, ref_dtok = _oracle_grads(sd, tokens, skip, r_logits, r_boxes)

    vitk._lib.set_attention_impl(impl)
    try:
        tok = tokens.cuda().requires_grad_(True)
        out = head.decode(tok, skip_tokens=skip)
        loss = (out["class_logits"] * r_logits.cuda()).sum() + (out["bbox_coords"] * r_boxes.cuda()).sum()
        loss.backward()
    finally:
        vitk._lib.set_attention_impl(0)
    torch.cuda.synchronize()
    assert (out["class_logits"].detach().cpu().double() - ref_out["class_logits"]).abs().max() < 2e-2
    assert (out["bbox_coords"].detach().cpu().double() - ref_out["bbox_coords"]).abs().max() < 1e-2

    # bf16 gradients between the GEMMs: the error grows with the number of layers a gradient has
    # crossed and shrinks with the width (observed worst: 6.7e-2 for object_queries - the deepest
    # gradient - at 6 layers of width 768; 4.2e-2 at 2 layers of width 256; cos >= 0.9977)
    tol = 9e-2
    worst = (0.0, "")
    for name, p in head.named_parameters():
        assert p.grad is not None, name
        rel, cos = _rel(p.grad.cpu(), ref_g[name])
        worst = max(worst, (rel, name))
        assert cos > 0.995 and rel < tol, (name, rel, cos)
    rel, cos = _rel(tok.grad.cpu(), ref_dtok)
    print("worst parameter gradient error", worst, "d tokens", rel, cos)
    assert cos > 0.995 and rel < tol, ("d_tokens", rel, cos)
    if skip:
        assert tok.grad[:, :skip].abs().max().item() == 0.0     # the memory starts after the prefix


def test_head_forward_train_equals_inference_forward(vitk):
    head, _ = _head(vitk, 768, 100, 6)
    _no_dropout(head)
    head = head.cuda()
    tokens = torch.randn(2, 197, 768, generator=torch.Generator().manual_seed(5)).cuda()
    with torch.no_grad():
        ref = head.eval().decode(tokens, skip_tokens=1)
    out = head.train().decode(tokens.clone().requires_grad_(True), skip_tokens=1)
    assert torch.equal(out["class_logits"].detach(), ref["class_logits"])
    assert torch.equal(out["bbox_coords"].detach(), ref["bbox_coords"])


def test_detector_trains_end_to_end(vitk):
    """model(images) -> loss -> backward() -> torch.optim step, the reference's loop
    (train.py:1441-1460), through the encoder bridge AND the head: the loss goes down."""
    torch.manual_seed(0)
    model = vitk.DeiTObjectDetector(image_size=32, patch_size=16, embed_dim=256, num_layers=2,
                                    num_heads=4, mlp_dim=512, dropout=0.0, num_classes=6,
                                    num_queries=10).cuda().train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    x = O.synthetic_images(4, 32).cuda()
    tgt_cls = torch.randint(0, 7, (4, 10), generator=torch.Generator().manual_seed(1)).cuda()
    tgt_box = torch.rand(4, 10, 4, generator=torch.Generator().manual_seed(2)).cuda()
    w = torch.ones(7, device="cuda")
    w[-1] = 0.1
    losses = []
    for _ in range(8):
        opt.zero_grad(set_to_none=True)
        pred, triplet = model(x)
        loss = vitk.weighted_cross_entropy(pred["class_logits"], tgt_cls, w) + \
            F.l1_loss(pred["bbox_coords"], tgt_box) + 0.01 * triplet.pow(2).sum()
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(p.grad is not None for p in model.parameters())
    assert losses[-1] < losses[0] - 0.05, losses


@pytest.mark.parametrize("rows_shape,C", [((2, 100), 7), ((5, 7), 81), ((1, 3), 2)])
def test_weighted_cross_entropy_matches_torch(vitk, rows_shape, C):
    g = torch.Generator().manual_seed(0)
    logits = (torch.randn(*rows_shape, C, generator=g) * 3).cuda().requires_grad_(True)
    targets = torch.randint(0, C, rows_shape, generator=g).cuda()
    targets.view(-1)[0] = C - 1                     # at least one background prediction
    weight = torch.ones(C, device="cuda")
    weight[-1] = 0.1                                # train.py:1215-1217
    ref_in = logits.detach().clone().requires_grad_(True)
    ref = F.cross_entropy(ref_in.transpose(1, 2), targets, weight)     # train.py:1236
    ref.backward()
    loss = vitk.weighted_cross_entropy(logits, targets, weight)
    (loss * 2.0).backward()
    torch.testing.assert_close(loss.detach(), ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(logits.grad, 2.0 * ref_in.grad, rtol=1e-4, atol=1e-6)
    # without weights: the plain mean
    l2 = vitk.weighted_cross_entropy(logits.detach(), targets, None)
    torch.testing.assert_close(l2, F.cross_entropy(logits.detach().transpose(1, 2), targets),
                               rtol=1e-5, atol=1e-6)


def _head_masks(vitk, p, seed, B, Q, P, D, F, H, L):
    """The library's own keep masks (vitk_dropout_keep_mask) for the six dropout sites of every
    decoder layer, scaled by 1 / (1 - p), in the oracle's layout."""
    import ctypes as C
    st = torch.cuda.current_stream().cuda_stream
    t = int(p * 65536 + 0.5)
    scale = 65536.0 / (65536 - t)

    def keep(site, layer, n):
        out = torch.empty(n, dtype=torch.uint8, device="cuda")
        vitk._lib.check(vitk._lib.lib().vitk_dropout_keep_mask(
            C.c_float(p), seed, site, layer, n, out.data_ptr(), st))
        return out.float() * scale

    pad = lambda n: (n + 15) // 16 * 16
    masks = {}
    for l in range(L):
        masks[("dec_sa_attn", l)] = keep(8, l, B * H * Q * pad(Q)).view(B, H, Q, pad(Q))[..., :Q].cpu()
        masks[("dec_sa_out", l)] = keep(9, l, B * Q * D).view(B, Q, D).cpu()
        masks[("dec_ca_attn", l)] = keep(10, l, B * H * Q * pad(P)).view(B, H, Q, pad(P))[..., :P].cpu()
        masks[("dec_ca_out", l)] = keep(11, l, B * Q * D).view(B, Q, D).cpu()
        masks[("dec_ffn", l)] = keep(12, l, B * Q * F).view(B, Q, F).cpu()
        masks[("dec_ff2", l)] = keep(13, l, B * Q * D).view(B, Q, D).cpu()
    return masks


@pytest.mark.parametrize("p,D,Q,P,B,skip,L", [(0.1, 256, 20, 30, 3, 1, 2), (0.25, 768, 100, 196, 2, 1, 2)])
def test_head_with_dropout_matches_oracle_given_the_same_masks(vitk, p, D, Q, P, B, skip, L):
    """nn.TransformerDecoderLayer's dropout at all six sites per layer (train.py:701-707): the
    kernels regenerate their masks from (seed, site, layer, index); the same masks injected into
    the oracle must reproduce the outputs, every parameter gradient and d features."""
    head, sd = _head(vitk, D, Q, 6, layers=L)
    for ly in head.decoder.layers:
        for m in (ly.dropout, ly.dropout1, ly.dropout2, ly.dropout3):
            m.p = p
        ly.self_attn.dropout = ly.multihead_attn.dropout = p
    head = head.cuda().train()
    seed = 4242
    head.__dict__["_vitk_forced_seed"] = seed
    F_ = head.decoder.layers[0].linear1.out_features
    masks = _head_masks(vitk, p, seed, B, Q, P, D, F_, 8, L)
    g = torch.Generator().manual_seed(3)
    tokens = torch.randn(B, P + skip, D, generator=g)
    r_logits = torch.randn(B, Q, 7, generator=g)
    r_boxes = torch.randn(B, Q, 4, generator=g)

    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    tok64 = tokens.double().clone().requires_grad_(True)
    ref = O.detection_head_forward(sd64, tok64[:, skip:, :], dtype=torch.float64, masks=masks)
    ((ref["class_logits"] * r_logits.double()).sum() + (ref["bbox_coords"] * r_boxes.double()).sum()).backward()

    tok = tokens.cuda().requires_grad_(True)
    out = head.decode(tok, skip_tokens=skip)
    ((out["class_logits"] * r_logits.cuda()).sum() + (out["bbox_coords"] * r_boxes.cuda()).sum()).backward()
    torch.cuda.synchronize()
    # bf16 ReLU outputs within rounding of zero flip their mask; everything else follows the oracle
    assert (out["class_logits"].detach().cpu().double() - ref["class_logits"]).abs().max() < 3e-2
    assert (out["bbox_coords"].detach().cpu().double() - ref["bbox_coords"]).abs().max() < 1e-2
    worst = (0.0, "")
    for name, prm in head.named_parameters():
        rel, cos = _rel(prm.grad.cpu(), sd64[name].grad)
        worst = max(worst, (rel, name))
        assert cos > 0.995 and rel < 9e-2, (name, rel, cos)
    rel, cos = _rel(tok.grad.cpu(), tok64.grad)
    print("dropout", p, "worst parameter gradient error", worst, "d tokens", rel, cos)
    assert cos > 0.995 and rel < 9e-2
    # and the masks really were applied: the eval-mode output differs
    with torch.no_grad():
        ev = head.eval().decode(tokens.cuda(), skip_tokens=skip)
    assert (ev["class_logits"] - out["class_logits"].detach()).abs().max() > 1e-2
