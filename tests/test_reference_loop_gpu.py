"""The reference's OWN training loop on the vitk detector (drop-in claim of SURVEY.md 8 row b under
training): `train_one_epoch` (train.py:1425-1479) is imported from the reference and run unchanged
- model(images, return_features=True), ObjectDetectionLoss with the HungarianMatcher, the triplet
loss, losses.backward(), optimizer.step() - once on the reference's DeiTObjectDetector (fp32, every
dropout probability set to 0) and once on vitk.DeiTObjectDetector loaded with the same state_dict.
The mean loss of an epoch of two batches must agree: the second batch's loss depends on the first
optimizer step, i.e. on every gradient that reached the parameters through the encoder bridge and
the detection head."""
import pytest
import torch

from oracle import ref_loader

pytestmark = pytest.mark.gpu

KW = dict(image_size=64, patch_size=16, embed_dim=256, num_layers=2, num_heads=4, mlp_dim=512,
          dropout=0.0, num_classes=6, num_queries=10)


def _no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, torch.nn.MultiheadAttention):
            m.dropout = 0.0


def _batches(device):
    g = torch.Generator().manual_seed(11)
    out = []
    for _ in range(2):
        images = torch.randn(4, 3, 64, 64, generator=g)
        targets = []
        for i in range(4):
            n = 1 + (i % 3)
            lo = torch.rand(n, 2, generator=g) * 0.4
            wh = 0.2 + torch.rand(n, 2, generator=g) * 0.3
            targets.append({"labels": torch.randint(0, 6, (n,), generator=g),
                            "boxes": torch.cat([lo, lo + wh], dim=1)})      # x1 y1 x2 y2 in [0, 1]
        out.append((images, targets))
    return out


@pytest.mark.skipif(not ref_loader.reference_available(), reason="reference scripts not shipped")
def test_reference_train_one_epoch_runs_on_the_vitk_detector(vitk, monkeypatch):
    ref = ref_loader.load("train")
    # the loop logs to wandb every 100 batches (train.py:1472): no run is open here
    monkeypatch.setattr(ref.wandb, "log", lambda *a, **k: None, raising=False)
    dev = torch.device("cuda")
    ref.Config.DEVICE = dev

    torch.manual_seed(0)
    theirs = ref.DeiTObjectDetector(**KW)
    _no_dropout(theirs)
    sd = {k: v.clone() for k, v in theirs.state_dict().items()}
    # predicted boxes valid (x2 > x1, y2 > y1) with a margin: the matcher drops invalid ones, a
    # discontinuity that a last-bit difference could otherwise flip
    sd["detection_head.bbox_head.weight"] *= 0.1
    sd["detection_head.bbox_head.bias"] = torch.tensor([-1.0, -1.0, 1.0, 1.0])
    theirs.load_state_dict(sd)
    mine = vitk.DeiTObjectDetector(**KW)
    _no_dropout(mine)               # (the mirror applies the decoder layers' dropout in train())
    mine.load_state_dict(sd)

    def epoch(model):
        model = model.to(dev)
        matcher = ref.HungarianMatcher(cost_class=1.0, cost_bbox=5.0, cost_giou=2.0)
        criterion = ref.ObjectDetectionLoss(num_classes=6, matcher=matcher,
                                            weight_dict=ref.Config.WEIGHT_DICT,
                                            use_triplet_loss=True, triplet_margin=0.3,
                                            triplet_mining="batch_hard").to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)   # train.py:1598-1602
        losses = [ref.train_one_epoch(model, _batches(dev), opt, criterion, dev, e, scaler=None)
                  for e in range(3)]
        return losses

    want = epoch(theirs)
    n0 = vitk.launch_count()
    got = epoch(mine)
    assert vitk.launch_count() - n0 > 500        # the vitk kernels did the work
    print("reference model:", want, " vitk model:", got)
    # observed: [6.84827, 5.66106, 5.16295] against [6.84825, 5.66127, 5.16341] (1e-4 relative)
    for a, b in zip(got, want):
        assert abs(a - b) <= 5e-3 * abs(b), (got, want)
    assert got[-1] < got[0]
