"""CPU tier: the multi-GPU host logic (batch sharding, gradient slice all-reduce, logit gather)
exercised with world_size 2 and 3 on the gloo backend."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, globals()[fn_name](rank, world)))
    finally:
        dist.destroy_process_group()


def _run(world, fn_name):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def _dist_mod():
    import importlib
    return importlib.import_module("automated-recycling-sorter-with-vision-transformers_b200.dist")


def test_shard_range_partitions_exactly():
    d = _dist_mod()
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [d.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        d.shard_range(10, 2, 2)


def job_allreduce(rank, world):
    d = _dist_mod()
    torch.manual_seed(0)
    full = torch.randn(world, 1000)              # every rank knows every rank's "gradient"
    flat = full[rank].clone()
    slices = [(900, 1000), (300, 900), (0, 300), (500, 500)]   # completion order, one empty
    d.allreduce_slices(flat, slices)
    return torch.allclose(flat, full.sum(0), atol=1e-6)


def job_pipelined_optimizer(rank, world):
    """on_slice_done: the per-slice update callback (FineTuner runs AdamW on a slice as soon as
    its reduction has finished) sees every non-empty slice once, in order, already summed."""
    d = _dist_mod()
    torch.manual_seed(4)
    full = torch.randn(world, 600)
    flat = full[rank].clone()
    p = torch.zeros(600)
    slices = [(400, 600), (100, 400), (250, 250), (0, 100)]
    seen = []

    def update(k, a, b):
        seen.append(k)
        p[a:b] -= 0.5 * flat[a:b]             # an SGD step on the reduced slice

    d.allreduce_slices(flat, slices, on_slice_done=update)
    return seen == [0, 1, 3] and torch.allclose(p, -0.5 * full.sum(0), atol=1e-6)


def job_data_parallel_mean(rank, world):
    """loss_scale = 1 / global batch + SUM all-reduce == gradient of the global-batch mean."""
    d = _dist_mod()
    torch.manual_seed(1)
    n, feat = 10, 16
    x, y = torch.randn(n, feat), torch.randn(n)
    w = torch.zeros(feat, requires_grad=True)
    ((x @ w - y) ** 2).mean().backward()
    want = w.grad.clone()
    a, b = d.shard_range(n, rank, world)
    w2 = torch.zeros(feat, requires_grad=True)
    (((x[a:b] @ w2 - y[a:b]) ** 2).sum() / n).backward()
    g = w2.grad.clone()
    d.allreduce_slices(g, [(0, feat)])
    return torch.allclose(g, want, atol=1e-6)


def job_sharded_inference(rank, world):
    d = _dist_mod()
    torch.manual_seed(2)
    batch = torch.randn(11, 5)
    w = torch.randn(5, 6)
    fn = lambda t: t @ w                      # stands in for the classifier
    out = d.sharded_apply(fn, batch)
    local = d.sharded_apply(fn, batch, gather=False)
    a, b = d.shard_range(11, rank, world)
    return bool(torch.allclose(out, batch @ w) and local.shape[0] == b - a)


def job_sharded_detector(rank, world):
    """A detector returns the reference's prediction dict; every entry is gathered by image.  Also
    with more ranks than images."""
    d = _dist_mod()
    torch.manual_seed(3)
    w1, w2 = torch.randn(5, 3 * 7), torch.randn(5, 3 * 4)
    fn = lambda t: {"class_logits": (t @ w1).reshape(-1, 3, 7),
                    "bbox_coords": torch.sigmoid(t @ w2).reshape(-1, 3, 4)}
    ok = True
    for n in (11, 1):
        batch = torch.randn(n, 5)
        out, want = d.sharded_apply(fn, batch), fn(batch)
        ok = ok and all(torch.allclose(out[k], want[k]) and out[k].shape == want[k].shape
                        for k in want)
    return bool(ok)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("job", ["job_allreduce", "job_pipelined_optimizer", "job_data_parallel_mean",
                                 "job_sharded_inference", "job_sharded_detector"])
def test_gloo(world, job):
    assert all(_run(world, job).values())


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_owned_pieces_partition_every_slice(world):
    """Ownership map of the peer-memory optimizer: the pieces of every slice are disjoint, in rank
    order, cover the slice exactly and start on 64-element boundaries (16-byte accesses in the fp32
    and bf16 arenas)."""
    from vitk.dist import owned_pieces
    slices = [(0, 64), (64, 64 * 1000 + 64), (64 * 1001, 64 * 1001 + 12352), (76416, 76416)]
    owned = owned_pieces(slices, world, 64)
    assert len(owned) == world and all(len(o) == len(slices) for o in owned)
    for k, (a, b) in enumerate(slices):
        cur = a
        for r in range(world):
            lo, hi = owned[r][k]
            assert lo == cur and lo <= hi <= b
            assert (lo - a) % 64 == 0 or lo == b
            cur = hi
        assert cur == b
