"""Multi-GPU host logic (one process per GPU, torch.distributed for the plumbing).

The reference is single-device (train.py:1392, no DDP).  The hot path shards over images:

  * inference (evaluation.py:498-502): rank r owns a contiguous slice of the batch, runs the whole
    network on it, and the [B, n_classes] logits are optionally all-gathered - no collective on
    the critical path;
  * training (train.py:1441-1460): data parallel - identical replicas, the loss is averaged over
    the global batch, ONE exchange per step: a sum all-reduce of the flat gradient arena, issued
    as a few large contiguous slices in the order the backward produces them.

Everything here is device-agnostic so that the logic is covered by world_size-2 gloo tests on CPU.
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced partition of n_items units: the first n % world ranks get one more."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def owned_pieces(slices: Sequence[tuple[int, int]], world: int, align: int = 64) -> list[list[tuple[int, int]]]:
    """Ownership map of the peer-memory optimizer (csrc/peer_optim.cu): every slice [a, b) of the
    flat arena is cut into `world` consecutive pieces whose inner boundaries are multiples of `align`
    elements past a; owned[r][k] = rank r's piece of slice k (possibly empty).  The pieces of a slice
    are disjoint and cover it exactly."""
    if world <= 0:
        raise ValueError(f"bad world size {world}")
    owned: list[list[tuple[int, int]]] = [[] for _ in range(world)]
    for a, b in slices:
        per = ((b - a) // world + align - 1) // align * align
        for r in range(world):
            lo = min(a + r * per, b)
            hi = b if r == world - 1 else min(a + (r + 1) * per, b)
            owned[r].append((lo, hi))
    return owned


def allreduce_slices(flat: torch.Tensor, slices: Sequence[tuple[int, int]], group=None,
                     comm_stream=None, ready_events=None, on_slice_done=None) -> None:
    """In-place sum all-reduce of flat[a:b] for every slice, in the given order.  On CUDA the
    collectives are enqueued on `comm_stream` and the current stream waits for them at the end.
    With `ready_events` (one CUDA event per slice, recorded on the compute stream when that slice
    is final) slice k only waits for its own event, so its reduction overlaps whatever the compute
    stream still has queued; without, every collective waits for all work queued so far.
    `on_slice_done(k, a, b)` is called, in order, once the current stream has been made to wait
    for slice k's reduction (and for nothing later): whatever it enqueues - the optimizer update
    of that slice - runs while the remaining slices are still being reduced."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    if ready_events is not None and len(ready_events) != len(slices):
        raise ValueError("ready_events needs one event per slice")
    if flat.is_cuda and comm_stream is not None:
        main = torch.cuda.current_stream(flat.device)
        if ready_events is None:
            comm_stream.wait_stream(main)
        works = []
        with torch.cuda.stream(comm_stream):
            for k, (a, b) in enumerate(slices):
                if ready_events is not None:
                    comm_stream.wait_event(ready_events[k])
                if b > a:
                    works.append((dist.all_reduce(flat[a:b], group=group, async_op=True), k))
        for w, k in works:
            w.wait()                       # the current stream waits for this collective only
            if on_slice_done is not None:
                on_slice_done(k, *slices[k])
        main.wait_stream(comm_stream)
    else:
        for k, (a, b) in enumerate(slices):
            if b > a:
                dist.all_reduce(flat[a:b], group=group)
                if on_slice_done is not None:
                    on_slice_done(k, a, b)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather row shards produced with shard_range(n_total, rank, world) back into [n_total, ...]
    on every rank (shards may differ by one row; they are padded for the collective)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    longest = (n_total + world - 1) // world
    pad = torch.zeros((longest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = []
    for r, part in enumerate(parts):
        a, b = shard_range(n_total, r, world)
        out.append(part[:b - a])
    return torch.cat(out, dim=0)


def sharded_apply(fn: Callable, batch: torch.Tensor, group=None, gather: bool = True):
    """Batch-sharded inference: apply `fn` (e.g. a ViTClassifier, or a ViTObjectDetector returning
    the reference's prediction dict of per-image tensors) to this rank's slice of `batch` and, if
    `gather`, assemble the full result - a tensor or a dict of tensors - on every rank."""
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    a, b = shard_range(batch.shape[0], rank, world)
    local = fn(batch[a:b]) if b > a else None
    if not gather or world == 1:
        return local
    if local is None:   # more ranks than images: contribute an empty shard of the right width
        probe = fn(batch[:1])
        local = {k: v[:0] for k, v in probe.items()} if isinstance(probe, dict) else probe[:0]
    if isinstance(local, dict):
        return {k: gather_rows(v, batch.shape[0], group) for k, v in local.items()}
    return gather_rows(local, batch.shape[0], group)
