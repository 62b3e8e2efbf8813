"""Host-buffer entry point of the classifier and of the detector: the evaluation loop of the
reference (evaluation.py:498-502 - `images.to(device)`; `model(images)`) with the host->device copy
of batch i+1 overlapped with the kernels of batch i.

    runner = HostBatchRunner(model, batch_size)
    for logits in runner.run(host_batches):      # host_batches: pinned f32 [B,3,S,S] CPU tensors
        ...                                      # logits: f32 [B, n_classes] CPU tensor
    # model = ViTObjectDetector / DeiTObjectDetector: each result is the reference's prediction
    # dict {'class_logits': [B,Q,C+1], 'bbox_coords': [B,Q,4]} as pinned CPU tensors

Two device input slots and a dedicated copy stream; every step moves its images host->device and
its logits device->host.  PyTorch is used for streams, events and memory only.

A batch may hold fewer images than `batch_size` (the reference's evaluation DataLoader keeps its
smaller last batch, evaluation.py:555-562: drop_last is False).  Buffer lifetime: every yielded
result is a view of one of TWO pinned host buffers and is overwritten two steps later - consume it
(or `.clone()` it) before asking for the batch after the next; `list(runner.run(...))` therefore
needs `copy_results=True`.
"""
from __future__ import annotations

from typing import Iterable, Iterator

import torch


class HostBatchRunner:
    def __init__(self, model, batch_size: int, device: torch.device | str = "cuda",
                 input_dtype: torch.dtype = torch.float32, copy_results: bool = False):
        """input_dtype float32: host batches are normalised f32 NCHW tensors, exactly what the
        reference copies to the device (evaluation.py:499).  uint8: host batches are raw decoder
        output, u8 NHWC [B, S, S, 3]; Normalize + ToTensorV2 run on the device (a quarter of the
        host->device bytes; same logits bit for bit)."""
        self.model = model
        self.device = torch.device(device)
        pe = model.backbone.patch_embedding
        if input_dtype == torch.uint8:
            shape = (batch_size, pe.image_size, pe.image_size, pe.projection.in_channels)
        elif input_dtype == torch.float32:
            shape = (batch_size, pe.projection.in_channels, pe.image_size, pe.image_size)
        else:
            raise ValueError("input_dtype must be torch.float32 or torch.uint8")
        self.shape, self.input_dtype = shape, input_dtype
        self.copy_results = bool(copy_results)
        self._count = [0, 0]       # images in each input slot
        self._out_count = [0, 0]   # images in each result slot
        self.detector = hasattr(model, "detection_head")
        if self.detector:
            head = model.detection_head
            out_shapes = {"class_logits": (batch_size, head.num_queries, head.class_head.out_features),
                          "bbox_coords": (batch_size, head.num_queries, 4)}
        else:
            self.n_classes = model.head.out_features
            out_shapes = {"logits": (batch_size, self.n_classes)}
        self._dev_in = [torch.empty(shape, dtype=input_dtype, device=self.device) for _ in range(2)]
        self._host_out = [{k: torch.empty(v, dtype=torch.float32).pin_memory()
                           for k, v in out_shapes.items()} for _ in range(2)]
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._in_ready = [torch.cuda.Event() for _ in range(2)]    # H2D of slot done
        self._in_free = [torch.cuda.Event() for _ in range(2)]     # compute finished reading slot
        self._out_ready = [torch.cuda.Event() for _ in range(2)]   # D2H of slot's logits done
        self.h2d_bytes_per_step = self._dev_in[0].element_size() * self._dev_in[0].numel()
        self.d2h_bytes_per_step = sum(4 * t.numel() for t in self._host_out[0].values())

    def _enqueue_copy(self, slot: int, host: torch.Tensor, first_use: bool):
        n = host.shape[0] if host.dim() == len(self.shape) else -1
        if tuple(host.shape[1:]) != self.shape[1:] or not 0 < n <= self.shape[0] or \
                host.dtype != self.input_dtype or host.is_cuda:
            raise ValueError(f"expected a CPU {self.input_dtype} tensor of shape [1..{self.shape[0]}, "
                             f"{', '.join(map(str, self.shape[1:]))}], got {tuple(host.shape)}")
        self._count[slot] = n
        with torch.cuda.stream(self._copy_stream):
            if not first_use:
                self._copy_stream.wait_event(self._in_free[slot])
            self._dev_in[slot][:n].copy_(host, non_blocking=True)
            self._in_ready[slot].record(self._copy_stream)

    @torch.no_grad()
    def run(self, host_batches: Iterable[torch.Tensor]) -> Iterator[torch.Tensor]:
        main = torch.cuda.current_stream(self.device)
        it = iter(host_batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        used = [False, False]
        self._enqueue_copy(0, nxt, True)
        used[0] = True
        i = 0
        pending = None  # (slot) whose logits are in flight to the host
        while nxt is not None:
            slot = i & 1
            try:
                after = next(it)
            except StopIteration:
                after = None
            if after is not None:
                self._enqueue_copy(slot ^ 1, after, not used[slot ^ 1])
                used[slot ^ 1] = True
            main.wait_event(self._in_ready[slot])
            n = self._count[slot]
            out = self.model(self._dev_in[slot][:n])
            self._in_free[slot].record(main)
            if pending is not None:
                # the previous step's results must have left the pinned buffers before their reuse
                self._out_ready[pending].synchronize()
                yield self._result(pending)
            if not self.detector:
                out = {"logits": out}
            for k, host in self._host_out[slot].items():
                host[:n].copy_(out[k], non_blocking=True)
            self._out_ready[slot].record(main)
            self._out_count[slot] = n
            pending = slot
            nxt = after
            i += 1
        if pending is not None:
            self._out_ready[pending].synchronize()
            yield self._result(pending)

    def _result(self, slot: int):
        n = self._out_count[slot]
        res = {k: (v[:n].clone() if self.copy_results else v[:n])
               for k, v in self._host_out[slot].items()}
        return res if self.detector else res["logits"]
