"""Host side of the training step (reference train.py:1425-1479, lines 1441-1460).

`TrainState` re-homes a module's parameters into one flat fp32 arena (plus flat gradient, Adam
moment, bf16-shadow and transposed-shadow arenas) so that

  * the optimizer is ONE kernel launch over 85.8 M parameters (vitk_adamw_step),
  * data-parallel gradient reduction is a handful of NCCL all-reduces over contiguous slices,
    issued on a side stream as soon as the backward has produced each slice,
  * every kernel reads weights through stable pointers (VitkWeights / VitkWeightsT / VitkGrads).

`FineTuner` is the fused 6-class fine-tune step of north_star: forward (saving activations) ->
cross-entropy on the CLS head -> backward -> gradient all-reduce -> AdamW.  PyTorch supplies
memory, streams and torch.distributed; all arithmetic runs in libvitk.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib
from ._lib import (VitkBlockGrads, VitkBlockWeights, VitkBlockWeightsT, VitkConfig, VitkGrads,
                   VitkPeerBuffers, VitkWeights, VitkWeightsT, check, lib)

_ALIGN = 64  # elements: keeps every parameter 256-byte (fp32) / 128-byte (bf16) aligned


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PeerArenas:
    """The flat parameter / gradient / bf16-shadow arenas (and the optimizer's guard flag) in
    symmetric memory: the same allocation on every GPU of the group, mapped into every process
    (torch.distributed._symmetric_memory; with NVSwitch also behind one multicast address).  What
    `vitk_peer_reduce_scan` / `vitk_peer_adamw_broadcast` (csrc/peer_optim.cu) read and write."""

    def __init__(self, numel: int, device, group):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        name = self.group.group_name
        self.flat = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.grad = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.shadow = symm_mem.empty(numel, dtype=torch.bfloat16, device=device)
        self.guard = symm_mem.empty(4, dtype=torch.int32, device=device)
        for t in (self.flat, self.grad, self.shadow, self.guard):
            t.zero_()
        self.handles = [symm_mem.rendezvous(t, name) for t in
                        (self.flat, self.grad, self.shadow, self.guard)]
        h = self.handles[0]
        self.world, self.rank = h.world_size, h.rank
        if self.world > _lib.MAX_PEERS:
            raise _lib.VitkError(f"peer optimizer: at most {_lib.MAX_PEERS} ranks")
        self.numel = numel
        self.set_buckets([(0, numel)])

    def set_buckets(self, slices) -> None:
        """Ownership: EVERY gradient bucket (a contiguous slice of the flat arena that the backward
        completes at one point) is cut into `world` pieces with 64-element aligned boundaries, rank
        r owning piece r - so that the reduction of a bucket, started as soon as the backward has
        produced it, is spread over all ranks and links.  owned[r][k] = (lo, hi) of bucket k."""
        from .dist import owned_pieces
        self.buckets = list(slices)
        self.owned = owned_pieces(self.buckets, self.world, _ALIGN)

    def buffers(self, with_guard: bool, multicast: bool = True) -> VitkPeerBuffers:
        pb = VitkPeerBuffers()
        pb.world, pb.rank = self.world, self.rank
        hf, hg, hs, hq = self.handles
        for r in range(self.world):
            pb.param[r], pb.grad[r], pb.shadow[r] = hf.buffer_ptrs[r], hg.buffer_ptrs[r], hs.buffer_ptrs[r]
            pb.guard[r] = hq.buffer_ptrs[r] if with_guard else None
        self.multicast = bool(multicast) and all(int(h.multicast_ptr or 0) != 0 for h in (hf, hg, hs))
        if self.multicast:
            pb.param_mc, pb.grad_mc, pb.shadow_mc = hf.multicast_ptr, hg.multicast_ptr, hs.multicast_ptr
        return pb

    def barrier(self, channel: int = 0) -> None:
        """Cross-rank barrier enqueued on the current stream (signal pads of the parameter arena)."""
        # (the timeout turns a rank that never arrives into a device-side trap - an error on the
        # host - instead of a hang)
        self.handles[0].barrier(channel=channel, timeout_ms=120000)


class TrainState:
    def __init__(self, backbone: torch.nn.Module, head: torch.nn.Linear | None, n_prefix: int,
                 bind_grads: bool = True, peer_group=None, use_peer_memory: bool = False):
        self.backbone, self.head, self.n_prefix = backbone, head, n_prefix
        named = [("backbone." + n, p) for n, p in backbone.named_parameters()]
        if head is not None:
            named += [("head." + n, p) for n, p in head.named_parameters()]
        dev = named[0][1].device
        if dev.type != "cuda":
            raise _lib.VitkError("training runs on CUDA only - call .to('cuda') first")
        self.device = dev
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        self.offsets, off = {}, 0
        for n, p in named:
            self.offsets[n] = off
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = off
        z = lambda dt: torch.zeros(self.numel, dtype=dt, device=dev)
        self.peer = PeerArenas(self.numel, dev, peer_group) if use_peer_memory else None
        if self.peer is not None:
            self.flat, self.grad = self.peer.flat, self.peer.grad
        else:
            self.flat, self.grad = z(torch.float32), z(torch.float32)
        # Adam moments only where the fused optimizer runs (FineTuner); the autograd bridge hands
        # gradients to whatever torch.optim optimizer the caller uses
        self.exp_avg, self.exp_avg_sq = (z(torch.float32), z(torch.float32)) if bind_grads \
            else (None, None)
        self.shadow = self.peer.shadow if self.peer is not None else z(torch.bfloat16)
        for n, p in named:                      # re-home parameters into the arena
            o = self.offsets[n]
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.detach().contiguous())
            p.data = view
            if bind_grads:   # FineTuner: .grad aliases the arena; autograd bridge: autograd owns .grad
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
        # transposed bf16 copies of the four Linear weights of every block
        self._t_jobs = []
        toff = 0
        for i, blk in enumerate(backbone.transformer_blocks):
            for key in ("attention.qkv", "attention.projection", "mlp.linear1", "mlp.linear2"):
                name = f"backbone.transformer_blocks.{i}.{key}.weight"
                w = dict(named)[name]
                self._t_jobs.append((name, w.shape[0], w.shape[1], toff))
                toff += w.numel()
        self.shadow_t = torch.zeros(toff, dtype=torch.bfloat16, device=dev)
        self._build_structs()
        self._bufs_batch = None
        self.step_count = 0
        self.refresh_shadows()

    # ------------------------------------------------------------------ pointer structs
    def _p(self, name):   # fp32 parameter
        return self.flat.data_ptr() + 4 * self.offsets[name]

    def _s(self, name):   # bf16 shadow
        return self.shadow.data_ptr() + 2 * self.offsets[name]

    def _g(self, name):   # fp32 gradient
        return self.grad.data_ptr() + 4 * self.offsets[name]

    def _build_structs(self):
        bb = self.backbone
        L = len(bb.transformer_blocks)
        self._bw = (VitkBlockWeights * L)()
        self._bt = (VitkBlockWeightsT * L)()
        self._bg = (VitkBlockGrads * L)()
        tptr = {n: self.shadow_t.data_ptr() + 2 * o for n, _, _, o in self._t_jobs}
        for i in range(L):
            pre = f"backbone.transformer_blocks.{i}."
            w, t, g = self._bw[i], self._bt[i], self._bg[i]
            for field, key in (("ln1", "layer_norm1"), ("ln2", "layer_norm2")):
                setattr(w, field + "_w", self._p(pre + key + ".weight"))
                setattr(w, field + "_b", self._p(pre + key + ".bias"))
                setattr(g, field + "_w", self._g(pre + key + ".weight"))
                setattr(g, field + "_b", self._g(pre + key + ".bias"))
            for field, key in (("qkv", "attention.qkv"), ("proj", "attention.projection"),
                               ("fc1", "mlp.linear1"), ("fc2", "mlp.linear2")):
                setattr(w, field + "_w", self._s(pre + key + ".weight"))
                setattr(w, field + "_b", self._p(pre + key + ".bias"))
                setattr(g, field + "_w", self._g(pre + key + ".weight"))
                setattr(g, field + "_b", self._g(pre + key + ".bias"))
                setattr(t, field + "_wt", tptr[pre + key + ".weight"])
        W, G = VitkWeights(), VitkGrads()
        W.patch_w = self._s("backbone.patch_embedding.projection.weight")
        W.patch_b = self._p("backbone.patch_embedding.projection.bias")
        G.patch_w = self._g("backbone.patch_embedding.projection.weight")
        G.patch_b = self._g("backbone.patch_embedding.projection.bias")
        W.cls_token, G.cls_token = self._p("backbone.cls_token"), self._g("backbone.cls_token")
        if self.n_prefix == 2:
            W.dist_token, G.dist_token = self._p("backbone.dist_token"), self._g("backbone.dist_token")
        W.pos_embed = self._p("backbone.position_embedding")
        G.pos_embed = self._g("backbone.position_embedding")
        W.blocks = C.cast(self._bw, C.POINTER(VitkBlockWeights))
        G.blocks = C.cast(self._bg, C.POINTER(VitkBlockGrads))
        W.ln_f_w, W.ln_f_b = self._p("backbone.layer_norm.weight"), self._p("backbone.layer_norm.bias")
        G.ln_f_w, G.ln_f_b = self._g("backbone.layer_norm.weight"), self._g("backbone.layer_norm.bias")
        if self.head is not None:
            W.head_w, W.head_b = self._p("head.weight"), self._p("head.bias")
            G.head_w, G.head_b = self._g("head.weight"), self._g("head.bias")
        T = VitkWeightsT()
        T.blocks = C.cast(self._bt, C.POINTER(VitkBlockWeightsT))
        self.W, self.G, self.T = W, G, T
        n = len(self._t_jobs)
        self._t_src = (C.c_void_p * n)(*[self._s(name) for name, _, _, _ in self._t_jobs])
        self._t_dst = (C.c_void_p * n)(*[tptr[name] for name, _, _, _ in self._t_jobs])
        self._t_rows = (C.c_int * n)(*[r for _, r, _, _ in self._t_jobs])
        self._t_cols = (C.c_int * n)(*[c for _, _, c, _ in self._t_jobs])

    def config(self, training: bool = False, seed: int = 0) -> VitkConfig:
        """training=True turns the module's dropout probability on (nn.Dropout is the identity in
        eval mode); `seed` selects the dropout masks and must be the same for the forward and the
        backward of one step."""
        m = self.backbone
        pe, blk = m.patch_embedding, m.transformer_blocks[0]
        return VitkConfig(
            image_size=pe.image_size, patch_size=pe.patch_size,
            in_channels=pe.projection.in_channels, embed_dim=pe.projection.out_channels,
            num_layers=len(m.transformer_blocks), num_heads=blk.attention.num_heads,
            mlp_dim=blk.mlp.linear1.out_features, n_prefix_tokens=self.n_prefix,
            n_classes=self.head.out_features if self.head is not None else 0, precision=0,
            ln_eps=m.layer_norm.eps, dropout_p=float(m.dropout.p) if training else 0.0,
            seed=int(seed) & 0x7FFFFFFF)

    # ------------------------------------------------------------------ shadows
    def refresh_transposes(self):
        check(lib().vitk_transpose_bf16_batched(len(self._t_jobs), self._t_src, self._t_dst,
                                                self._t_rows, self._t_cols, _stream()))

    def refresh_shadows(self):
        """bf16 shadow (and W^T copies) from the fp32 arena - after an external parameter update."""
        check(lib().vitk_cast_f32_to_bf16(self.flat.data_ptr(), self.shadow.data_ptr(), self.numel,
                                          _stream()))
        self.refresh_transposes()
        self._versions = [p._version for p in self.params]

    def owns_params(self) -> bool:
        """False once somebody else re-homed or replaced the module's parameters."""
        lo, hi = self.flat.data_ptr(), self.flat.data_ptr() + 4 * self.numel
        return all(lo <= p.data_ptr() < hi for p in self.params)

    def shadows_stale(self) -> bool:
        return any(p._version != v for p, v in zip(self.params, self._versions))

    # ------------------------------------------------------------------ buffers
    def buffer_sizes(self, batch: int) -> tuple[int, int]:
        cfg = self.config()
        sv, ws = C.c_size_t(0), C.c_size_t(0)
        check(lib().vitk_train_workspace_bytes(C.byref(cfg), batch, C.byref(sv), C.byref(ws)))
        return sv.value, ws.value

    def buffers(self, batch: int):
        """The step-private activation / workspace buffers of FineTuner (one forward in flight)."""
        if self._bufs_batch != batch:
            self._sizes = self.buffer_sizes(batch)
            self._saved = torch.empty(self._sizes[0] + 1024, dtype=torch.uint8, device=self.device)
            self._ws = torch.empty(self._sizes[1] + 1024, dtype=torch.uint8, device=self.device)
            self._bufs_batch = batch
        al = lambda t: (t.data_ptr() + 1023) // 1024 * 1024
        return al(self._saved), self._sizes[0], al(self._ws), self._sizes[1]

    def new_buffers(self, batch: int):
        """Fresh activation / workspace tensors for ONE forward of the autograd bridge: they live on
        that call's autograd context, so a second forward before the first backward (siamese /
        triplet pairs, summed micro-batches, other batch sizes) cannot overwrite them.  The
        caching allocator hands the same blocks back once the backward has released them."""
        sv, ws = self.buffer_sizes(batch)
        return (torch.empty(sv + 1024, dtype=torch.uint8, device=self.device), sv,
                torch.empty(ws + 1024, dtype=torch.uint8, device=self.device), ws)

    # ------------------------------------------------------------------ gradient slices
    def bucket_slices(self):
        """Contiguous gradient slices in the order the backward completes them:
        [final LN + head], block L-1, ..., block 0, [tokens + position + patch embedding]."""
        L = len(self.backbone.transformer_blocks)
        first_block = self.offsets["backbone.transformer_blocks.0.attention.qkv.weight"]
        # named_parameters order inside a block starts with attention.qkv.weight
        starts = [min(o for n, o in self.offsets.items()
                      if n.startswith(f"backbone.transformer_blocks.{i}.")) for i in range(L)]
        tail = self.offsets["backbone.layer_norm.weight"]
        out = [(tail, self.numel)]
        for i in reversed(range(L)):
            end = starts[i + 1] if i + 1 < L else tail
            out.append((starts[i], end))
        out.append((0, first_block if first_block == starts[0] else starts[0]))
        return out


class FineTuner:
    """Fused fine-tune step: AdamW(lr=1e-4, weight_decay=1e-4) as train.py:1598-1602."""

    def __init__(self, model, lr=1e-4, weight_decay=1e-4, betas=(0.9, 0.999), eps=1e-8,
                 process_group=None, overlap_allreduce: bool = True, reserve_sms: int = 0,
                 seed: int = 0, data_parallel: bool = True, skip_nonfinite: bool = True,
                 grad_sync: str = "auto"):
        """overlap_allreduce: start each gradient bucket's all-reduce on a side stream as soon as
        the backward has produced it (events recorded by vitk_classifier_loss_backward_ev), instead
        of reducing everything after the backward.  reserve_sms: SMs kept out of the persistent
        kernels' grids during the backward so that the NCCL kernels do not have to displace them
        (pair with a communicator limited to as many CTAs: ProcessGroupNCCL.Options.config.max_ctas)."""
        """grad_sync (data parallel only): "nccl" = all-reduce of the flat gradient slices, then
        the same AdamW on every rank; "peer" = sharded reduce + AdamW + broadcast through symmetric
        memory (csrc/peer_optim.cu: multimem.ld_reduce / multimem.st over NVSwitch when the fabric
        has multicast, peer loads / stores otherwise; Adam moments sharded over the ranks); "auto"
        (default; VITK_GRAD_SYNC overrides) = "peer" when every rank can set it up, else "nccl"."""
        import os
        self.model = model
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.pg = process_group
        # data_parallel=False: a single-process step even inside an initialised process group
        # (bench.py's dp_check compares it with the data-parallel step on the same weights)
        ddp = data_parallel and dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(process_group) if ddp else 1
        rank = dist.get_rank(process_group) if ddp else 0
        grad_sync = os.environ.get("VITK_GRAD_SYNC", grad_sync)
        if grad_sync not in ("auto", "nccl", "peer"):
            raise _lib.VitkError(f"grad_sync must be 'auto', 'nccl' or 'peer' (got {grad_sync!r})")
        self.state, self.peer, self.grad_sync = None, None, "none" if self.world == 1 else "nccl"
        if self.world > 1 and grad_sync in ("auto", "peer"):
            self.state, why = self._try_peer_state(model, process_group, rank)
            if self.state is not None:
                self.peer, self.grad_sync = self.state.peer, "peer"
            elif grad_sync == "peer":
                raise _lib.VitkError(f"grad_sync='peer' is not available here: {why}")
            elif rank == 0:
                import sys
                print(f"vitk: gradient exchange over NCCL (peer-memory path unavailable: {why})",
                      file=sys.stderr)
        if self.state is None:
            self.state = TrainState(model.backbone, model.head, model.backbone._n_prefix)
        # dropout mask stream: step k uses seed + k; every data-parallel rank draws its own masks
        # (the same masks on every shard would correlate the ranks' gradients)
        self.seed = (int(seed) + 0x9E3779B1 * rank) & 0x7FFFFFFF
        self._ranges_key, self._ranges = None, None
        # skip_nonfinite: the reference's scaler.step(optimizer) (train.py:1456) skips the update
        # when a gradient is inf / nan; device int[2] {flag, steps skipped}, never read by the
        # host inside step().  Off: every gradient slice is updated as soon as its own all-reduce
        # has finished (the optimizer then runs under the reductions still in flight).
        if not skip_nonfinite:
            self._guard = None
        elif self.peer is not None:
            self._guard = self.peer.guard[:2]       # symmetric: a rank raises the flag on all ranks
        else:
            self._guard = torch.zeros(2, dtype=torch.int32, device=self.state.device)
        self._peer_buffers = self.peer.buffers(skip_nonfinite) if self.peer is not None else None
        self._comm_stream = (torch.cuda.Stream(device=self.state.device, priority=-1)
                             if self.world > 1 else None)
        self._loss = torch.zeros(1, dtype=torch.float32, device=self.state.device)
        self.overlap = bool(overlap_allreduce) and self.world > 1
        if self.peer is not None:
            # pass A bucket by bucket under the backward: measured equal to reducing after the
            # backward at 8 GPUs (20.63 against 20.53 ms per step: per-bucket barriers against
            # 0.3 ms of hidden NVLink time), so it is opt-in
            self.overlap = self.overlap and os.environ.get("VITK_PEER_OVERLAP", "0") == "1"
        self.reserve_sms = int(reserve_sms) if self.overlap else 0
        self._slices = self.state.bucket_slices()
        if self.peer is not None:
            self.peer.set_buckets(self._slices if self.overlap else [(0, self.state.numel)])
        self._events, self._event_arr = None, None
        if self.overlap:
            # one event per gradient bucket; torch creates the cudaEvent lazily at the first record
            self._events = [torch.cuda.Event() for _ in self._slices]
            for ev in self._events:
                ev.record()
            self._event_arr = (C.c_void_p * len(self._events))(*[ev.cuda_event for ev in self._events])

    @torch.no_grad()
    def step(self, images: torch.Tensor, labels: torch.Tensor):
        """One optimisation step on this rank's shard; returns (loss tensor [1], logits [B, C]).
        The loss is this rank's contribution to the global-batch mean."""
        st = self.state
        if not images.is_cuda or not labels.is_cuda:
            raise _lib.VitkError("images / labels must be CUDA tensors (no CPU fallback)")
        images = images.float().contiguous()
        labels = labels.long().contiguous()
        B = images.shape[0]
        if not st.owns_params():
            raise _lib.VitkError("the model's parameters were re-homed after this FineTuner was "
                                 "built (e.g. .to(), load into new tensors, or a second trainer)")
        if st.shadows_stale():
            st.refresh_shadows()
        # dropout follows the module's train()/eval() state, as nn.Dropout does in the reference
        cfg = st.config(training=self.model.training, seed=self.seed + st.step_count)
        saved, saved_bytes, ws, ws_bytes = st.buffers(B)
        logits = torch.empty((B, cfg.n_classes), dtype=torch.float32, device=st.device)
        st.grad.zero_()
        self._loss.zero_()
        if self.peer is not None and self._guard is not None:
            self._guard[0:1].zero_()   # local flag; peers may raise it again after the barrier below
        s = _stream()
        check(lib().vitk_forward_train(C.byref(cfg), C.byref(st.W), images.data_ptr(), B, None,
                                       saved, saved_bytes, ws, ws_bytes, s))
        if self.overlap:
            if self.reserve_sms:
                check(lib().vitk_reserve_sms(self.reserve_sms))
            check(lib().vitk_classifier_loss_backward_ev(
                C.byref(cfg), C.byref(st.W), C.byref(st.T), C.byref(st.G), labels.data_ptr(), B,
                1.0 / (B * self.world), logits.data_ptr(), self._loss.data_ptr(), saved, ws,
                self._event_arr, len(self._events), s))
            if self.reserve_sms:
                check(lib().vitk_reserve_sms(0))
        else:
            check(lib().vitk_classifier_loss_backward(
                C.byref(cfg), C.byref(st.W), C.byref(st.T), C.byref(st.G), labels.data_ptr(), B,
                1.0 / (B * self.world), logits.data_ptr(), self._loss.data_ptr(), saved, ws, s))
        st.step_count += 1
        # torch.optim.AdamW(model.parameters()) skips parameters without a gradient: frozen ones
        # (requires_grad=False, the usual freeze-the-backbone fine-tune) get neither the Adam
        # update nor the decoupled weight decay.  One launch per maximal run of trainable
        # parameters - a single launch over the whole arena when nothing is frozen.
        ranges = self._trainable_ranges()

        guard = self._guard.data_ptr() if self._guard is not None else None

        def adamw(a: int, b: int):
            for lo, hi in ranges:
                lo, hi = max(lo, a), min(hi, b)
                if hi > lo:
                    check(lib().vitk_adamw_step_guarded(
                        st.flat.data_ptr() + 4 * lo, st.grad.data_ptr() + 4 * lo,
                        st.exp_avg.data_ptr() + 4 * lo, st.exp_avg_sq.data_ptr() + 4 * lo,
                        st.shadow.data_ptr() + 2 * lo, hi - lo, self.lr, self.betas[0],
                        self.betas[1], self.eps, self.wd, st.step_count, 1.0, guard, s))

        first = [True]

        def scan(a: int, b: int):     # non-finite check of one (reduced) gradient slice
            if b > a:
                check(lib().vitk_grad_guard_scan(st.grad.data_ptr() + 4 * a, b - a, guard,
                                                 1 if first[0] else 0, s))
                first[0] = False

        if self.peer is not None:
            self._peer_step(ranges, s)
        elif guard is None:
            if self.world > 1:
                self._allreduce_grads(on_slice_done=lambda k, a, b: adamw(a, b))
            else:
                adamw(0, st.numel)
        else:
            # all of the step's gradients are checked before any parameter moves; in the
            # data-parallel case slice by slice as the reductions finish (a non-finite value on
            # one rank reaches every rank through the sum, so all ranks skip together)
            if self.world > 1:
                self._allreduce_grads(on_slice_done=lambda k, a, b: scan(a, b))
            else:
                scan(0, st.numel)
            adamw(0, st.numel)
            check(lib().vitk_grad_guard_finish(guard, s))
        st.refresh_transposes()
        # The update went through raw pointers: bump the parameters' version counters so that
        # everything keyed on them re-reads the weights - the inference engine's packed bf16
        # copies (engine.py: pack) would otherwise serve the pre-step matrices to the next
        # validate / predict.  The shadows of THIS state were refreshed by the kernel itself.
        torch._C._increment_version(st.params)
        st._versions = [p._version for p in st.params]
        return self._loss, logits

    # ------------------------------------------------------------------ peer-memory optimizer step
    def _try_peer_state(self, model, group, rank):
        """TrainState with its arenas in symmetric memory, or (None, reason).  Collective: every
        rank calls it, and all ranks agree on the outcome."""
        state, why = None, ""
        try:
            state = TrainState(model.backbone, model.head, model.backbone._n_prefix,
                               peer_group=group, use_peer_memory=True)
        except Exception as e:   # no symmetric-memory support in this build / on this fabric
            why = f"{type(e).__name__}: {e}"
        ok = torch.tensor([1 if state is not None else 0], device=model.head.weight.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if state is not None:
                # parameters were re-homed into the symmetric arena: the NCCL state re-homes them again
                why = "another rank could not set it up"
            return None, why or "another rank could not set it up"
        return state, ""

    def reduced_grad_ranges(self) -> list[tuple[int, int]]:
        """The pieces of the flat gradient arena that hold the cross-rank SUM after step():
        everything with NCCL, the pieces this rank owns with the peer-memory optimizer."""
        if self.peer is None:
            return [(0, self.state.numel)]
        return [(a, b) for a, b in self.peer.owned[self.peer.rank] if b > a]

    def _peer_step(self, ranges, s):
        """Pass A (reduce + non-finite scan) bucket by bucket on the side stream, each bucket as soon
        as the backward has produced it on EVERY rank (its event, then a cross-rank barrier) - so it
        runs under the rest of the backward; pass B (AdamW + broadcast) once every gradient has been
        checked (train.py:1456: a non-finite gradient anywhere skips the whole step)."""
        st, pr = self.state, self.peer
        pb = C.byref(self._peer_buffers)
        lib_ = lib()

        def clip(a, b):   # trainable parts of an owned piece
            return [(max(a, lo), min(b, hi)) for lo, hi in ranges if min(b, hi) > max(a, lo)]

        main = torch.cuda.current_stream(st.device)
        own = pr.owned[pr.rank]
        if self.overlap:
            side = self._comm_stream
            with torch.cuda.stream(side):
                for k, (a, b) in enumerate(own):
                    side.wait_event(self._events[k])
                    pr.barrier(k % 8)          # bucket k is final on every rank
                    for lo, hi in clip(a, b):
                        check(lib_.vitk_peer_reduce_scan(pb, lo, hi, side.cuda_stream))
            main.wait_stream(side)
        else:
            pr.barrier(0)                      # every rank's backward has finished
            for a, b in own:
                for lo, hi in clip(a, b):
                    check(lib_.vitk_peer_reduce_scan(pb, lo, hi, s))
        pr.barrier(8)                          # every piece reduced: the skip flag is final everywhere
        for a, b in own:
            for lo, hi in clip(a, b):
                check(lib_.vitk_peer_adamw_broadcast(
                    pb, st.exp_avg.data_ptr(), st.exp_avg_sq.data_ptr(), lo, hi, self.lr,
                    self.betas[0], self.betas[1], self.eps, self.wd, st.step_count, 1.0, s))
        if self._guard is not None:
            check(lib_.vitk_grad_guard_finish(self._guard.data_ptr(), s))
        pr.barrier(9)                          # every rank's parameters and shadows are in place

    @property
    def skipped_steps(self) -> int:
        """Optimizer steps skipped because a gradient was inf / nan (synchronises)."""
        return int(self._guard[1].item()) if self._guard is not None else 0

    def _trainable_ranges(self) -> list[tuple[int, int]]:
        st = self.state
        key = tuple(p.requires_grad for p in st.params)
        if key != self._ranges_key:
            out: list[list[int]] = []
            for n, p in zip(st.names, st.params):
                if not p.requires_grad:
                    continue
                lo = st.offsets[n]
                hi = lo + (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
                if out and out[-1][1] == lo:
                    out[-1][1] = hi
                else:
                    out.append([lo, hi])
            self._ranges_key, self._ranges = key, [(a, b) for a, b in out]
        return self._ranges

    # ------------------------------------------------------------------ checkpoints
    # The reference saves {'epoch', 'model_state_dict', 'optimizer_state_dict', 'val_loss',
    # 'config'} (train.py:1647-1666) and evaluation.py:375-391 loads 'model_state_dict'.  The
    # module mirrors keep the reference's state_dict keys; the optimizer state below uses
    # torch.optim.AdamW's own layout (per-parameter 'step', 'exp_avg', 'exp_avg_sq' in
    # model.parameters() order), so a checkpoint written by either implementation resumes in the
    # other.
    def optimizer_state_dict(self) -> dict:
        """torch.optim.AdamW's state layout.  With the peer-memory optimizer the Adam moments are
        sharded over the ranks: this call is then COLLECTIVE (every rank must make it) - each shard
        is broadcast from its owner first."""
        st = self.state
        if self.peer is not None:
            for r in range(self.peer.world):
                src = dist.get_global_rank(self.peer.group, r)
                for a, b in self.peer.owned[r]:
                    if b > a:
                        dist.broadcast(st.exp_avg[a:b], src=src, group=self.peer.group)
                        dist.broadcast(st.exp_avg_sq[a:b], src=src, group=self.peer.group)
        state = {}
        for i, (name, p) in enumerate(zip(st.names, st.params)):
            o = st.offsets[name]
            state[i] = {"step": torch.tensor(float(st.step_count)),
                        "exp_avg": st.exp_avg[o:o + p.numel()].view(p.shape).clone(),
                        "exp_avg_sq": st.exp_avg_sq[o:o + p.numel()].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps,
                 "weight_decay": self.wd, "amsgrad": False, "maximize": False, "foreach": None,
                 "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": True, "params": list(range(len(st.params)))}
        return {"state": state if st.step_count > 0 else {}, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd: dict) -> None:
        st = self.state
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(st.params):
            raise _lib.VitkError("optimizer state does not match this model: expected one parameter "
                                 f"group with {len(st.params)} parameters (train.py:1598-1602)")
        g = groups[0]
        self.lr, self.wd, self.eps = float(g["lr"]), float(g["weight_decay"]), float(g["eps"])
        self.betas = tuple(float(b) for b in g["betas"])
        st.exp_avg.zero_()
        st.exp_avg_sq.zero_()
        steps = set()
        for i, (name, p) in enumerate(zip(st.names, st.params)):
            e = sd["state"].get(g["params"][i])
            if e is None:
                continue
            o = st.offsets[name]
            st.exp_avg[o:o + p.numel()].view(p.shape).copy_(e["exp_avg"])
            st.exp_avg_sq[o:o + p.numel()].view(p.shape).copy_(e["exp_avg_sq"])
            steps.add(int(float(e["step"])))
        if len(steps) > 1:
            raise _lib.VitkError(f"per-parameter step counts differ ({sorted(steps)}): the fused "
                                 "AdamW keeps one step count for the whole arena")
        st.step_count = steps.pop() if steps else 0

    def save_checkpoint(self, path, epoch: int = 0, val_loss: float = float("nan"), config=None):
        """Same dictionary as train.py:1647-1654."""
        torch.save({"epoch": epoch, "model_state_dict": self.model.state_dict(),
                    "optimizer_state_dict": self.optimizer_state_dict(), "val_loss": val_loss,
                    "config": dict(config or {})}, path)

    def load_checkpoint(self, path, strict: bool = False) -> dict:
        """Loads a checkpoint written by save_checkpoint or by the reference's training loop
        (evaluation.py:375-391 semantics: strict=False on the model)."""
        ck = torch.load(path, map_location=self.state.device, weights_only=False)
        sd = ck["model_state_dict"] if "model_state_dict" in ck else ck
        with torch.no_grad():
            self.model.load_state_dict(sd, strict=strict)   # copies in place into the arena
        if isinstance(ck, dict) and "optimizer_state_dict" in ck:
            self.load_optimizer_state_dict(ck["optimizer_state_dict"])
        return ck

    def _allreduce_grads(self, on_slice_done=None):
        """Sum gradients over the data-parallel group: a few large all-reduces over contiguous
        slices of the flat gradient arena on a side stream (NCCL over NVLink/NVSwitch)."""
        from .dist import allreduce_slices
        allreduce_slices(self.state.grad, self._slices, self.pg, self._comm_stream,
                         ready_events=self._events if self.overlap else None,
                         on_slice_done=on_slice_done)
