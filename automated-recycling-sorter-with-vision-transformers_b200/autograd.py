"""Autograd bridge: `features = backbone(images)` with gradients (reference train.py:831 inside the
autocast forward of train.py:1441-1444, differentiated by `losses.backward()` at train.py:1455).

The forward runs `vitk_forward_train` (saves activations), the backward receives
dLoss/d features from PyTorch autograd - whatever loss the caller built on top, e.g. the
reference's own detection head and Hungarian loss - and runs `vitk_backward_tokens`; parameter
gradients are handed back to autograd so that any `torch.optim` optimizer works unchanged.
`FineTuner` (trainer.py) is the fused fast path for the 6-class fine-tune; this bridge is the
general drop-in.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import check, lib
from .trainer import TrainState


def _align1k(t: torch.Tensor) -> int:
    return (t.data_ptr() + 1023) // 1024 * 1024


def _state_for(backbone) -> TrainState:
    st = backbone.__dict__.get("_vitk_train_state")
    if st is None or st.flat.device != backbone.cls_token.device or not st.owns_params():
        st = TrainState(backbone, None, backbone._n_prefix, bind_grads=False)
        backbone.__dict__["_vitk_train_state"] = st
    return st


class _BackboneFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, state, *params):
        if not images.is_cuda:
            raise _lib.VitkError("images must be a CUDA tensor (no CPU fallback)")
        images = images.detach().float().contiguous()
        B = images.shape[0]
        if state.shadows_stale():
            state.refresh_shadows()
        # dropout as in the reference: active in train() mode; the mask seed is drawn from torch's
        # generator so that torch.manual_seed makes runs reproducible
        training = bool(state.backbone.training) and float(state.backbone.dropout.p) > 0.0
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if training else 0
        cfg = state.config(training=training, seed=seed)
        # this call's own activation / workspace buffers, kept alive by the autograd context
        saved_t, saved_bytes, ws_t, ws_bytes = state.new_buffers(B)
        saved, ws = _align1k(saved_t), _align1k(ws_t)
        pe = state.backbone.patch_embedding
        N = pe.n_patches + state.n_prefix
        tokens = torch.empty((B, N, cfg.embed_dim), dtype=torch.float32, device=images.device)
        check(lib().vitk_forward_train(C.byref(cfg), C.byref(state.W), images.data_ptr(), B,
                                       tokens.data_ptr(), saved, saved_bytes, ws, ws_bytes,
                                       torch.cuda.current_stream().cuda_stream))
        ctx.state, ctx.batch, ctx.training, ctx.seed = state, B, training, seed
        ctx.saved_t, ctx.ws_t = saved_t, ws_t
        return tokens

    @staticmethod
    def backward(ctx, d_tokens):
        st, B = ctx.state, ctx.batch
        d_tokens = d_tokens.float().contiguous()
        cfg = st.config(training=ctx.training, seed=ctx.seed)
        if ctx.saved_t is None:
            raise _lib.VitkError("backward through the vitk encoder a second time: the saved "
                                 "activations were released (retain_graph is not supported)")
        saved, ws = _align1k(ctx.saved_t), _align1k(ctx.ws_t)
        st.grad.zero_()
        check(lib().vitk_backward_tokens(C.byref(cfg), C.byref(st.W), C.byref(st.T), C.byref(st.G),
                                         d_tokens.data_ptr(), B, saved, ws,
                                         torch.cuda.current_stream().cuda_stream))
        grads = []
        for name, p in zip(st.names, st.params):
            o = st.offsets[name]
            grads.append(st.grad[o:o + p.numel()].view(p.shape).clone() if p.requires_grad else None)
        ctx.saved_t = ctx.ws_t = None   # release the activations now, not when ctx is collected
        return (None, None, *grads)


def encoder_forward_train(module, images):
    """tokens f32 [B, N, D] = backbone(images), differentiable w.r.t. every backbone parameter."""
    st = _state_for(module)
    return _BackboneFunction.apply(images, st, *st.params)


class _HeadFunction(torch.autograd.Function):
    """Linear(D, n_classes) on the CLS rows: vitk_linear_rows forward, vitk_linear_rows_backward."""

    @staticmethod
    def forward(ctx, cls_rows, weight, bias):
        x = cls_rows.detach().float().contiguous()
        w, b = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        out = torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
        check(lib().vitk_linear_rows(x.data_ptr(), x.shape[1], w.data_ptr(), b.data_ptr(),
                                     out.data_ptr(), x.shape[0], x.shape[1], w.shape[0], 0,
                                     torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.float().contiguous()
        dx, dw = torch.empty_like(x), torch.empty_like(w)
        db = torch.empty(w.shape[0], dtype=torch.float32, device=x.device)
        check(lib().vitk_linear_rows_backward(x.data_ptr(), x.shape[1], w.data_ptr(), dy.data_ptr(),
                                              dx.data_ptr(), dw.data_ptr(), db.data_ptr(),
                                              x.shape[0], x.shape[1], w.shape[0],
                                              torch.cuda.current_stream().cuda_stream))
        return dx, dw, db


def classifier_forward_train(model, images):
    """ViTClassifier under autograd: backbone through the bridge, the head on the CLS rows through
    the library's row-wise linear kernels (forward and backward) - the same composition a
    user-defined head would use on `tokens[:, 0]`."""
    tokens = encoder_forward_train(model.backbone, images)
    return _HeadFunction.apply(tokens[:, 0], model.head.weight, model.head.bias)
