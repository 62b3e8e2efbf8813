"""Drop-in mirrors of the reference's detection head and detector wrappers:

    ObjectDetectionHead     evaluation.py:160-200 == train.py:691-731
    ViTObjectDetector       evaluation.py:203-241
    DeiTObjectDetector      train.py:798-838 (eval-mode forward)

Constructor arguments, attribute names and `state_dict()` keys are the reference's, so its
checkpoints (`checkpoint['model_state_dict']`, evaluation.py:375-391) load with `load_state_dict`.
`self.decoder` is a real `nn.TransformerDecoder` - used ONLY as the parameter container (same
initialisation and key names as the reference); its `forward` is never called.  The arithmetic is
`vitk_detection_head_forward` (csrc/detection_head.cu): tcgen05 GEMMs with fused bias / ReLU /
residual epilogues, fused attention at head_dim D/8, one K/V projection GEMM of the encoder tokens
for all six layers.

Under autograd (grad enabled and the features or a head parameter require grad) the call goes
through `vitk_detection_head_forward_train` / `vitk_detection_head_backward`: gradients reach every
head parameter and the encoder features (and, through the encoder bridge, the backbone), so the
reference's training loop - model(images) -> SetCriterion -> losses.backward() -> optimizer.step(),
train.py:1441-1460 - runs on this head unchanged.  In train() mode the decoder layers' dropout
(nn.TransformerDecoderLayer(dropout=0.1), train.py:701-707) is applied at its six sites per layer
with masks regenerated from a per-call seed (never stored).
`weighted_cross_entropy` is the device-side SetCriterion.loss_labels (train.py:1220-1239).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import (VitkDecoderLayerGrads, VitkDecoderLayerWeights, VitkDecoderLayerWeightsT,
                   VitkDetectionHeadConfig, VitkDetectionHeadGrads, VitkDetectionHeadWeights,
                   VitkDetectionHeadWeightsT, check, lib)
from .modules import DataEfficientImageTransformer, VisionTransformer


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _align1k(t: torch.Tensor) -> int:
    return (t.data_ptr() + 1023) // 1024 * 1024


class _DetectionHeadFunction(torch.autograd.Function):
    """class_logits, bbox_coords = head(tokens[:, skip:, :]) with gradients for the tokens and every
    head parameter (vitk_detection_head_forward_train / vitk_detection_head_backward)."""

    @staticmethod
    def forward(ctx, tokens, head, skip, *params):
        tokens = tokens.detach().float().contiguous()
        B, N, _ = tokens.shape
        cfg0, (w, _, _) = head._pack()
        # nn.TransformerDecoderLayer's dropout is active in train() mode, as in the reference; the
        # mask seed is drawn from torch's generator so that torch.manual_seed makes runs
        # reproducible (the backward of this call regenerates the same masks)
        cfg = VitkDetectionHeadConfig.from_buffer_copy(cfg0)
        p_drop = float(head.decoder.layers[0].dropout1.p) if head.training else 0.0
        if p_drop > 0.0:
            cfg.dropout_p = p_drop
            cfg.seed = head.__dict__.get("_vitk_forced_seed") or \
                int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
        saved_b, ws_b = C.c_size_t(0), C.c_size_t(0)
        check(lib().vitk_detection_head_train_bytes(C.byref(cfg), B, N, skip, C.byref(saved_b),
                                                    C.byref(ws_b)))
        dev = tokens.device
        # this call's own buffers, kept alive by the autograd context
        saved_t = torch.empty(saved_b.value + 1024, dtype=torch.uint8, device=dev)
        ws_t = torch.empty(ws_b.value + 1024, dtype=torch.uint8, device=dev)
        Q = head.num_queries
        logits = torch.empty((B, Q, cfg.num_outputs), dtype=torch.float32, device=dev)
        boxes = torch.empty((B, Q, 4), dtype=torch.float32, device=dev)
        check(lib().vitk_detection_head_forward_train(
            C.byref(cfg), C.byref(w), tokens.data_ptr(), B, N, skip, logits.data_ptr(),
            boxes.data_ptr(), _align1k(saved_t), saved_b.value, _align1k(ws_t), ws_b.value,
            _stream()))
        ctx.head, ctx.skip, ctx.shape, ctx.cfg = head, skip, (B, N), cfg
        ctx.saved_t, ctx.ws_t, ctx.boxes = saved_t, ws_t, boxes
        ctx.need_tokens = ctx.needs_input_grad[0]
        return logits, boxes

    @staticmethod
    def backward(ctx, d_logits, d_boxes):
        head, (B, N) = ctx.head, ctx.shape
        if ctx.saved_t is None:
            raise _lib.VitkError("backward through the vitk detection head a second time: the saved "
                                 "activations were released (retain_graph is not supported)")
        _, (w, _, _) = head._pack()
        cfg = ctx.cfg
        wt = head._pack_transposed()
        dev = ctx.boxes.device
        Q, D, L = head.num_queries, cfg.embed_dim, cfg.num_layers
        d_logits = (torch.zeros((B, Q, cfg.num_outputs), device=dev) if d_logits is None
                    else d_logits.float().contiguous())
        d_boxes = (torch.zeros((B, Q, 4), device=dev) if d_boxes is None
                   else d_boxes.float().contiguous())
        grads = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in head.named_parameters()}
        kv_w = torch.zeros((L * 2 * D, D), dtype=torch.float32, device=dev)
        kv_b = torch.zeros(L * 2 * D, dtype=torch.float32, device=dev)
        arr = (VitkDecoderLayerGrads * L)()
        for i in range(L):
            a, pre = arr[i], f"decoder.layers.{i}."
            g = lambda n: grads[pre + n].data_ptr()   # noqa: E731
            a.sa_in_w, a.sa_in_b = g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias")
            a.sa_out_w, a.sa_out_b = g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias")
            # rows [0, D) of multihead_attn.in_proj_* are the query projection
            a.ca_q_w, a.ca_q_b = g("multihead_attn.in_proj_weight"), g("multihead_attn.in_proj_bias")
            a.ca_out_w, a.ca_out_b = (g("multihead_attn.out_proj.weight"),
                                      g("multihead_attn.out_proj.bias"))
            a.ff1_w, a.ff1_b = g("linear1.weight"), g("linear1.bias")
            a.ff2_w, a.ff2_b = g("linear2.weight"), g("linear2.bias")
            a.norm1_w, a.norm1_b = g("norm1.weight"), g("norm1.bias")
            a.norm2_w, a.norm2_b = g("norm2.weight"), g("norm2.bias")
            a.norm3_w, a.norm3_b = g("norm3.weight"), g("norm3.bias")
        G = VitkDetectionHeadGrads()
        G.object_queries = grads["object_queries"].data_ptr()
        G.layers = C.cast(arr, C.POINTER(VitkDecoderLayerGrads))
        G.ca_kv_w, G.ca_kv_b = kv_w.data_ptr(), kv_b.data_ptr()
        G.class_w, G.class_b = grads["class_head.weight"].data_ptr(), grads["class_head.bias"].data_ptr()
        G.bbox_w, G.bbox_b = grads["bbox_head.weight"].data_ptr(), grads["bbox_head.bias"].data_ptr()
        d_tokens = torch.empty((B, N, D), dtype=torch.float32, device=dev) if ctx.need_tokens else None
        check(lib().vitk_detection_head_backward(
            C.byref(cfg), C.byref(w), C.byref(wt[0]), C.byref(G), d_logits.data_ptr(),
            d_boxes.data_ptr(), ctx.boxes.data_ptr(), B, N, ctx.skip,
            d_tokens.data_ptr() if d_tokens is not None else None, _align1k(ctx.saved_t),
            _align1k(ctx.ws_t), _stream()))
        # the K / V rows of multihead_attn.in_proj_* were one GEMM for all layers
        for i in range(L):
            pre = f"decoder.layers.{i}.multihead_attn."
            grads[pre + "in_proj_weight"][D:] = kv_w[i * 2 * D:(i + 1) * 2 * D]
            grads[pre + "in_proj_bias"][D:] = kv_b[i * 2 * D:(i + 1) * 2 * D]
        ctx.saved_t = ctx.ws_t = None
        out = [grads[n] if p.requires_grad else None for n, p in head.named_parameters()]
        return (d_tokens, None, None, *out)


class _WeightedCrossEntropy(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, targets, weight):
        lg = logits.detach().float().contiguous()
        C_ = lg.shape[-1]
        rows = lg.numel() // C_
        tg = targets.detach().to(torch.int64).contiguous()
        wv = weight.detach().float().contiguous() if weight is not None else None
        loss = torch.empty((), dtype=torch.float32, device=lg.device)
        sums = torch.empty(2, dtype=torch.float32, device=lg.device)
        dlg = torch.empty_like(lg)
        check(lib().vitk_weighted_cross_entropy(lg.data_ptr(), tg.data_ptr(),
                                                wv.data_ptr() if wv is not None else None, rows, C_,
                                                loss.data_ptr(), sums.data_ptr(), dlg.data_ptr(),
                                                1.0, _stream()))
        ctx.save_for_backward(dlg)
        return loss

    @staticmethod
    def backward(ctx, g):
        (dlg,) = ctx.saved_tensors
        return dlg * g, None, None


def weighted_cross_entropy(logits: torch.Tensor, targets: torch.Tensor,
                           weight: torch.Tensor | None = None) -> torch.Tensor:
    """SetCriterion.loss_labels' loss (train.py:1236):
    `F.cross_entropy(src_logits.transpose(1, 2), target_classes, self.empty_weight)` as
    `weighted_cross_entropy(src_logits, target_classes, self.empty_weight)` - logits [..., C] with
    the classes LAST (no transpose), targets [...] int64, weight [C]; the weighted mean over every
    prediction, differentiable w.r.t. the logits; loss and gradient are computed on the device by
    vitk_weighted_cross_entropy."""
    if not logits.is_cuda:
        raise _lib.VitkError("weighted_cross_entropy: CUDA tensors only (no CPU fallback)")
    return _WeightedCrossEntropy.apply(logits, targets, weight)


class ObjectDetectionHead(nn.Module):
    def __init__(self, embed_dim=768, num_classes=80, num_queries=100):
        super().__init__()
        self.num_classes = num_classes
        self.num_queries = num_queries
        self.object_queries = nn.Parameter(torch.randn(num_queries, embed_dim))
        decoder_layer = nn.TransformerDecoderLayer(d_model=embed_dim, nhead=8, dim_feedforward=2048,
                                                   dropout=0.1, batch_first=True)
        self.decoder = nn.TransformerDecoder(decoder_layer, num_layers=6)
        self.class_head = nn.Linear(embed_dim, num_classes + 1)
        self.bbox_head = nn.Linear(embed_dim, 4)

    # ------------------------------------------------------------------ weights
    def _pack(self):
        ps = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in ps)
        st = self.__dict__.get("_vitk_pack")
        if st is not None and st[0] == key:
            return st[1], st[2]
        if self.object_queries.device.type != "cuda":
            raise _lib.VitkError("the vitk detection head runs on CUDA only - call .to('cuda') "
                                 "first (no CPU fallback)")
        keep = []

        def f32(t):
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def bf16(t):
            s = ops.cast_bf16(t.detach().float().contiguous())
            keep.append(s)
            return s.data_ptr()

        layers = self.decoder.layers
        L = len(layers)
        D = self.object_queries.shape[1]
        arr = (VitkDecoderLayerWeights * L)()
        kv_w, kv_b = [], []
        for i, ly in enumerate(layers):
            a = arr[i]
            sa, ca = ly.self_attn, ly.multihead_attn
            a.sa_in_w, a.sa_in_b = bf16(sa.in_proj_weight), f32(sa.in_proj_bias)
            a.sa_out_w, a.sa_out_b = bf16(sa.out_proj.weight), f32(sa.out_proj.bias)
            a.ca_q_w, a.ca_q_b = bf16(ca.in_proj_weight[:D]), f32(ca.in_proj_bias[:D])
            a.ca_out_w, a.ca_out_b = bf16(ca.out_proj.weight), f32(ca.out_proj.bias)
            a.ff1_w, a.ff1_b = bf16(ly.linear1.weight), f32(ly.linear1.bias)
            a.ff2_w, a.ff2_b = bf16(ly.linear2.weight), f32(ly.linear2.bias)
            a.norm1_w, a.norm1_b = f32(ly.norm1.weight), f32(ly.norm1.bias)
            a.norm2_w, a.norm2_b = f32(ly.norm2.weight), f32(ly.norm2.bias)
            a.norm3_w, a.norm3_b = f32(ly.norm3.weight), f32(ly.norm3.bias)
            kv_w.append(ca.in_proj_weight.detach()[D:])
            kv_b.append(ca.in_proj_bias.detach()[D:])
        w = VitkDetectionHeadWeights()
        w.object_queries = f32(self.object_queries)
        w.layers = C.cast(arr, C.POINTER(VitkDecoderLayerWeights))
        w.ca_kv_w = bf16(torch.cat(kv_w, dim=0))        # [L*2D, D]: K rows then V rows per layer
        w.ca_kv_b = f32(torch.cat(kv_b, dim=0))
        w.class_w, w.class_b = f32(self.class_head.weight), f32(self.class_head.bias)
        w.bbox_w, w.bbox_b = f32(self.bbox_head.weight), f32(self.bbox_head.bias)
        ly0 = layers[0]
        cfg = VitkDetectionHeadConfig(
            embed_dim=D, num_heads=ly0.self_attn.num_heads, ffn_dim=ly0.linear1.out_features,
            num_layers=L, num_queries=self.num_queries, num_outputs=self.class_head.out_features,
            ln_eps=ly0.norm1.eps)
        self.__dict__["_vitk_pack"] = (key, cfg, (w, arr, keep))
        return cfg, (w, arr, keep)

    def _pack_transposed(self):
        """W^T (bf16 [in, out]) of every matrix weight, for the input-gradient GEMMs; cached like
        `_pack` and made by the library's batched transpose kernel."""
        cfg, (w, arr, keep) = self._pack()
        key = self.__dict__["_vitk_pack"][0]
        st = self.__dict__.get("_vitk_pack_t")
        if st is not None and st[0] == key:
            return st[1]
        D, F, L = cfg.embed_dim, cfg.ffn_dim, cfg.num_layers
        dev = self.object_queries.device
        tarr = (VitkDecoderLayerWeightsT * L)()
        jobs, tkeep = [], []

        def transposed(src_ptr, rows, cols):
            t = torch.empty((cols, rows), dtype=torch.bfloat16, device=dev)
            tkeep.append(t)
            jobs.append((src_ptr, t.data_ptr(), rows, cols))
            return t.data_ptr()

        for i in range(L):
            a, t = arr[i], tarr[i]
            t.sa_in_wt = transposed(a.sa_in_w, 3 * D, D)
            t.sa_out_wt = transposed(a.sa_out_w, D, D)
            t.ca_q_wt = transposed(a.ca_q_w, D, D)
            t.ca_out_wt = transposed(a.ca_out_w, D, D)
            t.ff1_wt = transposed(a.ff1_w, F, D)
            t.ff2_wt = transposed(a.ff2_w, D, F)
        wt = VitkDetectionHeadWeightsT()
        wt.layers = C.cast(tarr, C.POINTER(VitkDecoderLayerWeightsT))
        wt.ca_kv_wt = transposed(w.ca_kv_w, L * 2 * D, D)
        n = len(jobs)
        check(lib().vitk_transpose_bf16_batched(
            n, (C.c_void_p * n)(*[j[0] for j in jobs]), (C.c_void_p * n)(*[j[1] for j in jobs]),
            (C.c_int * n)(*[j[2] for j in jobs]), (C.c_int * n)(*[j[3] for j in jobs]), _stream()))
        out = (wt, tarr, tkeep)
        self.__dict__["_vitk_pack_t"] = (key, out)
        return out

    # ------------------------------------------------------------------ the call
    def decode(self, tokens: torch.Tensor, skip_tokens: int = 0):
        """tokens f32 [B, N, D] (CUDA); the first `skip_tokens` rows of every image are not part
        of the memory (`features[:, 1:, :]`, evaluation.py:235, without the copy).  With autograd
        recording (grad enabled, and the features or a parameter requiring grad) the differentiable
        path runs (module docstring)."""
        if not tokens.is_cuda:
            raise _lib.VitkError("encoder features must be a CUDA tensor (no CPU fallback)")
        if tokens.dim() != 3 or tokens.shape[2] != self.object_queries.shape[1]:
            raise _lib.VitkError(f"expected encoder features [B, N, {self.object_queries.shape[1]}], "
                                 f"got {tuple(tokens.shape)}")
        if torch.is_grad_enabled() and (tokens.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            params = list(self.parameters())
            logits, boxes = _DetectionHeadFunction.apply(tokens, self, skip_tokens, *params)
            return {"class_logits": logits, "bbox_coords": boxes}
        tokens = tokens.detach().float().contiguous()
        B, N, _ = tokens.shape
        cfg, (w, _, _) = self._pack()
        need = C.c_size_t(0)
        check(lib().vitk_detection_head_workspace_bytes(C.byref(cfg), B, N, skip_tokens,
                                                        C.byref(need)))
        ws = self.__dict__.get("_vitk_ws")
        if ws is None or ws.numel() < need.value + 1024 or ws.device != tokens.device:
            ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=tokens.device)
            self.__dict__["_vitk_ws"] = ws
        base = (ws.data_ptr() + 1023) // 1024 * 1024
        Q = self.num_queries
        logits = torch.empty((B, Q, cfg.num_outputs), dtype=torch.float32, device=tokens.device)
        boxes = torch.empty((B, Q, 4), dtype=torch.float32, device=tokens.device)
        check(lib().vitk_detection_head_forward(
            C.byref(cfg), C.byref(w), tokens.data_ptr(), B, N, skip_tokens, logits.data_ptr(),
            boxes.data_ptr(), base, need.value, torch.cuda.current_stream().cuda_stream))
        return {"class_logits": logits, "bbox_coords": boxes}

    def forward(self, encoder_features):
        return self.decode(encoder_features, 0)


class ViTObjectDetector(nn.Module):
    """evaluation.py:203-241."""

    def __init__(self, image_size=224, patch_size=16, in_channels=3, embed_dim=768,
                 num_layers=12, num_heads=12, mlp_dim=3072, dropout=0.1,
                 num_classes=80, num_queries=100):
        super().__init__()
        self.backbone = VisionTransformer(image_size=image_size, patch_size=patch_size,
                                          in_channels=in_channels, embed_dim=embed_dim,
                                          num_layers=num_layers, num_heads=num_heads,
                                          mlp_dim=mlp_dim, dropout=dropout)
        self.detection_head = ObjectDetectionHead(embed_dim=embed_dim, num_classes=num_classes,
                                                  num_queries=num_queries)
        self.num_classes = num_classes
        self.num_queries = num_queries

    def forward(self, images):
        features = self.backbone(images)                      # [B, P + 1, D]
        return self.detection_head.decode(features, 1)       # memory = features[:, 1:, :]


class DeiTObjectDetector(nn.Module):
    """train.py:798-849, eval-mode forward; `return_features=True` also returns the L2-normalised
    triplet projection of the CLS token, as the reference does (train.py:847-848)."""

    def __init__(self, image_size=224, patch_size=16, in_channels=3, embed_dim=768,
                 num_layers=12, num_heads=12, mlp_dim=3072, dropout=0.1,
                 num_classes=80, num_queries=100):
        super().__init__()
        self.backbone = DataEfficientImageTransformer(
            image_size=image_size, patch_size=patch_size, in_channels=in_channels,
            embed_dim=embed_dim, num_layers=num_layers, num_heads=num_heads, mlp_dim=mlp_dim,
            dropout=dropout)
        self.detection_head = ObjectDetectionHead(embed_dim=embed_dim, num_classes=num_classes,
                                                  num_queries=num_queries)
        self.num_classes = num_classes
        self.num_queries = num_queries
        self.triplet_projection = nn.Linear(embed_dim, 256)   # state_dict compatibility

    def forward(self, images, return_features=False):
        features = self.backbone(images)                      # [B, P + 2, D]
        predictions = self.detection_head.decode(features, 2)  # drop CLS and DIST (train.py:842)
        if not (return_features or self.training):             # train.py:835,847
            return predictions
        if torch.is_grad_enabled() and features.requires_grad:
            # differentiable: the row-wise linear kernels under autograd, then F.normalize
            from .autograd import _HeadFunction
            tp = self.triplet_projection
            triplet = torch.nn.functional.normalize(
                _HeadFunction.apply(features[:, 0], tp.weight, tp.bias), p=2, dim=1)
            return predictions, triplet
        # triplet_projection on the CLS row + F.normalize (train.py:833-838), read in place
        feats = features.detach().float().contiguous()
        B, N, D = feats.shape
        tp = self.triplet_projection
        w = tp.weight.detach().float().contiguous()
        b = tp.bias.detach().float().contiguous()
        triplet = torch.empty((B, tp.out_features), dtype=torch.float32, device=feats.device)
        check(lib().vitk_linear_rows(feats.data_ptr(), N * D, w.data_ptr(), b.data_ptr(),
                                     triplet.data_ptr(), B, D, tp.out_features, 1,
                                     torch.cuda.current_stream().cuda_stream))
        return predictions, triplet


@torch.no_grad()
def post_process_predictions(outputs, confidence_threshold=0.5, nms_threshold=0.5):
    """Drop-in for evaluation.py:393-426: same arguments (nms_threshold is unused there too), same
    return value - one dict per image with 'boxes' [k,4], 'labels' [k] (int64), 'scores' [k] of
    the queries whose best non-background probability exceeds the threshold, in query order; an
    image without detections gets the reference's empty CPU tensors.  One kernel launch and ONE
    host synchronisation per batch instead of a softmax / max / mask / `.sum() > 0` round trip per
    image."""
    logits, boxes = outputs["class_logits"], outputs["bbox_coords"]
    if not logits.is_cuda:
        raise _lib.VitkError("post_process_predictions: outputs must be CUDA tensors (no CPU fallback)")
    logits = logits.detach().float().contiguous()
    boxes = boxes.detach().float().contiguous()
    B, Q, n_out = logits.shape
    dev = logits.device
    counts = torch.empty(B, dtype=torch.int32, device=dev)
    boxes_out = torch.empty((B, Q, 4), dtype=torch.float32, device=dev)
    labels_out = torch.empty((B, Q), dtype=torch.int64, device=dev)
    scores_out = torch.empty((B, Q), dtype=torch.float32, device=dev)
    check(lib().vitk_postprocess_detections(
        logits.data_ptr(), boxes.data_ptr(), B, Q, n_out, float(confidence_threshold),
        counts.data_ptr(), boxes_out.data_ptr(), labels_out.data_ptr(), scores_out.data_ptr(),
        torch.cuda.current_stream().cuda_stream))
    result = []
    for i, k in enumerate(counts.tolist()):          # the one synchronisation
        if k > 0:
            result.append({"boxes": boxes_out[i, :k], "labels": labels_out[i, :k],
                           "scores": scores_out[i, :k]})
        else:                                         # evaluation.py:419-424
            result.append({"boxes": torch.zeros((0, 4)),
                           "labels": torch.zeros((0,), dtype=torch.long),
                           "scores": torch.zeros((0,))})
    return result
