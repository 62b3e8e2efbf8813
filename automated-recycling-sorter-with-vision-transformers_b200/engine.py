"""Host-side marshalling between an encoder nn.Module and `vitk_forward`.

Packs the module's parameters into the `VitkWeights` pointer struct (fp32 parameters are passed
in place; matrix weights get a bf16 shadow made by the library's own cast kernel), owns the
workspace buffer and issues the single C-ABI call that replaces `self.backbone(images)`
(reference evaluation.py:231, train.py:831).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from ._lib import VitkBlockWeights, VitkConfig, VitkWeights, check, lib


class EncoderEngine:
    def __init__(self, module: torch.nn.Module, n_prefix_tokens: int):
        self.module = module
        self.n_prefix = n_prefix_tokens
        self._pack_key = None
        self._keep = []          # tensors that the pointer structs reference
        self._weights = None
        self._blocks = None
        self._ws = None
        self._ws_bytes = 0
        self.head: torch.nn.Linear | None = None
        self.precision = 0       # 0 = bf16 operands, 1 = fp32-parity (three-term bf16 split GEMMs)

    # ------------------------------------------------------------------ config
    def config(self, n_classes: int = 0) -> VitkConfig:
        m = self.module
        pe = m.patch_embedding
        blk = m.transformer_blocks[0]
        return VitkConfig(
            image_size=pe.image_size, patch_size=pe.patch_size,
            in_channels=pe.projection.in_channels, embed_dim=pe.projection.out_channels,
            num_layers=len(m.transformer_blocks), num_heads=blk.attention.num_heads,
            mlp_dim=blk.mlp.linear1.out_features, n_prefix_tokens=self.n_prefix,
            n_classes=n_classes, precision=self.precision, ln_eps=m.layer_norm.eps,
            dropout_p=float(m.dropout.p), seed=0)

    # ------------------------------------------------------------------ weights
    def _params_for_key(self, head):
        ps = list(self.module.parameters())
        if head is not None:
            ps += list(head.parameters())
        return ps

    def pack(self, head: torch.nn.Linear | None = None) -> VitkWeights:
        ps = self._params_for_key(head)
        key = (self.precision,) + tuple((p.data_ptr(), p._version) for p in ps)
        if key == self._pack_key:
            return self._weights
        m = self.module
        dev = m.cls_token.device
        if dev.type != "cuda":
            raise _lib.VitkError("the vitk encoder runs on CUDA only - call .to('cuda') first "
                                 "(no CPU fallback)")
        keep = []

        def f32(t: torch.Tensor) -> int:
            t = t.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            keep.append(t)
            return t.data_ptr()

        def bf16(t: torch.Tensor) -> int:
            # channels_last conv weights etc.: reshape to a dense [out, in] matrix first
            t = t.detach().float().reshape(t.shape[0], -1).contiguous()
            s = ops.split3(t, is_weight=True) if self.precision == 1 else ops.cast_bf16(t)
            keep.append(s)
            return s.data_ptr()

        L = len(m.transformer_blocks)
        blocks = (VitkBlockWeights * L)()
        for i, blk in enumerate(m.transformer_blocks):
            b = blocks[i]
            b.ln1_w, b.ln1_b = f32(blk.layer_norm1.weight), f32(blk.layer_norm1.bias)
            b.qkv_w, b.qkv_b = bf16(blk.attention.qkv.weight), f32(blk.attention.qkv.bias)
            b.proj_w, b.proj_b = (bf16(blk.attention.projection.weight),
                                  f32(blk.attention.projection.bias))
            b.ln2_w, b.ln2_b = f32(blk.layer_norm2.weight), f32(blk.layer_norm2.bias)
            b.fc1_w, b.fc1_b = bf16(blk.mlp.linear1.weight), f32(blk.mlp.linear1.bias)
            b.fc2_w, b.fc2_b = bf16(blk.mlp.linear2.weight), f32(blk.mlp.linear2.bias)
            if self.precision == 0:
                # layer_norm1 folded into qkv, layer_norm2 into linear1 (vitk_fold_layernorm):
                # vitk_forward then runs no LayerNorm pass between the blocks
                for name, ln, lin in (("qkv", blk.layer_norm1, blk.attention.qkv),
                                      ("fc1", blk.layer_norm2, blk.mlp.linear1)):
                    w_ln, b_ln, _ = ops.fold_layernorm(lin.weight, ln.weight, ln.bias, lin.bias)
                    keep.extend((w_ln, b_ln))
                    setattr(b, name + "_w_ln", w_ln.data_ptr())
                    setattr(b, name + "_b_ln", b_ln.data_ptr())
        w = VitkWeights()
        w.patch_w = bf16(m.patch_embedding.projection.weight)
        w.patch_b = f32(m.patch_embedding.projection.bias)
        w.cls_token = f32(m.cls_token.reshape(-1))
        dist = getattr(m, "dist_token", None)
        w.dist_token = f32(dist.reshape(-1)) if dist is not None else None
        w.pos_embed = f32(m.position_embedding.reshape(-1, m.position_embedding.shape[-1]))
        w.blocks = C.cast(blocks, C.POINTER(VitkBlockWeights))
        w.ln_f_w, w.ln_f_b = f32(m.layer_norm.weight), f32(m.layer_norm.bias)
        if head is not None:
            w.head_w, w.head_b = f32(head.weight), f32(head.bias)
        self._keep, self._blocks, self._weights, self._pack_key = keep, blocks, w, key
        return w

    # ------------------------------------------------------------------ workspace
    def _workspace(self, cfg: VitkConfig, batch: int, device) -> tuple[int, int]:
        need = C.c_size_t(0)
        check(lib().vitk_workspace_bytes(C.byref(cfg), batch, C.byref(need)))
        if self._ws is None or self._ws_bytes < need.value or self._ws.device != device:
            self._ws = torch.empty(need.value + 1024, dtype=torch.uint8, device=device)
            self._ws_bytes = need.value
        base = (self._ws.data_ptr() + 1023) // 1024 * 1024
        return base, self._ws_bytes

    # ------------------------------------------------------------------ the model call
    #: A.Normalize defaults of the reference's pipelines (evaluation.py:363, train.py:442)
    NORM_MEAN = (0.485, 0.456, 0.406)
    NORM_STD = (0.229, 0.224, 0.225)

    def forward(self, images: torch.Tensor, head: torch.nn.Linear | None = None,
                want_tokens: bool = True, want_logits: bool = False, cls_only_tail: bool = False):
        """images: f32 NCHW [B, C, S, S] (what `images.to(device)` hands the reference's model,
        evaluation.py:499), or u8 NHWC [B, S, S, 3] straight from the decoder - then Normalize +
        ToTensorV2 run on the device inside the patch gather (vitk_forward_u8)."""
        if not images.is_cuda:
            raise _lib.VitkError("images must be a CUDA tensor (no CPU fallback)")
        m = self.module
        pe = m.patch_embedding
        u8 = images.dtype == torch.uint8
        if u8:
            if images.dim() != 4 or images.shape[3] != 3 or images.shape[1] != pe.image_size or \
                    images.shape[2] != pe.image_size or pe.projection.in_channels != 3:
                raise _lib.VitkError(f"expected u8 images of shape [B, {pe.image_size}, "
                                     f"{pe.image_size}, 3], got {tuple(images.shape)}")
        elif images.dtype != torch.float32:
            images = images.float()
        images = images.contiguous()
        if not u8 and (images.dim() != 4 or images.shape[1] != pe.projection.in_channels or
                       images.shape[2] != pe.image_size or images.shape[3] != pe.image_size):
            raise _lib.VitkError(
                f"expected images of shape [B, {pe.projection.in_channels}, {pe.image_size}, "
                f"{pe.image_size}], got {tuple(images.shape)}")
        B = images.shape[0]
        n_classes = head.out_features if (head is not None and want_logits) else 0
        cfg = self.config(n_classes)
        w = self.pack(head if want_logits else None)
        ws, ws_bytes = self._workspace(cfg, B, images.device)
        D = cfg.embed_dim
        N = pe.n_patches + self.n_prefix
        tokens = torch.empty((B, N, D), dtype=torch.float32, device=images.device) \
            if want_tokens else None
        logits = torch.empty((B, n_classes), dtype=torch.float32, device=images.device) \
            if want_logits else None
        if cls_only_tail:
            if u8 or want_tokens or not want_logits:
                raise _lib.VitkError("cls_only_tail: f32 images in, logits only out")
            check(lib().vitk_forward_cls(C.byref(cfg), C.byref(w), images.data_ptr(), B,
                                         logits.data_ptr(), ws, ws_bytes,
                                         torch.cuda.current_stream().cuda_stream))
        elif u8:
            mean, std = (C.c_float * 3)(*self.NORM_MEAN), (C.c_float * 3)(*self.NORM_STD)
            check(lib().vitk_forward_u8(C.byref(cfg), C.byref(w), images.data_ptr(), mean, std, B,
                                        tokens.data_ptr() if tokens is not None else None,
                                        logits.data_ptr() if logits is not None else None,
                                        ws, ws_bytes, torch.cuda.current_stream().cuda_stream))
        else:
            check(lib().vitk_forward(C.byref(cfg), C.byref(w), images.data_ptr(), B,
                                     tokens.data_ptr() if tokens is not None else None,
                                     logits.data_ptr() if logits is not None else None,
                                     ws, ws_bytes, torch.cuda.current_stream().cuda_stream))
        return tokens, logits
