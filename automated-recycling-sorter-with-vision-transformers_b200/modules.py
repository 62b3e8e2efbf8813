"""Drop-in nn.Module mirrors of the reference encoder classes.

Constructor signatures, attribute names, parameter initialisation order (hence the random-init
values under a fixed `torch.manual_seed`) and `state_dict()` keys are those of

    PatchEmbedding            train.py:498-515  == evaluation.py:26-43
    MultiHeadSelfAttention    train.py:518-555  == evaluation.py:46-80
    MLPBlock                  train.py:558-573  == evaluation.py:83-98
    TransformerBlock          train.py:576-593  == evaluation.py:101-117
    VisionTransformer         evaluation.py:120-157
    DataEfficientImageTransformer   train.py:637-688

so a reference checkpoint loads with `load_state_dict` and the reference's own detector wrappers
(`ViTObjectDetector`, `DeiTObjectDetector`) can sit on top unchanged.  The arithmetic is NOT
PyTorch: `forward` marshals pointers into libvitk (hand-written sm_100a kernels).  The sub-modules
are parameter containers whose own `forward` runs the matching per-operator kernels.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops
from .engine import EncoderEngine


class PatchEmbedding(nn.Module):
    def __init__(self, image_size=224, patch_size=16, in_channels=3, embed_dim=768):
        super().__init__()
        self.image_size = image_size
        self.patch_size = patch_size
        self.n_patches = (image_size // patch_size) ** 2
        # parameter container with the reference's shapes/initialisation; never called as a conv
        self.projection = nn.Conv2d(in_channels, embed_dim, kernel_size=patch_size,
                                    stride=patch_size)

    def forward(self, x):
        """[B,C,S,S] f32 -> [B,P,D] f32 : im2col gather + tcgen05 GEMM (+bias)."""
        B = x.shape[0]
        D = self.projection.out_channels
        a = ops.patchify(x.float(), self.patch_size)
        w = ops.cast_bf16(self.projection.weight.detach().reshape(D, -1).contiguous())
        out = ops.gemm(a, w, _lib.EPI_F32, bias=self.projection.bias.detach().float().contiguous())
        return out.view(B, self.n_patches, D)


class MultiHeadSelfAttention(nn.Module):
    def __init__(self, embed_dim=768, num_heads=12, dropout=0.1):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        assert embed_dim % num_heads == 0
        self.qkv = nn.Linear(embed_dim, embed_dim * 3)
        self.attention_dropout = nn.Dropout(dropout)
        self.projection = nn.Linear(embed_dim, embed_dim)
        self.projection_dropout = nn.Dropout(dropout)

    def forward(self, x):
        """[B,N,D] f32 -> [B,N,D] f32 (inference semantics: dropout is the identity)."""
        B, N, D = x.shape
        a = ops.cast_bf16(x.reshape(B * N, D).float().contiguous())
        qkv = ops.gemm(a, ops.cast_bf16(self.qkv.weight.detach()), _lib.EPI_BF16,
                       bias=self.qkv.bias.detach())
        ctx = ops.attention(qkv, B, N, self.num_heads)
        out = ops.gemm(ctx, ops.cast_bf16(self.projection.weight.detach()), _lib.EPI_F32,
                       bias=self.projection.bias.detach())
        return out.view(B, N, D)


class MLPBlock(nn.Module):
    def __init__(self, embed_dim=768, mlp_dim=3072, dropout=0.1):
        super().__init__()
        self.linear1 = nn.Linear(embed_dim, mlp_dim)
        self.gelu = nn.GELU()
        self.dropout1 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(mlp_dim, embed_dim)
        self.dropout2 = nn.Dropout(dropout)

    def forward(self, x):
        shp = x.shape
        a = ops.cast_bf16(x.reshape(-1, shp[-1]).float().contiguous())
        h = ops.gemm(a, ops.cast_bf16(self.linear1.weight.detach()), _lib.EPI_GELU_BF16,
                     bias=self.linear1.bias.detach())
        out = ops.gemm(h, ops.cast_bf16(self.linear2.weight.detach()), _lib.EPI_F32,
                       bias=self.linear2.bias.detach())
        return out.view(shp)


class TransformerBlock(nn.Module):
    def __init__(self, embed_dim=768, num_heads=12, mlp_dim=3072, dropout=0.1):
        super().__init__()
        self.attention = MultiHeadSelfAttention(embed_dim, num_heads, dropout)
        self.mlp = MLPBlock(embed_dim, mlp_dim, dropout)
        self.layer_norm1 = nn.LayerNorm(embed_dim)
        self.layer_norm2 = nn.LayerNorm(embed_dim)

    def forward(self, x):
        B, N, D = x.shape
        x2 = x.reshape(B * N, D).float().contiguous()
        ln1, ln2 = self.layer_norm1, self.layer_norm2
        att, mlp = self.attention, self.mlp
        xn = ops.layernorm(x2, ln1.weight.detach(), ln1.bias.detach(), ln1.eps)
        qkv = ops.gemm(xn, ops.cast_bf16(att.qkv.weight.detach()), _lib.EPI_BF16,
                       bias=att.qkv.bias.detach())
        ctx = ops.attention(qkv, B, N, att.num_heads)
        x2 = ops.gemm(ctx, ops.cast_bf16(att.projection.weight.detach()), _lib.EPI_RESID_F32,
                      bias=att.projection.bias.detach(), resid=x2)
        xn = ops.layernorm(x2, ln2.weight.detach(), ln2.bias.detach(), ln2.eps)
        h = ops.gemm(xn, ops.cast_bf16(mlp.linear1.weight.detach()), _lib.EPI_GELU_BF16,
                     bias=mlp.linear1.bias.detach())
        x2 = ops.gemm(h, ops.cast_bf16(mlp.linear2.weight.detach()), _lib.EPI_RESID_F32,
                      bias=mlp.linear2.bias.detach(), resid=x2)
        return x2.view(B, N, D)


class _EncoderBase(nn.Module):
    _n_prefix = 1

    def _engine(self) -> EncoderEngine:
        eng = self.__dict__.get("_vitk_engine")
        if eng is None:
            eng = EncoderEngine(self, self._n_prefix)
            self.__dict__["_vitk_engine"] = eng  # not a sub-module / not in state_dict
        return eng

    def set_precision(self, mode: str):
        """'bf16' (default): bf16 operands, fp32 accumulation / residual / LayerNorm statistics.
        'fp32': parity mode - fp32 activations, contractions on three-term bf16 splits (logits
        within 1e-4 of the fp32 reference); inference only."""
        if mode not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self._engine().precision = 1 if mode == "fp32" else 0
        return self

    def forward(self, x):
        """images f32 [B,C,S,S] -> all tokens after the final LayerNorm, f32 [B,N,D]."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if self._engine().precision == 1:
                raise _lib.VitkError("the fp32-parity mode is inference-only: the differentiable "
                                     "path computes in bf16 (call under torch.no_grad(), or "
                                     "set_precision('bf16'))")
            from .autograd import encoder_forward_train  # training path (saves activations)
            return encoder_forward_train(self, x)
        tokens, _ = self._engine().forward(x, want_tokens=True)
        return tokens

    def classify(self, x, head: nn.Linear, cls_only_tail: bool = False):
        """logits f32 [B, n_classes] = head(LN(tokens)[:, 0]) in one library call.
        cls_only_tail: evaluate the last block for the CLS rows only (vitk_forward_cls) - the same
        logits, without the 71 % of the last block's work that feeds nothing the head reads."""
        _, logits = self._engine().forward(x, head=head, want_tokens=False, want_logits=True,
                                           cls_only_tail=cls_only_tail)
        return logits


class VisionTransformer(_EncoderBase):
    """evaluation.py:120-157 (CLS + patches, randn tokens)."""
    _n_prefix = 1

    def __init__(self, image_size=224, patch_size=16, in_channels=3, embed_dim=768,
                 num_layers=12, num_heads=12, mlp_dim=3072, dropout=0.1, num_classes=1000):
        super().__init__()
        self.patch_embedding = PatchEmbedding(image_size, patch_size, in_channels, embed_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.position_embedding = nn.Parameter(
            torch.randn(1, self.patch_embedding.n_patches + 1, embed_dim))
        self.dropout = nn.Dropout(dropout)
        self.transformer_blocks = nn.ModuleList([
            TransformerBlock(embed_dim, num_heads, mlp_dim, dropout) for _ in range(num_layers)])
        self.layer_norm = nn.LayerNorm(embed_dim)


class DataEfficientImageTransformer(_EncoderBase):
    """train.py:637-688 (CLS + DIST + patches, trunc-normal(0.02) tokens)."""
    _n_prefix = 2

    def __init__(self, image_size=224, patch_size=16, in_channels=3, embed_dim=768,
                 num_layers=12, num_heads=12, mlp_dim=3072, dropout=0.1, num_classes=1000):
        super().__init__()
        self.patch_embedding = PatchEmbedding(image_size, patch_size, in_channels, embed_dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.randn(1, 1, embed_dim))
        self.position_embedding = nn.Parameter(
            torch.randn(1, self.patch_embedding.n_patches + 2, embed_dim))
        self.dropout = nn.Dropout(dropout)
        self.transformer_blocks = nn.ModuleList([
            TransformerBlock(embed_dim, num_heads, mlp_dim, dropout) for _ in range(num_layers)])
        self.layer_norm = nn.LayerNorm(embed_dim)
        self._init_weights()

    def _init_weights(self):
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        nn.init.trunc_normal_(self.dist_token, std=0.02)
        nn.init.trunc_normal_(self.position_embedding, std=0.02)


class ViTClassifier(nn.Module):
    """north_star's recyclable-waste classifier: `Linear(D, n_classes)` on the CLS row of the
    backbone output (the reference has no such head; closest analogues train.py:711,827).  The
    head is constructed right after the backbone so that, under a fixed seed, weights equal the
    oracle's."""

    def __init__(self, num_classes=6, deit=False, **backbone_kwargs):
        super().__init__()
        cls = DataEfficientImageTransformer if deit else VisionTransformer
        self.backbone = cls(**backbone_kwargs)
        self.head = nn.Linear(self.backbone.layer_norm.normalized_shape[0], num_classes)

    def forward(self, images):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if self.backbone._engine().precision == 1:
                raise _lib.VitkError("the fp32-parity mode is inference-only: the differentiable "
                                     "path computes in bf16 (call under torch.no_grad(), or "
                                     "set_precision('bf16'))")
            from .autograd import classifier_forward_train
            return classifier_forward_train(self, images)
        return self.backbone.classify(images, self.head)

    @torch.no_grad()
    def classify_pruned(self, images):
        """Inference logits with the last encoder block evaluated for the CLS rows only (opt-in;
        `forward` always evaluates every token, as the reference does)."""
        return self.backbone.classify(images, self.head, cls_only_tail=True)

    @torch.no_grad()
    def predict(self, images):
        """top-1 class per image (cf. evaluation.py:403-404 argmax over class scores)."""
        return self.forward(images).argmax(dim=-1)

    @torch.no_grad()
    def predict_with_scores(self, images):
        """(scores, labels): softmax confidence and class of the top-1 prediction per image - the
        device-side form of post_process_predictions' softmax / max (evaluation.py:403-404)."""
        from . import ops
        return ops.postprocess_scores(self.forward(images))
