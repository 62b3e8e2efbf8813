"""ORACLE - test infrastructure only.  NOT part of the product path.

A CPU restatement (PyTorch tensor arithmetic, fp32 or fp64) of the reference's ViT/DeiT encoder
forward, the north-star 6-class CLS head, and one fine-tune train step, written from the
reference's algorithm and citing the lines it follows.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu-baseline / `--impl reference` legs may import this file; the product
(`automated-recycling-sorter-with-vision-transformers_b200/`) never does and has no CPU fallback.

Pinning: the reference ships no tests, golden vectors or fixtures for this path, so
the oracle is pinned against the reference's OWN classes executed in the build container:
`oracle/gen_golden.py` imports /root/reference/{evaluation,train}.py, runs them on seeded inputs and
commits the outputs under tests/golden/; `tests/test_oracle.py` checks this restatement against
those fixtures (and, when /root/reference is present, against the live classes).

The third-party arithmetic underneath the reference is PyTorch (unpinned by the reference;
2.11.0 in this image): nn.Conv2d, nn.Linear, torch.matmul, torch.softmax, nn.GELU (erf form),
nn.LayerNorm (eps 1e-5, biased variance), torch.optim.AdamW.  Each is restated below from its
published definition rather than by calling the nn.Module.
"""
from __future__ import annotations

import math

import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------------
# synthetic inputs: what Resize->Normalize->ToTensorV2 (evaluation.py:360-366)
# yields for an 8-bit RGB image
# --------------------------------------------------------------------------------------------
def synthetic_images(batch: int, size: int = 224, seed: int = 1234) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (batch, size, size, 3), generator=g, dtype=torch.uint8)
    x = u8.to(torch.float32) / 255.0
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32)
    x = (x - mean) / std
    return x.permute(0, 3, 1, 2).contiguous()


def synthetic_labels(batch: int, n_classes: int = 6, seed: int = 1) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, n_classes, (batch,), generator=g)


# --------------------------------------------------------------------------------------------
# operators
# --------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5):
    """nn.LayerNorm(D) - train.py:581-582,586,590,656,687: per-row mean, biased variance,
    eps inside the square root, affine."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: torch.Tensor) -> torch.Tensor:
    """nn.GELU() default (approximate='none') - train.py:562: 0.5 x (1 + erf(x / sqrt 2))."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def patch_embed(images: torch.Tensor, w: torch.Tensor, b: torch.Tensor, p: int) -> torch.Tensor:
    """PatchEmbedding.forward - train.py:511-515.  Conv2d(k=p, stride=p) over non-overlapping
    patches == one matrix product of the gathered patches [B*P, C*p*p] (column order c, kh, kw,
    the order of conv.weight.reshape(D, -1)) with the weight; `flatten(2).transpose(1, 2)` makes
    the token order ph * (S/p) + pw."""
    B, C, S, _ = images.shape
    g = S // p
    cols = images.reshape(B, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, C * p * p)
    return cols @ w.reshape(w.shape[0], -1).t() + b


def _drop(x: torch.Tensor, masks, site: str, layer: int) -> torch.Tensor:
    """nn.Dropout in training mode with an explicit mask: masks[(site, layer)] holds 0 for a
    dropped element and 1 / (1 - p) for a kept one (identity when masks is None / site absent)."""
    if not masks or (site, layer) not in masks:
        return x
    return x * masks[(site, layer)].to(x.dtype)


def attention(x: torch.Tensor, qkv_w, qkv_b, proj_w, proj_b, num_heads: int, masks=None,
              layer: int = 0) -> torch.Tensor:
    """MultiHeadSelfAttention.forward - train.py:532-555 (dropout = identity unless masks given:
    attention_dropout on the probabilities :545, projection_dropout on the output :553)."""
    B, N, D = x.shape
    hd = D // num_heads
    qkv = x @ qkv_w.t() + qkv_b                                    # :536
    qkv = qkv.reshape(B, N, 3, num_heads, hd).permute(2, 0, 3, 1, 4)  # :537-538
    q, k, v = qkv[0], qkv[1], qkv[2]                               # :540
    scores = (q @ k.transpose(-2, -1)) / (hd ** 0.5)               # :543 (divide AFTER the product)
    probs = torch.softmax(scores, dim=-1)                          # :544
    probs = _drop(probs, masks, "attn", layer)                     # :545
    ctx = probs @ v                                                # :548
    ctx = ctx.transpose(1, 2).reshape(B, N, D)                     # :549
    return _drop(ctx @ proj_w.t() + proj_b, masks, "proj", layer)  # :552-553


def mlp(x: torch.Tensor, w1, b1, w2, b2, masks=None, layer: int = 0) -> torch.Tensor:
    """MLPBlock.forward - train.py:567-573 (dropout1 after the GELU :570, dropout2 at the end :572)."""
    h = _drop(gelu_erf(x @ w1.t() + b1), masks, "gelu", layer)
    return _drop(h @ w2.t() + b2, masks, "fc2", layer)


def encoder_block(x: torch.Tensor, sd: dict, prefix: str, num_heads: int, masks=None,
                  layer: int = 0) -> torch.Tensor:
    """TransformerBlock.forward - train.py:584-593 (pre-LN residual block)."""
    g = lambda k: sd[prefix + k]
    x = x + attention(layer_norm(x, g("layer_norm1.weight"), g("layer_norm1.bias")),
                      g("attention.qkv.weight"), g("attention.qkv.bias"),
                      g("attention.projection.weight"), g("attention.projection.bias"), num_heads,
                      masks, layer)
    x = x + mlp(layer_norm(x, g("layer_norm2.weight"), g("layer_norm2.bias")),
                g("mlp.linear1.weight"), g("mlp.linear1.bias"),
                g("mlp.linear2.weight"), g("mlp.linear2.bias"), masks, layer)
    return x


def infer_dims(sd: dict, prefix: str = "") -> dict:
    w = sd[prefix + "patch_embedding.projection.weight"]
    n_layers = 1 + max(int(k[len(prefix):].split(".")[1]) for k in sd
                       if k.startswith(prefix + "transformer_blocks."))
    return dict(embed_dim=w.shape[0], in_channels=w.shape[1], patch_size=w.shape[2],
                num_layers=n_layers, n_tokens=sd[prefix + "position_embedding"].shape[1],
                n_prefix=2 if (prefix + "dist_token") in sd else 1)


def backbone_forward(sd: dict, images: torch.Tensor, num_heads: int, prefix: str = "",
                     dtype: torch.dtype = torch.float32, masks=None) -> torch.Tensor:
    """VisionTransformer.forward (evaluation.py:138-157) / DataEfficientImageTransformer.forward
    (train.py:666-688): returns all tokens after the final LayerNorm.  Eval mode unless `masks`
    (explicit dropout masks keyed by (site, layer), see _drop) is given."""
    sd = {k: v.to(dtype) for k, v in sd.items() if k.startswith(prefix)}
    dims = infer_dims(sd, prefix)
    g = lambda k: sd[prefix + k]
    x = patch_embed(images.to(dtype), g("patch_embedding.projection.weight"),
                    g("patch_embedding.projection.bias"), dims["patch_size"])
    B = x.shape[0]
    toks = [g("cls_token").expand(B, -1, -1)]
    if dims["n_prefix"] == 2:
        toks.append(g("dist_token").expand(B, -1, -1))         # train.py:670-673
    x = torch.cat(toks + [x], dim=1) + g("position_embedding")  # evaluation.py:145-149
    x = _drop(x, masks, "embed", 0)                              # evaluation.py:150 / train.py:681
    for i in range(dims["num_layers"]):                          # evaluation.py:153-154
        x = encoder_block(x, sd, f"{prefix}transformer_blocks.{i}.", num_heads, masks, i)
    return layer_norm(x, g("layer_norm.weight"), g("layer_norm.bias"))  # evaluation.py:156


def classifier_forward(sd: dict, images: torch.Tensor, num_heads: int,
                       dtype: torch.dtype = torch.float32, masks=None):
    """RefClassifier of SURVEY.md 8c: Linear(D, C)(backbone(images)[:, 0]).
    sd uses the keys of ViTClassifier: 'backbone.*', 'head.weight', 'head.bias'."""
    tokens = backbone_forward(sd, images, num_heads, prefix="backbone.", dtype=dtype, masks=masks)
    logits = tokens[:, 0] @ sd["head.weight"].to(dtype).t() + sd["head.bias"].to(dtype)
    return tokens, logits


# --------------------------------------------------------------------------------------------
# train step (train.py:1441-1460 with the north-star's 6-class cross-entropy in place of the
# detection loss; optimizer per train.py:1598-1602)
# --------------------------------------------------------------------------------------------
def cross_entropy(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """mean over the batch of -log softmax(logits)[label] (F.cross_entropy default reduction)."""
    lse = torch.logsumexp(logits, dim=-1)
    return (lse - logits.gather(1, labels[:, None]).squeeze(1)).mean()


def adamw_step(params: dict, grads: dict, m: dict, v: dict, step: int, lr=1e-4, betas=(0.9, 0.999),
               eps=1e-8, weight_decay=1e-4) -> None:
    """torch.optim.AdamW as train.py:1598-1602 configures it: one group, decoupled decay on every
    parameter.  In place on params / m / v; `step` is 1-based."""
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k, p in params.items():
        g = grads[k]
        p.mul_(1.0 - lr * weight_decay)
        m[k].mul_(b1).add_(g, alpha=1.0 - b1)
        v[k].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = (v[k].sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(m[k], denom, value=-lr / bc1)


def train_step(sd: dict, images: torch.Tensor, labels: torch.Tensor, num_heads: int,
               dtype: torch.dtype = torch.float32, lr=1e-4, weight_decay=1e-4, masks=None):
    """One fine-tune step (dropout = 0, or the given explicit masks): returns (loss, grads,
    new_params)."""
    params = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    _, logits = classifier_forward(params, images, num_heads, dtype=dtype, masks=masks)
    loss = cross_entropy(logits, labels)
    grads_list = torch.autograd.grad(loss, list(params.values()))
    grads = dict(zip(params.keys(), grads_list))
    new = {k: v.detach().clone() for k, v in params.items()}
    m = {k: torch.zeros_like(v) for k, v in new.items()}
    vv = {k: torch.zeros_like(v) for k, v in new.items()}
    adamw_step(new, grads, m, vv, step=1, lr=lr, weight_decay=weight_decay)
    return loss.detach(), grads, new


# --------------------------------------------------------------------------------------------
# detection head: ObjectDetectionHead, evaluation.py:160-200 == train.py:691-731
#
# The decoder is torch's nn.TransformerDecoder(nn.TransformerDecoderLayer(d_model=D, nhead=8,
# dim_feedforward=2048, dropout=0.1, batch_first=True), num_layers=6) - third-party code (PyTorch,
# unpinned by the reference; 2.11.0 here).  Restated from its published definition for the default
# norm_first=False, activation=relu, no masks, eval mode:
#     x = norm1(x + self_attn(x, x, x)); x = norm2(x + multihead_attn(x, memory, memory));
#     x = norm3(x + linear2(relu(linear1(x))))
# nn.MultiheadAttention: q/k/v = rows [0,D) / [D,2D) / [2D,3D) of in_proj_weight (+ in_proj_bias),
# heads split along the feature axis, softmax(q k^T / sqrt(head_dim)) v, then out_proj.
# Pinned by tests/golden/det_head_*.npz (the reference's own class run in the build container).
# --------------------------------------------------------------------------------------------
def multihead_attention(q_in: torch.Tensor, kv_in: torch.Tensor, in_w, in_b, out_w, out_b,
                        num_heads: int, masks=None, site: str = "", layer: int = 0) -> torch.Tensor:
    B, Nq, D = q_in.shape
    Nk = kv_in.shape[1]
    hd = D // num_heads
    q = q_in @ in_w[:D].t() + in_b[:D]
    k = kv_in @ in_w[D:2 * D].t() + in_b[D:2 * D]
    v = kv_in @ in_w[2 * D:].t() + in_b[2 * D:]
    q = q.reshape(B, Nq, num_heads, hd).transpose(1, 2)
    k = k.reshape(B, Nk, num_heads, hd).transpose(1, 2)
    v = v.reshape(B, Nk, num_heads, hd).transpose(1, 2)
    probs = torch.softmax((q @ k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    probs = _drop(probs, masks, site, layer)       # nn.MultiheadAttention(dropout=0.1), train mode
    ctx = (probs @ v).transpose(1, 2).reshape(B, Nq, D)
    return ctx @ out_w.t() + out_b


def decoder_layer(x: torch.Tensor, memory: torch.Tensor, sd: dict, prefix: str,
                  num_heads: int = 8, masks=None, layer: int = 0) -> torch.Tensor:
    """Train mode (train.py:701-707, dropout = 0.1) when `masks` is given: explicit dropout masks
    keyed (site, layer), sites dec_sa_attn / dec_sa_out (dropout1) / dec_ca_attn / dec_ca_out
    (dropout2) / dec_ffn (after the ReLU) / dec_ff2 (dropout3) - see _drop."""
    g = lambda k: sd[prefix + k]
    sa = multihead_attention(x, x, g("self_attn.in_proj_weight"), g("self_attn.in_proj_bias"),
                             g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"),
                             num_heads, masks, "dec_sa_attn", layer)
    x = layer_norm(x + _drop(sa, masks, "dec_sa_out", layer), g("norm1.weight"), g("norm1.bias"))
    ca = multihead_attention(x, memory, g("multihead_attn.in_proj_weight"),
                             g("multihead_attn.in_proj_bias"), g("multihead_attn.out_proj.weight"),
                             g("multihead_attn.out_proj.bias"), num_heads, masks, "dec_ca_attn",
                             layer)
    x = layer_norm(x + _drop(ca, masks, "dec_ca_out", layer), g("norm2.weight"), g("norm2.bias"))
    h = _drop(torch.relu(x @ g("linear1.weight").t() + g("linear1.bias")), masks, "dec_ffn", layer)
    ff = h @ g("linear2.weight").t() + g("linear2.bias")
    return layer_norm(x + _drop(ff, masks, "dec_ff2", layer), g("norm3.weight"), g("norm3.bias"))


def detection_head_forward(sd: dict, encoder_features: torch.Tensor, prefix: str = "",
                           dtype: torch.dtype = torch.float32, num_heads: int = 8,
                           masks=None) -> dict:
    """ObjectDetectionHead.forward - evaluation.py:183-200.  encoder_features [B, P, D] is the
    memory (the callers strip the CLS / DIST rows first: evaluation.py:235, train.py:829)."""
    sd = {k: v.to(dtype) for k, v in sd.items() if k.startswith(prefix)}
    g = lambda k: sd[prefix + k]
    memory = encoder_features.to(dtype)
    B = memory.shape[0]
    x = g("object_queries").unsqueeze(0).expand(B, -1, -1)          # :186
    n_layers = 1 + max(int(k[len(prefix):].split(".")[2]) for k in sd
                       if k.startswith(prefix + "decoder.layers."))
    for i in range(n_layers):                                       # :189
        x = decoder_layer(x, memory, sd, f"{prefix}decoder.layers.{i}.", num_heads, masks, i)
    class_logits = x @ g("class_head.weight").t() + g("class_head.bias")            # :192
    bbox = torch.sigmoid(x @ g("bbox_head.weight").t() + g("bbox_head.bias"))      # :193-194
    return {"class_logits": class_logits, "bbox_coords": bbox}


def randomize_head_state(sd: dict, seed: int) -> dict:
    """Deterministic non-degenerate parameters for the head fixtures: torch's constructor clones
    one decoder layer six times and zero-initialises the attention biases, which would leave the
    bias paths and per-layer indexing untested.  Matrices ~ N(0, 1/fan_in), biases ~ N(0, 0.1^2),
    LayerNorm weights 1 + N(0, 0.1^2), queries N(0, 1); keys visited in sorted order."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k in sorted(sd):
        v = sd[k]
        r = torch.randn(v.shape, generator=g, dtype=torch.float32)
        if k.endswith("object_queries"):
            out[k] = r
        elif ".norm" in k and k.endswith("weight"):
            out[k] = 1.0 + 0.1 * r
        elif v.dim() == 2:
            out[k] = r / math.sqrt(v.shape[1])
        else:
            out[k] = 0.1 * r
    return out


def post_process_predictions(outputs: dict, confidence_threshold: float = 0.5) -> list:
    """evaluation.py:393-426 restated: per image softmax over the class logits (:403), best
    non-background probability and class (:404), `max_probs > confidence_threshold` (:407),
    boolean-mask selection of boxes / labels / scores (:410-412), empty tensors otherwise."""
    res = []
    for logits, boxes in zip(outputs["class_logits"], outputs["bbox_coords"]):
        probs = torch.softmax(logits, dim=-1)
        max_probs, labels = torch.max(probs[:, :-1], dim=-1)
        keep = max_probs > confidence_threshold
        res.append({"boxes": boxes[keep], "labels": labels[keep], "scores": max_probs[keep]})
    return res


def triplet_features(cls_tokens: torch.Tensor, w: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """DeiTObjectDetector.forward, train.py:833-838: `triplet_projection` on the CLS row, then
    F.normalize(p=2, dim=1) (x / max(||x||_2, 1e-12))."""
    y = cls_tokens @ w.t() + b
    return y / y.norm(dim=1, keepdim=True).clamp_min(1e-12)
