"""LayerNorm folded into the GEMMs around it (vitk_fold_layernorm, vitk_row_stats,
vitk_gemm_resid_stats, vitk_gemm_layernorm_folded) against torch:  x = x + f(..);
Linear(layer_norm(x))  (reference train.py:584-591)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (M, N = features of x, K of the residual GEMM)
SHAPES = [
    (300, 768, 768),        # ragged M, three 256-column tiles
    (197 * 8, 768, 3072),   # linear2 shape
    (1000, 384, 1536),      # 128-column tiles (64 columns per epilogue warp)
    (77, 64, 128),          # single CTA, one tile
    (600, 1024, 1024),      # ViT-L width
    (333, 400, 400),        # the reference's own Config width (train.py:1345): ragged columns
]


@pytest.fixture(params=[1, 2], ids=["cta1", "cta2"], autouse=True)
def gemm_mode(request, vitk):
    vitk._lib.set_gemm_cta_group(request.param)
    yield request.param
    vitk._lib.set_gemm_cta_group(0)


def _mk(M, N, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    w = (torch.randn(N, K, generator=g, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, generator=g, device="cuda")
    x = torch.randn(M, N, generator=g, device="cuda") * 2 + 0.3
    return a, w, bias, x


def test_fold_layernorm(vitk):
    g = torch.Generator(device="cuda").manual_seed(3)
    W = torch.randn(200, 400, generator=g, device="cuda") / 20
    gamma = torch.rand(400, generator=g, device="cuda") + 0.5
    beta = torch.randn(400, generator=g, device="cuda") * 0.1
    b = torch.randn(200, generator=g, device="cuda")
    w_ln, b_ln, rowsum = vitk.ops.fold_layernorm(W, gamma, beta, b)
    v = W * gamma
    ref_w = v - v.mean(1, keepdim=True)
    # centred rows, rounded to bf16 (a last-bit difference where fma and mul + sub round apart)
    torch.testing.assert_close(w_ln.float(), ref_w, rtol=2 ** -7, atol=1e-6)
    torch.testing.assert_close(rowsum, w_ln.float().sum(1), rtol=1e-4, atol=1e-5)
    assert rowsum.abs().max().item() < 5e-3      # what rounding leaves of a zero row sum
    torch.testing.assert_close(b_ln, b + W @ beta, rtol=1e-5, atol=1e-5)
    _, b0, _ = vitk.ops.fold_layernorm(W, gamma, beta, None)
    torch.testing.assert_close(b0, W @ beta, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("rows,D", [(197 * 4, 768), (50, 400), (33, 1024)])
def test_row_stats(vitk, rows, D):
    x = torch.randn(rows, D, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    xb, st = vitk.ops.row_stats(x)
    assert torch.equal(xb, x.bfloat16())
    torch.testing.assert_close(st[0, :, 0], x.sum(1), rtol=1e-5, atol=1e-4)
    torch.testing.assert_close(st[0, :, 1], (x * x).sum(1), rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_resid_stats(vitk, M, N, K):
    a, w, bias, x = _mk(M, N, K)
    ref = x + a.float() @ w.float().t() + bias
    xb, st = vitk.ops.gemm_resid_stats(a, w, x, bias=bias)
    torch.testing.assert_close(x, ref, rtol=1e-4, atol=1e-4)
    assert torch.equal(xb, x.bfloat16())          # the bf16 copy is of the value that was stored
    assert st.shape[0] == vitk._lib.lib().vitk_stats_parts(N)
    tot = st.sum(0)
    torch.testing.assert_close(tot[:, 0], x.sum(1), rtol=1e-5, atol=2e-3)
    torch.testing.assert_close(tot[:, 1], (x * x).sum(1), rtol=1e-5, atol=2e-3)
    # partial p covers a contiguous group of columns
    cols = -(-N // st.shape[0])
    cols = 64 if cols <= 64 else 128
    for p in range(st.shape[0]):
        torch.testing.assert_close(st[p, :, 0], x[:, p * cols:(p + 1) * cols].sum(1), rtol=1e-5,
                                   atol=2e-3)


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("epi", ["bf16", "gelu"])
def test_linear_of_layernorm_folded(vitk, M, N, K, epi):
    """The whole chain: residual GEMM with statistics, then Linear(LayerNorm(x)) with the
    normalisation in the consumer's epilogue, against torch on the fp32 x."""
    a, w, bias, x = _mk(M, N, K, seed=2)
    g = torch.Generator(device="cuda").manual_seed(5)
    N2 = 3 * N if epi == "bf16" else 4 * N
    W2 = torch.randn(N2, N, generator=g, device="cuda") / N ** 0.5
    b2 = torch.randn(N2, generator=g, device="cuda") * 0.1
    gamma = torch.rand(N, generator=g, device="cuda") + 0.5
    beta = torch.randn(N, generator=g, device="cuda") * 0.1
    xb, st = vitk.ops.gemm_resid_stats(a, w, x, bias=bias)
    w_ln, b_ln, _ = vitk.ops.fold_layernorm(W2, gamma, beta, b2)
    e = vitk._lib.EPI_BF16 if epi == "bf16" else vitk._lib.EPI_GELU_TANH_BF16
    out = vitk.ops.gemm_layernorm_folded(xb, w_ln, b_ln, st, epilogue=e)
    ref = F.linear(F.layer_norm(x, (N,), gamma, beta, 1e-5), W2, b2)
    if epi == "gelu":
        ref = F.gelu(ref)
    err = (out.float() - ref).abs().max().item()
    # same bar as the unfolded bf16 path (bf16 operands, fp32 accumulate, bf16 output)
    two = F.linear(F.layer_norm(x, (N,), gamma, beta, 1e-5).bfloat16().float(),
                   W2.bfloat16().float(), b2)
    if epi == "gelu":
        two = F.gelu(two)
    err_two = (two.bfloat16().float() - ref).abs().max().item()
    assert err < max(3 * err_two, 3e-2), (err, err_two)
    # entry of the chain (one partial per row) gives the same answer from the same x
    xb1, st1 = vitk.ops.row_stats(x)
    out1 = vitk.ops.gemm_layernorm_folded(xb1, w_ln, b_ln, st1, epilogue=e)
    torch.testing.assert_close(out1.float(), out.float(), rtol=2e-2, atol=2e-2)


def test_rows_with_a_large_common_offset(vitk):
    """The folded form relies on zero-sum weight rows to drop the row mean inside the contraction;
    bf16 rounding leaves ~1e-3 of the row sum, and the bf16 copy of x is rounded relative to |x|
    and not |x - mean|: rows whose mean is large against their spread are where it loses accuracy
    first.  Still inside the bar at |mean| / std = 4 (encoder rows sit far below 1)."""
    M, N = 256, 768
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn(M, N, generator=g, device="cuda") + 4.0
    W2 = torch.randn(N, N, generator=g, device="cuda") / N ** 0.5
    gamma, beta = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
    xb, st = vitk.ops.row_stats(x)
    w_ln, b_ln, _ = vitk.ops.fold_layernorm(W2, gamma, beta, None)
    out = vitk.ops.gemm_layernorm_folded(xb, w_ln, b_ln, st)
    ref = F.linear(F.layer_norm(x, (N,)), W2)
    assert (out.float() - ref).abs().max().item() < 6e-2


def test_model_folded_matches_unfolded(vitk):
    """vitk_forward with and without the folding on the same ViT-B/16-width model."""
    torch.manual_seed(0)
    model = vitk.ViTClassifier(num_classes=6, dropout=0.0, image_size=224, patch_size=16,
                               embed_dim=768, num_layers=4, num_heads=12, mlp_dim=3072).cuda().eval()
    x = torch.randn(5, 3, 224, 224, device="cuda")
    with torch.no_grad():
        model(x)                       # packs the weights (launches of its own)
    n0 = vitk.launch_count()
    with torch.no_grad():
        a = model(x)
    n_folded = vitk.launch_count() - n0
    vitk._lib.set_layernorm_folding(False)
    try:
        n0 = vitk.launch_count()
        with torch.no_grad():
            b = model(x)
        n_plain = vitk.launch_count() - n0
    finally:
        vitk._lib.set_layernorm_folding(True)
    assert n_folded < n_plain          # the LayerNorm launches between the blocks are gone
    assert (a - b).abs().max().item() < 1e-2
