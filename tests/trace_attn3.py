"""Not a pytest file: unit timeline of attn_fwd_tc3_kernel (library built with -DVITK_ATTN_TRACE)."""
import sys, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import vitk
B, N, H = 64, 577, 12
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
for _ in range(2):
    ctx, lse = vitk.ops.attention(qkv, B, N, H, return_lse=True)
torch.cuda.synchronize()
tr = lse.flatten()[:1536].view(torch.int64).reshape(3, 64, 4)
t0 = tr[1, 0, 0].item()
for u in range(0, 32):
    # five key blocks (units) per q-tile at N = 577
    print("unit %2d %s  S-issue: start %6d end %6d (%4d) | softmax: s_full %6d done %6d (%5d) | complete: done-seen %6d issued %6d (%5d)" % (
        u, "kb%d" % (u % 5), tr[2, u, 0] - t0, tr[2, u, 1] - t0, tr[2, u, 1] - tr[2, u, 0],
        tr[1, u, 0] - t0, tr[1, u, 1] - t0, tr[1, u, 1] - tr[1, u, 0],
        tr[0, u, 0] - t0, tr[0, u, 1] - t0, tr[0, u, 1] - tr[0, u, 0]))
