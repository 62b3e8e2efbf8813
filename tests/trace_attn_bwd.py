"""Not a pytest file: phase clocks of the pipelined attention backward.  Needs a library whose
attention_bwd_tc2.cu was compiled with -DVITK_ATTN_TRACE, selected with VITK_LIB:
    cd automated-...; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
        -Xcompiler -fPIC --expt-relaxed-constexpr -DVITK_ATTN_TRACE -c csrc/attention_bwd_tc2.cu -o /tmp/t.o
    nvcc -shared -o libvitk_trace.so $(ls csrc/obj/*.o | grep -v attention_bwd_tc2.o) /tmp/t.o -lcudart
Prints, per item, where the first softmax warp, the MMA warp and a row-vector warp spend their
time (batch 128, 197 tokens, 12 heads)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import vitk  # noqa: E402
from vitk import _lib as E  # noqa: E402

B, N, H = 128, 197, 12
qkv = (torch.randn(B * N, 3 * H * 64, device="cuda") * 1.5).bfloat16()
ctx, lse = vitk.ops.attention(qkv, B, N, H, return_lse=True)
dctx = (torch.randn_like(ctx.float()) * 0.1).bfloat16()
for _ in range(3):
    vitk.ops.attention_bwd(qkv, ctx, dctx, lse, B, N, H)
torch.cuda.synchronize()
fn = E.lib().vitk_debug_attn_bwd_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_longlong * (320 * 16))()
n = fn(buf, 320 * 16)
assert n > 0
ctas = 148
items = B * H / ctas
rows = [buf[16 * i:16 * i + 16] for i in range(ctas)]
avg = [sum(r[k] for r in rows) / ctas / items for k in range(16)]
print(f"per item ({items:.1f} items per CTA), SM clocks")
print(f"  MMA warp:     total {avg[0]:7.0f}  waits: operands {avg[1]:6.0f}  softmax done {avg[2]:6.0f}  "
      f"accumulators free {avg[3]:6.0f}")
rows2 = [buf[16 * (160 + i):16 * (160 + i) + 16] for i in range(ctas)]
avg2 = [sum(r[k] for r in rows2) / ctas / items for k in range(2)]
print(f"                issuing (blocks while the tensor pipe's queue is full): score MMAs {avg2[0]:6.0f}  "
      f"dV / dK / dQ MMAs {avg2[1]:6.0f}")
print(f"  softmax warp: total {avg[4]:7.0f}  waits: row vectors {avg[5]:6.0f}  dS chunk free {avg[6]:6.0f}  "
      f"scores {avg[7]:6.0f}  dK/dV complete {avg[11]:6.0f}  dQ complete {avg[13]:6.0f}")
print(f"                sub-block bodies {avg[10]:7.0f} (TMEM loads {avg[8]:6.0f}, arithmetic {avg[9]:6.0f})  "
      f"dK/dV store {avg[12]:6.0f}  dQ store {avg[14]:6.0f}")
print(f"  row-vector warp: busy {avg[15]:6.0f}")
