"""Not a pytest file: prints the role timeline of attn_fwd_tc2_kernel (library built with -DVITK_ATTN_TRACE)."""
import sys, torch
sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import vitk
B,N,H=256,197,12
qkv=(torch.randn(B*N,3*H*64,device='cuda')*1.5).bfloat16()
for _ in range(3):
    ctx,lse=vitk.ops.attention(qkv,B,N,H,return_lse=True)
torch.cuda.synchronize()
tr=lse.flatten()[:512].view(torch.int64).cpu().reshape(4,8,8)
t0=tr[1,0,0].item()
names={0:'MMA',1:'SM0',2:'SM1',3:'EPI'}
for it in range(6):
    print('item',it)
    print('  SM0 start %6d end %6d'%(tr[1,it,0]-t0,tr[1,it,1]-t0), '  SM1 start %6d end %6d'%(tr[2,it,0]-t0,tr[2,it,1]-t0))
    print('  SM0 chunks', [(tr[1,it,k]-t0).item() for k in range(2,8)], ' SM1 chunks', [(tr[2,it,k]-t0).item() for k in range(2,8)])
    print('  MMA t0: p_full %6d o_free %6d pv_issued %6d s_issued %6d | t1: p_full %6d o_free %6d pv_issued %6d s_issued %6d'%tuple((tr[0,it,k]-t0).item() for k in (0,1,3,2,4,5,7,6)))
    print('  EPI t0: o_full %6d o_free %6d stored %6d | t1: o_full %6d o_free %6d stored %6d'%tuple((tr[3,it,k]-t0).item() for k in (0,1,2,4,5,6)))
