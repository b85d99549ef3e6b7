"""One forward + one backward launch of the text self-attention shape (B = 80, packed) for ncu."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['CRCT_ATTN_TC_POLICY'] = '15'
from cqa_crct_b200 import _lib as L
DEV = 'cuda'
B, nh, dh, Lq = 80, 16, 48, 124
H = nh * dh
g = torch.Generator().manual_seed(0)
lt = torch.randint(48, 125, (B,), generator=g)
cu = torch.zeros(B + 1, dtype=torch.int32); cu[1:] = lt.cumsum(0); cu = cu.to(DEV)
q = torch.randn(B * Lq, 3 * H, device=DEV).to(torch.bfloat16)
do = torch.randn(B * Lq, H, device=DEV).to(torch.bfloat16)
out = torch.zeros(B * Lq, H, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, nh, Lq, device=DEV)
dq = torch.zeros_like(q)
kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lq, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H, dropout_p=0.1, seed=1, cu_q=cu, cu_k=cu)
for _ in range(2):
    L.attn_fwd(q, q[:, H:], q[:, 2 * H:], None, out, lse, **kw)
    L.attn_bwd(q, q[:, H:], q[:, 2 * H:], None, out, do, lse, dq, dq[:, H:], dq[:, 2 * H:], lddo=H, lddq=3 * H, lddk=3 * H, lddv=3 * H, **kw)
torch.cuda.synchronize()
