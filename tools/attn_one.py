"""A few attention launches at the text-layer shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L
B, nh, dh, T = 80, 16, 48, 124
H = nh * dh
qkv = torch.randn(B * T, 3 * H, device='cuda').bfloat16()
mask = torch.zeros(B, T, device='cuda')
out = torch.empty(B * T, H, device='cuda', dtype=torch.bfloat16)
lse = torch.empty(B, nh, T, device='cuda')
dout = torch.randn(B * T, H, device='cuda').bfloat16()
dqkv = torch.empty_like(qkv)
kw = dict(B=B, nh=nh, dh=dh, Lq=T, Lk=T, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H)
for _ in range(3):
    L.attn_fwd(qkv, qkv[:, H:], qkv[:, 2 * H:], mask, out, lse, dropout_p=0.1, seed=5, **kw)
    L.attn_bwd(qkv, qkv[:, H:], qkv[:, 2 * H:], mask, out, dout, lse, dqkv, dqkv[:, H:], dqkv[:, 2 * H:], lddo=H, lddq=3 * H, lddk=3 * H, lddv=3 * H, dropout_p=0.1, seed=5, **kw)
torch.cuda.synchronize()
