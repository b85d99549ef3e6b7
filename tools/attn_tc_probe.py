"""Bring-up probe for the tcgen05 attention: forward, then backward, tiny to full shapes, against the mma.sync kernels."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['CRCT_ATTN_TC_POLICY'] = '15'
from cqa_crct_b200 import _lib as L
DEV = 'cuda'
bf = lambda x: x.to(torch.bfloat16)
rel = lambda a, b: float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
which = sys.argv[1] if len(sys.argv) > 1 else 'both'
SH = [(1, 1, 64, 128, 128), (1, 2, 48, 124, 124), (2, 2, 32, 124, 44), (2, 2, 32, 44, 124), (3, 16, 64, 44, 44), (80, 16, 48, 124, 124)]
if len(sys.argv) > 2:
    SH = [tuple(int(x) for x in sys.argv[2].split(','))]
for (B, nh, dh, Lq, Lk) in SH:
    torch.manual_seed(1)
    H = nh * dh
    q, k, v, do = (bf(torch.randn(B * n, H, device=DEV)) for n in (Lq, Lk, Lk, Lq))
    mask = torch.zeros(B, Lk, device=DEV)
    mask[:, Lk - 3:] = -10000.0
    kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H)
    res = {}
    for impl in ('mma', 'tc'):
        if impl == 'mma':
            os.environ['CRCT_ATTN_LEGACY_NOW'] = '1'
        else:
            os.environ.pop('CRCT_ATTN_LEGACY_NOW', None)
        out = torch.zeros(B * Lq, H, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, nh, Lq, device=DEV)
        L.attn_fwd(q, k, v, mask, out, lse, **kw)
        torch.cuda.synchronize()
        print((B, nh, dh, Lq, Lk), impl, 'fwd done', flush=True)
        dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
        if which != 'fwd':
            L.attn_bwd(q, k, v, mask, out, do, lse, dq, dk, dv, lddo=H, lddq=H, lddk=H, lddv=H, **kw)
            torch.cuda.synchronize()
            print((B, nh, dh, Lq, Lk), impl, 'bwd done', flush=True)
        res[impl] = (out, lse, dq, dk, dv)
    print('   err out %.2e lse %.2e dq %.2e dk %.2e dv %.2e' % tuple(rel(a.float(), b.float()) for a, b in zip(res['tc'], res['mma'])), flush=True)
