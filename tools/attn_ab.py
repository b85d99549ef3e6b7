"""Isolated A/B timing of the attention kernels: tcgen05 (attention_tc.cu) vs mma.sync (attention.cu), CRCT shapes at B = 80, packed rows
with the synthetic length distribution (text U[48..124], regions U[4..44]) and the padded layout; dropout 0.1 as in training."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ['CRCT_ATTN_TC_POLICY'] = '15'
from cqa_crct_b200 import _lib as L
DEV = 'cuda'
bf = lambda x: x.to(torch.bfloat16)
B = 80
g = torch.Generator().manual_seed(0)
lt = torch.randint(48, 125, (B,), generator=g)
lv = torch.randint(4, 45, (B,), generator=g)
def cu(l):
    c = torch.zeros(B + 1, dtype=torch.int32); c[1:] = l.cumsum(0); return c.to(DEV)
cases = {'text self 16x48': (16, 48, 124, 124, lt, lt), 'visual self 16x64': (16, 64, 44, 44, lv, lv),
         'co q=text k=vis 32x32': (32, 32, 124, 44, lt, lv), 'co q=vis k=text 32x32': (32, 32, 44, 124, lv, lt)}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, (nh, dh, Lq, Lk, lq, lk) in cases.items():
    H = nh * dh
    q = bf(torch.randn(B * Lq, 3 * H, device=DEV)); kv = q if Lq == Lk else bf(torch.randn(B * Lk, 3 * H, device=DEV))
    do = bf(torch.randn(B * Lq, H, device=DEV))
    for layout in ('packed', 'padded'):
        cq, ck = (cu(lq), cu(lk)) if layout == 'packed' else (None, None)
        mask = None if layout == 'packed' else torch.zeros(B, Lk, device=DEV)
        line = f'{name:24s} {layout:7s}'
        for impl in ('tc', 'mma'):
            if impl == 'mma':
                os.environ['CRCT_ATTN_LEGACY_NOW'] = '1'
            else:
                os.environ.pop('CRCT_ATTN_LEGACY_NOW', None)
            out = torch.zeros(B * Lq, H, device=DEV, dtype=torch.bfloat16); lse = torch.zeros(B, nh, Lq, device=DEV)
            dq, dkv = torch.zeros_like(q), torch.zeros_like(kv)
            kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H, dropout_p=0.1, seed=1, cu_q=cq, cu_k=ck)
            f = lambda: L.attn_fwd(q, kv[:, H:], kv[:, 2 * H:], mask, out, lse, **kw)
            bw = lambda: L.attn_bwd(q, kv[:, H:], kv[:, 2 * H:], mask, out, do, lse, dq, dkv[:, H:], dkv[:, 2 * H:], lddo=H, lddq=3 * H, lddk=3 * H, lddv=3 * H, **kw)
            for fn, tag in ((f, 'fwd'), (bw, 'bwd')):
                for _ in range(3):
                    fn()
                ts = []
                for _ in range(10):
                    flush.zero_()                      # cold L2
                    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); fn(); b_.record(); torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b_) * 1e3)
                line += f'  {impl} {tag} {sorted(ts)[len(ts) // 2]:6.1f} us'
        print(line, flush=True)
