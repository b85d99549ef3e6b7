"""Times the attention kernels at the CRCT shapes (B=80), rotating over 4 operand sets (> L2) — CUDA events, median of 20."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L

def run(name, B, nh, dh, Lq, Lk, cross):
    H = nh * dh
    sets = []
    for _ in range(4):
        qa = torch.randn(B * Lq, 3 * H, device='cuda').bfloat16()
        kb = torch.randn(B * Lk, 3 * H, device='cuda').bfloat16() if cross else qa
        sets.append((qa, kb, torch.empty(B * Lq, H, device='cuda', dtype=torch.bfloat16), torch.empty(B, nh, Lq, device='cuda'),
                     torch.randn(B * Lq, H, device='cuda').bfloat16(), torch.empty_like(qa), torch.empty_like(kb)))
    mask = torch.zeros(B, Lk, device='cuda')
    kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H)
    tf, tb = [], []
    for i in range(24):
        qa, kb, out, lse, dout, dqa, dkb = sets[i % 4]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        L.attn_fwd(qa, kb[:, H:], kb[:, 2 * H:], mask, out, lse, dropout_p=0.1, seed=5, **kw)
        e[1].record()
        L.attn_bwd(qa, kb[:, H:], kb[:, 2 * H:], mask, out, dout, lse, dqa, dkb[:, H:], dkb[:, 2 * H:], lddo=H, lddq=3 * H, lddk=3 * H,
                   lddv=3 * H, dropout_p=0.1, seed=5, **kw)
        e[2].record()
        torch.cuda.synchronize()
        if i >= 4:
            tf.append(e[0].elapsed_time(e[1]) * 1e3); tb.append(e[1].elapsed_time(e[2]) * 1e3)
    print(f'{name:22s} fwd {statistics.median(tf):7.1f} us   bwd {statistics.median(tb):7.1f} us')

run('text 16x48 124x124', 80, 16, 48, 124, 124, False)
run('co t->v 32x32 124x44', 80, 32, 32, 124, 44, True)
run('co v->t 32x32 44x124', 80, 32, 32, 44, 124, True)
run('vis 16x64 44x44', 80, 16, 64, 44, 44, False)
