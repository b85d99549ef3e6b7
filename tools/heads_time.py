"""Event-timed head launches (crct_linear_f32_batched) of one forward + backward at B=80, cold and warm L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.synthetic import default_params, make_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
params = default_params(os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json'), device='cuda', L1=True)
m = VisualDialogEncoder(params).to('cuda').eval()
B, T, R = 80, 124, 44
gb = make_batch(B, T, R, 1024, seed=5)
t = torch.randn(B * T, 768, device='cuda').bfloat16(); v = torch.randn(B * R, 1024, device='cuda').bfloat16()
labels = gb['next_sentence_labels'].view(-1).cuda(); Rt = gb['R'].cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
m._bind_grads()
for cold in (True, False):
    rec = []
    orig = L.linear_f32_batched
    def timed(probs):
        if cold: flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); orig(probs); b.record(); rec.append((a, b, len(probs), max(p.K for p in probs), max(p.N for p in probs), max(p.M for p in probs)))
    for it in range(3):
        rec.clear()
        L.linear_f32_batched = timed
        logits, outs, scalars, s = m._heads_fwd(t, v, B, T, R, labels, Rt, 'L1_smooth', True)
        d_nsp = torch.ones(1, device='cuda'); d_reg = torch.full((B,), 1.0 / B, device='cuda')
        m._heads_bwd(s, d_nsp, d_reg, B, T, R)
        L.linear_f32_batched = orig
        torch.cuda.synchronize()
    print('cold L2' if cold else 'warm L2', 'total %.1f us' % sum(a.elapsed_time(b) * 1e3 for a, b, *_ in rec))
    print('  ', ' '.join('%d:%.0f' % (n, a.elapsed_time(b) * 1e3) for a, b, n, *_ in rec))
