"""Per-shape timing of every GEMM launch of one real train step (CUDA events around each launch, warm caches)."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cqa_crct_b200 import _lib as L
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward
from cqa_crct_b200.optim import FusedAdamW
from cqa_crct_b200.synthetic import default_params, make_batch
cfg = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
params = default_params(cfg, device='cuda', L1=True)
torch.manual_seed(0)
m = VisualDialogEncoder(params).to('cuda').train()
opt = FusedAdamW(m)
gb = {k: v.to('cuda') for k, v in make_batch(80, 124, 44, 1024, seed=5).items()}
def step():
    opt.zero_grad(); glue_forward(m, gb, params)[0].backward(); opt.step()
for _ in range(3): step()
ev = []
orig = L.gemm
def timed(A, B, D, **kw):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); orig(A, B, D, **kw); b.record()
    ev.append((a, b, kw['M'], kw['N'], kw['K'], kw.get('a_major', 0), kw.get('b_major', 0), kw.get('epilogue', 0)))
L.gemm = timed
step(); torch.cuda.synchronize()
L.gemm = orig
agg = collections.defaultdict(lambda: [0, 0.0])
for a, b, M, N, K, am, bm, ep in ev:
    k = (M, N, K, am, bm, ep); agg[k][0] += 1; agg[k][1] += a.elapsed_time(b) * 1e3
tot = sum(v for c, v in agg.values())
print(f'{len(ev)} GEMMs, {tot/1e3:.2f} ms')
names = {0: 'bias', 1: 'gelu', 2: 'res', 3: 'mul', 4: 'f32'}
for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    M, N, K, am, bm, ep = k
    print(f'{v:8.0f} us n={c:3d} avg={v/c:7.1f} us {2*M*N*K*c/v/1e6:7.0f} TFLOP/s  M={M} N={N} K={K} a_mn={am} b_mn={bm} {names[ep]}')
