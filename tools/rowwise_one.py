"""Row-wise kernels at the text-stream shape (9920 x 768): LayerNorm fwd / bwd (with output dropout + bias grad), colsum.
With `time` as argv[1]: CUDA-event timings, cold (L2 flushed) and warm; otherwise a few launches for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L
dev = 'cuda'
rows, H = 9920, 768
z = torch.randn(rows, H, device=dev).bfloat16(); dy = torch.randn(rows, H, device=dev).bfloat16()
y = torch.empty_like(z); dz = torch.empty_like(z); dzm = torch.empty_like(z)
gamma = torch.ones(H, device=dev); beta = torch.zeros(H, device=dev)
mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
dg = torch.zeros(H, device=dev); db = torch.zeros(H, device=dev); dbias = torch.zeros(H, device=dev)
wide = torch.randn(rows, 3072, device=dev).bfloat16(); cs = torch.zeros(3072, device=dev)
L.SALT = torch.zeros(1, dtype=torch.int64, device=dev)
ops = {
    'ln_fwd 9920x768': lambda: L.layernorm_fwd(z, gamma, beta, y, mean, rstd),
    'ln_bwd 9920x768 (+dropout, dbias)': lambda: L.layernorm_bwd(dy, z, mean, rstd, gamma, dz, dg, db, dbias=dbias, dzm=dzm, p_out=0.1, seed_out=3),
    'ln_bwd 9920x768 (plain)': lambda: L.layernorm_bwd(dy, z, mean, rstd, gamma, dz, dg, db),
    'colsum 9920x768': lambda: L.colsum_bf16(dy, dg),
    'colsum 9920x3072': lambda: L.colsum_bf16(wide, cs),
}
if len(sys.argv) > 1 and sys.argv[1] == 'time':
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, fn in ops.items():
        for cold in (True, False):
            ts = []
            for r in range(12):
                if cold: flush.fill_(r)
                else: fn()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3)
            ts.sort()
            print(f'{name:36s} {"cold" if cold else "warm"} {ts[len(ts)//2]:7.1f} us', flush=True)
else:
    for _ in range(2):
        for fn in ops.values():
            fn()
    torch.cuda.synchronize()
