"""One process, few steps of the bench workload — meant to be wrapped by ncu for the launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file gpurun_out/launches.csv python tools/step_profile.py
Also prints the host-side enqueue time per step (is the step launch-bound?)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from cqa_crct_b200 import _lib as L  # noqa: E402
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward  # noqa: E402
from cqa_crct_b200.optim import FusedAdamW  # noqa: E402
from cqa_crct_b200.synthetic import default_params, make_batch  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 80
cfg = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
from cqa_crct_b200.graph import fill_fractions  # noqa: E402
hb = make_batch(B, 124, 44, 1024, seed=5)
params = default_params(cfg, device='cuda', L1=True, row_fill_hint=fill_fractions(hb))      # tile shapes as in the captured bench step
torch.manual_seed(0)
m = VisualDialogEncoder(params).to('cuda').train()
m.overlap_streams = os.environ.get('CRCT_PROFILE_LANES', '0') == '1'       # one stream: ncu serialises the launches anyway
opt = FusedAdamW(m)
gb = {k: v.to('cuda') for k, v in hb.items()}
for i in range(steps):
    torch.cuda.synchronize()
    l0, t0 = L.LAUNCHES, time.perf_counter()
    opt.zero_grad()
    loss = glue_forward(m, gb, params)[0]
    t1 = time.perf_counter()
    loss.backward()
    opt.step()
    t2 = time.perf_counter()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f'step {i}: launches={L.LAUNCHES - l0} host enqueue fwd={1e3 * (t1 - t0):.2f} ms bwd+opt={1e3 * (t2 - t1):.2f} ms, wall incl. sync={1e3 * (t3 - t0):.2f} ms', flush=True)
