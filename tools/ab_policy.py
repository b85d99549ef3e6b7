"""In-process A/B of GEMM tile/cluster policies on the real train step (same GPU, interleaved, CUDA-graph replay)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.graph import GraphedTrainStep
from cqa_crct_b200.optim import FusedAdamW
from cqa_crct_b200.synthetic import default_params, make_batch
cfg = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
params = default_params(cfg, device='cuda', L1=True)
torch.manual_seed(0)
gb = {k: v.to('cuda') for k, v in make_batch(80, 124, 44, 1024, seed=5).items()}
steps = {}
for pol in sys.argv[1:] or ['0', '1', '2', '4']:
    os.environ['CRCT_GEMM_PAIR_POLICY'] = pol
    m = VisualDialogEncoder(params).to('cuda').train()
    steps[pol] = GraphedTrainStep(m, FusedAdamW(m), params, gb, warmup_steps=1)
for rnd in range(3):
    for pol, g in steps.items():
        for _ in range(3):
            g.step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.step()
        b.record(); torch.cuda.synchronize()
        print(f'round {rnd} policy {pol}: {a.elapsed_time(b) / 10:.3f} ms/step', flush=True)
