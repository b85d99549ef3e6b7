"""In-process A/B of encoder parameters / environment switches on the real train step (same GPU, interleaved, CUDA-graph replay).
usage: python tools/ab_params.py name:key=value[,key=value][,ENV_VAR=value] ...      e.g.  base: grouped:group_wgrads=1"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.graph import GraphedTrainStep
from cqa_crct_b200.optim import FusedAdamW
from cqa_crct_b200.synthetic import default_params, make_batch
cfg = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
torch.manual_seed(0)
gb = {k: v.to('cuda') for k, v in make_batch(80, 124, 44, 1024, seed=5).items()}
steps = {}
for spec in sys.argv[1:]:
    name, _, kv = spec.partition(':')
    params = default_params(cfg, device='cuda', L1=True)
    env = {}
    for item in filter(None, kv.split(',')):
        k, v = item.split('=')
        if k.isupper():
            env[k] = v
        else:
            params[k] = type(params[k])(int(v)) if k in params and isinstance(params[k], (bool, int)) else (int(v) if v.lstrip('-').isdigit() else v)
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    m = VisualDialogEncoder(params).to('cuda').train()
    steps[name] = GraphedTrainStep(m, FusedAdamW(m), params, gb, warmup_steps=1)      # the graph is captured under this environment
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
for rnd in range(3):
    for name, g in steps.items():
        for _ in range(3):
            g.step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            g.step()
        b.record(); torch.cuda.synchronize()
        print(f'round {rnd} {name}: {a.elapsed_time(b) / 10:.3f} ms/step', flush=True)
