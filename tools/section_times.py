"""Warm, in-step time of every kernel family of one train step: CUDA events recorded as GRAPH NODES around each library call of a captured
single-stream replay of the step (no host in the loop, L2 state as in the real step; the ncu launch list is cold-cache).  Sections:
GEMM, grouped weight gradients, attention fwd / bwd, LayerNorm fwd / bwd, column sums, heads fwd / bwd (whole sections), embeddings,
AdamW.  usage: python tools/section_times.py [B]"""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cqa_crct_b200 import _lib as L
from cqa_crct_b200 import encoder as E
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.graph import fill_fractions
from cqa_crct_b200.optim import FusedAdamW
from cqa_crct_b200.synthetic import default_params, make_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 80
cfg = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
hb = make_batch(B, 124, 44, 1024, seed=5)
params = default_params(cfg, device='cuda', L1=True, row_fill_hint=fill_fractions(hb))
torch.manual_seed(0)
enc = VisualDialogEncoder(params).to('cuda').train()
enc.overlap_streams = False
opt = FusedAdamW(enc)
opt.enable_device_scalars()
gb = {k: v.to('cuda') for k, v in hb.items()}
pairs = {}


def wrap(section, fn):
    def timed(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        pairs.setdefault(section, []).append((e0, e1))
        return r
    return timed


def step():
    enc.zero_grad()
    for _ in enc.train_step_stages(gb, 1.0, 1.0):
        pass
    opt.step_captured()


for _ in range(2):                      # eager warm-up (allocator, lazy attributes)
    opt.push_device_scalars()
    step()
torch.cuda.synchronize()
names = {'gemm': 'GEMM (fwd / dgrad)', 'gemm_wgrad_grouped': 'grouped weight gradients', 'attn_fwd': 'attention fwd', 'attn_bwd': 'attention bwd',
         'layernorm_fwd': 'LayerNorm fwd', 'layernorm_bwd': 'LayerNorm bwd (dz)', 'layernorm_bwd_params': 'LayerNorm bwd (column sums)',
         'colsum_bf16': 'bias-gradient column sums', 'embed_text_fwd': 'embeddings', 'embed_vis_fwd': 'embeddings', 'embed_text_bwd': 'embeddings',
         'embed_vis_bwd': 'embeddings', 'softmax_rows': 'embeddings', 'adamw': 'AdamW', 'row_map': 'row maps',
         'linear_f32_batched': 'heads: batched fp32 linear launches', 'layernorm_rows_f32': 'heads: other', 'pool_mul_fwd': 'heads: other',
         'pool_mul_bwd': 'heads: other', 'hybrid_loss': 'heads: other', 'scale_rows': 'heads: other', 'scatter_rows_f32': 'heads: other',
         'fill_zero': 'memsets'}
saved = {n: getattr(L, n) for n in names}
hf, hb_ = enc._heads_fwd, enc._heads_bwd
try:
    for n, sec in names.items():
        setattr(L, n, wrap(sec, saved[n]))
    enc._heads_fwd, enc._heads_bwd = wrap('heads fwd (section)', hf), wrap('heads bwd (section)', hb_)
    nulls = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        g.capture_begin()
        t0, t1 = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
        t0.record()
        step()
        t1.record()
        for _ in range(64):
            a, b = torch.cuda.Event(enable_timing=True, external=True), torch.cuda.Event(enable_timing=True, external=True)
            a.record(); b.record(); nulls.append((a, b))
        g.capture_end()
        for _ in range(3):
            opt.push_device_scalars()
            g.replay()
    torch.cuda.synchronize()
finally:
    for n, f in saved.items():
        setattr(L, n, f)
    enc._heads_fwd, enc._heads_bwd = hf, hb_
ovh = statistics.median(a.elapsed_time(b) for a, b in nulls)
total = t0.elapsed_time(t1)
print(f'# one train step B = {B}, ONE stream, captured graph with event nodes (adds ~{ovh * 1e3:.1f} us per pair, removed): {total:.2f} ms incl. event nodes')
rows = []
for sec, ps in pairs.items():
    raw = sum(a.elapsed_time(b) for a, b in ps)
    nested = sec.startswith('heads')
    ms = raw - ovh * len(ps)
    rows.append((ms, sec, len(ps)))
acc = 0.0
for ms, sec, n in sorted(rows, reverse=True):
    print(f'{ms:8.3f} ms  {n:4d} calls  {sec}')
    if not sec.startswith('heads'):
        acc += ms
    if os.environ.get('SECTION_DETAIL') and sec.startswith(os.environ['SECTION_DETAIL']):
        print('           per call (us):', ' '.join(f'{(a.elapsed_time(b) - ovh) * 1e3:.1f}' for a, b in pairs[sec]))
print(f'{acc:8.3f} ms  sum of the kernel families (heads sections contain their own launches only; GEMM etc. do not run inside them)')
