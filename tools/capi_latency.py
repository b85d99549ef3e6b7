"""Inference latency / throughput of the library-scheduled forward (crct_forward, csrc/model.cu) against the Python host schedule
(VisualDialogEncoder.forward in evaluation mode) on the full model: wall time per call with a host sync (latency) and device time
of back-to-back calls (throughput).  usage: python tools/capi_latency.py [B ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cqa_crct_b200.capi import CModel
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward
from cqa_crct_b200.synthetic import default_params, make_batch
cfg = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
params = default_params(cfg, device='cuda', L1=True)
torch.manual_seed(0)
m = VisualDialogEncoder(params).to('cuda').eval()
cm = CModel(m)
for B in [int(x) for x in sys.argv[1:]] or [1, 32, 128, 512]:
    gb = {k: v.cuda() for k, v in make_batch(B, 124, 44, 1024, seed=7).items()}
    seq_len = torch.gather(gb['sep_indices'], 1, gb['hist_len'].view(-1, 1)).squeeze(1) + 1
    inp = {'tokens': gb['tokens'], 'segments': gb['segments'], 'loc': gb['loc'],
           'attention_mask': torch.arange(124, device='cuda').unsqueeze(0) < seq_len.unsqueeze(1), 'image_feat': gb['image_feat'],
           'image_loc': gb['image_loc'], 'image_target': gb['image_target'], 'image_mask': gb['image_mask'], 'R': gb['R']}
    fill = (float(seq_len.sum()) / (B * 124), float((gb['image_mask'] != 0).sum()) / (B * 44))
    m.row_fill_hint = fill

    def py():
        with torch.no_grad():
            return glue_forward(m, gb, params, evaluation=True)[4]

    def c():
        return cm.forward(inp, fill=fill)['logits']
    assert torch.equal(py(), c())
    def cg():
        return cm.forward_graphed(inp, fill=fill)['logits']
    assert torch.equal(py(), cg())
    for name, fn in (('python schedule', py), ('crct_forward  ', c), ('crct_forward in a CUDA graph', cg)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        n = 20
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
            torch.cuda.synchronize()
        lat = (time.perf_counter() - t0) / n * 1e3
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        dev = a.elapsed_time(b) / n
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        host = (time.perf_counter() - t0) / n * 1e3
        torch.cuda.synchronize()
        print(f'B={B:4d} {name}: latency {lat:7.3f} ms/call (synchronous), back-to-back {dev:7.3f} ms/call, host enqueue {host:7.3f} ms/call', flush=True)
