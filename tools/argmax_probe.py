"""GPU probe: per-question candidate margins (fp32 oracle) vs the CUDA path's probability error on the full model."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.evaluate import evaluate_batch, expand_question_batch
from cqa_crct_b200.spec import ModelConfig, synth_state_dict
from cqa_crct_b200.synthetic import default_params, make_question_batch
from oracle import crct_oracle as O
cfg_path = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
cfg = ModelConfig(cfg_path)
params = default_params(cfg_path, device='cuda', max_seq_len=124, max_vis_features=44, L1=True)
qb = make_question_batch(16, 124, 44, cfg.v_feature_size, seed=77, total=512, distinct=True)
full = expand_question_batch(qb)
for style in ('mild', 'trained'):
    sd = synth_state_dict(cfg, 228, 1, style)
    m = VisualDialogEncoder(params)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to('cuda').eval()
    out = evaluate_batch(m, qb, params, eval_batch_size=512)
    with torch.no_grad():
        o, _ = O.forward(sd, O.Config(cfg.__dict__), full, train=False, l1=True, keep_cache=False)
    p = torch.softmax(o['logits'], 1)[:, 0]
    pc = out['prob'].cpu()
    print(style, 'logit scale %.3f  logits err %.2e  prob err max %.2e' % (float(o['logits'].abs().max()), float((out['logits'].cpu() - o['logits']).abs().max() / o['logits'].abs().max()), float((pc - p).abs().max())))
    off = 0
    for q, n in enumerate(qb['num_ans'].tolist()):
        top = torch.sort(p[off:off + n], descending=True)
        d = (pc[off:off + n] - p[off:off + n])
        print('  q%02d n=%d margin %.2e spread %.2e  common-mode err %.2e  differential err %.2e  same=%s regret %.1e' % (
            q, n, float(top.values[0] - top.values[1]), float(top.values[0] - top.values[-1]), float(d.mean()), float((d - d.mean()).abs().max()),
            int(out['answers'][q]) == int(top.indices[0]), float(top.values[0] - p[off + int(out['answers'][q])])))
        off += n
    del m
