"""compute-sanitizer --tool memcheck python tools/sanitize_*.py : tiny-model runs of the library-scheduled forward / one training step (grouped weight gradients, 128-bit AdamW) for out-of-bounds and misaligned accesses."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cqa_crct_b200.capi import CModel
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.evaluate import candidate_groups, expand_question_batch
from cqa_crct_b200.spec import ModelConfig, synth_state_dict
from cqa_crct_b200.synthetic import default_params, make_batch, make_question_batch
cfgp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'cqa_crct_b200', 'config', 'tiny.json')
cfg = ModelConfig(cfgp)
params = default_params(cfgp, device='cuda', max_seq_len=32, max_vis_features=12, L1=True)
m = VisualDialogEncoder(params)
m.load_state_dict({'bert_pretrained.' + k: v for k, v in synth_state_dict(cfg, 228, 3, 'mild').items()})
m.to('cuda').eval()
cm = CModel(m)
def inp(gb):
    seq_len = torch.gather(gb['sep_indices'], 1, gb['hist_len'].view(-1, 1)).squeeze(1) + 1
    am = torch.arange(gb['tokens'].shape[1], device='cuda').unsqueeze(0) < seq_len.unsqueeze(1)
    return {'tokens': gb['tokens'], 'segments': gb['segments'], 'loc': gb['loc'], 'attention_mask': am, 'image_feat': gb['image_feat'],
            'image_loc': gb['image_loc'], 'image_target': gb['image_target'], 'image_mask': gb['image_mask'], 'R': gb['R']}
gb = {k: v.cuda() for k, v in make_batch(5, 32, 12, cfg.v_feature_size, seed=3, vocab_size=cfg.vocab_size).items()}
print(cm.forward(inp(gb))['logits'].sum().item())
qb = make_question_batch(4, 32, 12, cfg.v_feature_size, seed=5, vocab_size=cfg.vocab_size, max_ans=5)
full = {k: v.cuda() for k, v in expand_question_batch(qb).items()}
qd = {k: v.cuda() for k, v in qb.items() if torch.is_tensor(v)}
i2 = inp(qd); i2['R'] = full['R']
print(cm.forward(i2, group=candidate_groups(qb['num_ans']).cuda())['logits'].sum().item())
torch.cuda.synchronize()
