"""compute-sanitizer --tool memcheck python tools/sanitize_*.py : tiny-model runs of the library-scheduled forward / one training step (grouped weight gradients, 128-bit AdamW) for out-of-bounds and misaligned accesses."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward
from cqa_crct_b200.optim import FusedAdamW
from cqa_crct_b200.spec import ModelConfig, synth_state_dict
from cqa_crct_b200.synthetic import default_params, make_batch
cfgp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'cqa_crct_b200', 'config', 'tiny.json')
cfg = ModelConfig(cfgp)
params = default_params(cfgp, device='cuda', max_seq_len=32, max_vis_features=12, L1=True)
m = VisualDialogEncoder(params)
m.load_state_dict({'bert_pretrained.' + k: v for k, v in synth_state_dict(cfg, 228, 3, 'mild').items()})
m.to('cuda').train()
opt = FusedAdamW(m)
gb = {k: v.cuda() for k, v in make_batch(6, 32, 12, cfg.v_feature_size, seed=3, vocab_size=cfg.vocab_size).items()}
for _ in range(2):
    opt.zero_grad()
    loss = glue_forward(m, gb, params)[0]
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print('loss', float(loss))
