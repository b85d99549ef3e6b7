"""Live check of the oracle against the UNMODIFIED reference (only where /root/reference exists,
i.e. in the build container; skipped on the GPU box)."""
import os

import pytest
import torch

from oracle import ref_shim, crct_oracle as O
from cqa_crct_b200.spec import ModelConfig, synth_state_dict, param_spec
from cqa_crct_b200.synthetic import make_batch, default_params
from tests.helpers import CONFIG_DIR, rel_err

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason='reference checkout not present')


@pytest.mark.parametrize('l1', [True, False])
def test_full_gradients_against_reference_autograd(l1):
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    cfg = ModelConfig(cfg_path)
    params = default_params(cfg_path, max_seq_len=24, max_vis_features=9, L1=l1)
    enc = ref_shim.RefEncoder(cfg_path, params)
    sd = synth_state_dict(cfg, 228, seed=5, style='trained')
    enc.module.bert_pretrained.load_state_dict(sd)
    enc.module.eval()
    batch = make_batch(7, 24, 9, cfg.v_feature_size, seed=3, vocab_size=cfg.vocab_size)
    loss, _, nsp, _, scores, reg, _ = enc.glue_forward(enc.module, batch, params)
    loss.backward()
    out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=l1, dtype=torch.float64)
    g = O.backward(cache)
    assert abs(float(loss) - float(out['loss'])) < 1e-6
    assert rel_err(out['logits'], scores.detach()) < 1e-5
    for k, p in enc.module.bert_pretrained.named_parameters():
        if p.grad is None:
            assert k not in g
            continue
        if p.grad.norm() < 1e-9:
            continue
        # the reference runs in fp32: cancellation noise dominates very small gradients
        assert rel_err(g[k], p.grad) < (1e-4 if p.grad.norm() > 1e-3 else 5e-3), k


def test_param_spec_is_the_reference_named_parameters():
    for f in ('tiny.json', 'vilbert.json'):
        cfg_path = os.path.join(CONFIG_DIR, f)
        m = ref_shim.build_reference_model(cfg_path, default_params(cfg_path))
        ref = [(k, tuple(v.shape)) for k, v in m.named_parameters()]
        mine = [(p.name, p.shape) for p in param_spec(ModelConfig(cfg_path))]
        assert ref == mine
        dead = {p.name for p in param_spec(ModelConfig(cfg_path)) if not p.live}
        assert len(dead) == (36 if f == 'vilbert.json' else 20)
