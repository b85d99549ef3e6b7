"""The oracle restatement against the golden vectors produced by the reference itself
(oracle/make_golden.py).  CPU only."""
import pytest
import torch

from oracle import crct_oracle as O
from tests.helpers import load_golden, golden_inputs, sample_idx

TRAIN = ['tiny_train_l1', 'tiny_train_smooth', 'tiny_ragged']
EVAL = ['tiny_eval', 'full_eval_b8', 'full_eval_b8_mild']


def _check_outputs(rec, out, tol=2e-5):
    assert torch.allclose(out['logits'].float(), rec['logits'], atol=tol, rtol=1e-4)
    assert torch.allclose(out['reg_pred'].float(), rec['reg_pred'], rtol=1e-4, atol=1e-3)
    assert torch.allclose(out['reg_loss'].float(), rec['reg_loss'], atol=tol)
    assert torch.allclose(out['reg_l1'].float(), rec['reg_l1'], atol=tol)
    assert torch.allclose(out['reg_dist'].float(), rec['reg_dist'], rtol=1e-4, atol=tol)
    assert tuple(out['reg_right']) == tuple(rec['reg_right'])
    assert torch.equal(out['logits'].argmax(1), rec['logits'].argmax(1))


@pytest.mark.parametrize('name', EVAL)
def test_eval_matches_reference_golden(name):
    rec = load_golden(name)
    _, cfg, sd, batch = golden_inputs(rec)
    out, _ = O.forward(sd, O.Config(cfg.__dict__), batch, train=False, l1=rec['l1'], keep_cache=False)
    _check_outputs(rec, out)


@pytest.mark.parametrize('name', TRAIN)
def test_train_matches_reference_golden(name):
    rec = load_golden(name)
    _, cfg, sd, batch = golden_inputs(rec)
    out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=rec['l1'], dtype=torch.float64)
    _check_outputs(rec, out)
    assert abs(float(out['loss']) - rec['loss']) < 1e-5
    assert abs(float(out['nsp_loss']) - rec['nsp_loss']) < 1e-5
    g = O.backward(cache)
    checked = 0
    for k, s in rec['grads'].items():
        if s['norm'] < 1e-9:          # key biases: softmax is shift-invariant, the reference holds rounding noise
            continue
        assert k in g, k
        f = g[k].double().flatten()
        assert abs(float(f.norm()) - s['norm']) <= 2e-4 * s['norm'] + 1e-9, k
        assert torch.allclose(f[sample_idx(k, f.numel())].float(), s['samples'], rtol=2e-3, atol=2e-4 * s['norm'] / max(1.0, f.numel() ** 0.5) + 1e-9), k
        checked += 1
    assert checked > 100
    dead = set(g) - set(rec['grads'])
    assert not dead, dead


def test_schedule_matches_reference_config():
    from cqa_crct_b200.spec import ModelConfig
    import os
    from tests.helpers import CONFIG_DIR
    cfg = ModelConfig(os.path.join(CONFIG_DIR, 'vilbert.json'))
    s = cfg.schedule()
    assert s[:7] == [('t', 0), ('t', 1), ('t', 2), ('t', 3), ('t', 4), ('t', 5), ('c', 0)]
    assert s[7:10] == [('v', 0), ('t', 6), ('c', 1)]
    assert s[-2:] == [('v', 5), ('t', 11)]
    assert len(s) == 24
