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


def _reference_eval_lines(lo, hi):
    """Source lines [lo, hi] (1-based) of the reference's evaluation.py, de-indented — executed where they lie, not copied:
    the module itself cannot be imported offline (fig_dataloader -> pytorch_transformers, matplotlib)."""
    import textwrap
    path = os.path.join(ref_shim.REFERENCE_ROOT, 'CRCT', 'evaluation.py')
    src = open(path).read().splitlines()
    return textwrap.dedent('\n'.join(src[lo - 1:hi]))


@pytest.mark.parametrize('seed,force', [(0, False), (1, False), (2, True)])
def test_eval_oracle_against_the_reference_selection_lines(seed, force, monkeypatch):
    """oracle/eval_oracle.py vs the reference's own answer-selection block (evaluation.py:274-312) and
    `reduce_total_acc` (:494-525) run on the same arrays."""
    import numpy as np
    from oracle.eval_oracle import select_and_score
    from cqa_crct_b200.synthetic import make_question_batch
    qb = make_question_batch(23, 16, 4, 8, seed=seed, vocab_size=2048, max_ans=40)
    if force:
        qb['gt_id'] = qb['gt_id'].clamp(min=0)
    N = int(qb['num_ans'].sum())
    g = torch.Generator().manual_seed(seed)
    scores = torch.randn(N, 2, generator=g) * 3
    scores[5] = scores[4]                                     # an exact tie inside a question: first maximum wins
    reg_pred, reg_dist, reg_l1 = torch.randn(N, generator=g), torch.rand(N, generator=g) * 0.1, torch.rand(N, generator=g) * 0.02
    mine = select_and_score(scores, reg_pred, reg_dist, reg_l1, qb['num_ans'], qb['gt_id'], qb['needs_reg'], qb['tolerance_margin'],
                            force_gt=force)
    ns = {'torch': torch, 'np': np, 'params': {'binary_answers': False, 'qa_file': 'qa_pairs_REGS.json' if force else 'qa_pairs.json'},
          'batch': {'num_ans': qb['num_ans'], 'gt_id': qb['gt_id'].view(-1, 1), 'tokens': qb['tokens'], 'id': qb['id'].view(-1, 1),
                    'needs_reg': qb['needs_reg'].view(-1, 1), 'tolerance_margin': qb['tolerance_margin'].view(-1, 1),
                    'next_sentence_labels': qb['next_sentence_labels']},
          'output': torch.softmax(scores, 1)[:, 0], 'reg_output': reg_pred, 'reg_loss_lst': reg_dist, 'reg_t_loss_lst': reg_l1}
    exec(_reference_eval_lines(274, 312), ns)
    assert torch.equal(ns['answers'], mine['answers'])
    assert torch.equal(ns['reg_answers_output'], mine['reg_output'])
    assert torch.equal(ns['reg_answers_loss'], mine['reg_loss']) and torch.equal(ns['reg_answers_t_loss'], mine['reg_t_loss'])
    flags = torch.stack([ns['nsp_right'], ns['reg_right'], ns['reg_t_right'], ns['correct_answers'], ns['correct_answers_t_loss']], 1)
    assert torch.equal(flags.to(torch.uint8), mine['flags'])
    # reduce_total_acc allocates with .cuda() and all-reduces: identity / no-op on this CPU-only check
    monkeypatch.setattr(torch.Tensor, 'cuda', lambda self, *a, **k: self)
    fn_ns = {'torch': torch, 'dist': type('D', (), {'all_reduce': staticmethod(lambda *a, **k: None)})}
    exec(_reference_eval_lines(494, 525), fn_ns)
    total = fn_ns['reduce_total_acc'](torch.zeros(6, 2).double(), ns['needs_regression'], ns['nsp_right'], ns['reg_right'],
                                      ns['reg_t_right'], None)
    assert torch.equal(total, mine['total_correct'])


def test_train_flags_are_the_reference_flags_with_the_same_defaults():
    """cqa_crct_b200.train.read_command_line vs CRCT/options.py:9-81: every flag kept from the reference has the reference's
    type and default (so existing command lines keep their meaning); the only additions are the synthetic-data / launch
    flags, and the only changed defaults are the ones the synthetic data replaces."""
    import argparse
    import re
    from cqa_crct_b200 import train as T
    src = open(os.path.join(ref_shim.REFERENCE_ROOT, 'CRCT', 'options.py')).read()
    body = src[src.index('parser = argparse.ArgumentParser'):src.index('try:')]
    ns = {'argparse': argparse, 'sys': __import__('sys')}
    exec(re.sub(r'^    ', '', body, flags=re.M), ns)                    # the reference's own add_argument calls
    ref = {a.dest: a for a in ns['parser']._actions if a.dest != 'help'}
    captured = {}
    orig = argparse.ArgumentParser.parse_args

    def grab(self, args=None, namespace=None):
        captured['actions'] = {a.dest: a for a in self._actions if a.dest != 'help'}
        return orig(self, args, namespace)
    argparse.ArgumentParser.parse_args = grab
    try:
        T.read_command_line([])
    finally:
        argparse.ArgumentParser.parse_args = orig
    mine = captured['actions']
    added = set(mine) - set(ref)
    assert added == {'max_vis_features', 'iters_per_epoch', 'eval_questions', 'graph'}, added
    changed = {}
    for k, a in mine.items():
        if k in ref:
            assert type(a) is type(ref[k]), k                           # store / store_true
            if a.default != ref[k].default:
                changed[k] = (ref[k].default, a.default)
    # dataset-dependent defaults: config path, sequence length of config/plotqa.json, candidate chunk, qa file, class count
    assert set(changed) <= {'model_config', 'max_seq_len', 'eval_batch_size', 'qa_file', 'categories'}, changed


def _edge_batches(cfg):
    from cqa_crct_b200.synthetic import make_batch
    mk = lambda B, seed: make_batch(B, 24, 9, cfg.v_feature_size, seed=seed, vocab_size=cfg.vocab_size)
    b1 = mk(1, 3)
    none = mk(5, 4); none['R'][:, 1] = 0; none['needs_reg'][:] = False
    allr = mk(5, 5); allr['R'][:, 1] = 1
    img = mk(4, 6); img['image_mask'][:, 1:] = 0
    zero = mk(4, 8); zero['R'][0] = torch.tensor([0.0, 1.0, 0.01, 1.0])
    return {'single sequence': b1, 'no regression rows': none, 'all regression rows': allr, 'only the <IMG> region visible': img,
            'zero regression target': zero}


@pytest.mark.parametrize('l1', [True, False])
def test_edge_case_batches_against_reference(l1):
    """Degenerate batches the reference handles in its loss bookkeeping (vilbert.py:1586-1657): B = 1, no row / every row
    needing regression (empty boolean gather, regressor.py:36-42 on [0, H]), a fully masked visual stream but the <IMG>
    token, a zero target (the 0/0 rule of the relative distance, :1632-1636)."""
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, seed=5, style='trained')
    params = default_params(cfg_path, max_seq_len=24, max_vis_features=9, L1=l1)
    enc = ref_shim.RefEncoder(cfg_path, params)
    enc.module.bert_pretrained.load_state_dict(sd)
    enc.module.eval()
    for name, batch in _edge_batches(cfg).items():
        loss, _, nsp, _, scores, reg, _ = enc.glue_forward(enc.module, batch, params)
        out, _ = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=l1, dtype=torch.float64)
        assert abs(float(loss) - float(out['loss'])) < 1e-6, name
        assert rel_err(out['logits'], scores.detach()) < 1e-5, name
        for k, i in (('reg_pred', 0), ('reg_loss', 1), ('reg_l1', 2), ('reg_dist', 4)):
            assert float((out[k].float() - reg[i].detach()).abs().max()) < 2e-5 * max(1.0, float(reg[i].detach().abs().max())), (name, k)
        assert (int(reg[3][0]), int(reg[3][1])) == tuple(out['reg_right']), name


def test_all_labels_ignored_is_a_documented_deviation():
    """CrossEntropyLoss(ignore_index=-1) over a batch whose labels are ALL -1 is 0/0 = NaN in the reference
    (vilbert.py:1513,1655-1657); the data loader never produces that batch (labels are 0/1, fig_dataloader.py:53-54).
    This implementation returns nsp_loss = 0 there (deviations) — pinned here so it stays a decision."""
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, seed=5, style='trained')
    params = default_params(cfg_path, max_seq_len=24, max_vis_features=9)
    enc = ref_shim.RefEncoder(cfg_path, params)
    enc.module.bert_pretrained.load_state_dict(sd)
    enc.module.eval()
    batch = make_batch(4, 24, 9, cfg.v_feature_size, seed=7, vocab_size=cfg.vocab_size)
    batch['next_sentence_labels'][:] = -1
    loss = enc.glue_forward(enc.module, batch, params)[0]
    out, _ = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=True, dtype=torch.float64)
    assert torch.isnan(loss) and float(out['nsp_loss']) == 0.0 and torch.isfinite(out['loss'])
