"""f4 — checkpoint layout (CRCT/train.py:91-130,282-291; CRCT/evaluation.py:30-53): the files this package writes load
into `torch.optim.AdamW` built the reference's way (`get_optimizer`, CRCT/utils.py:228-249) and vice versa."""
import json
import os

import pytest
import torch

from cqa_crct_b200 import checkpoint as C
from cqa_crct_b200.encoder import VisualDialogEncoder
from cqa_crct_b200.optim import FusedAdamW, WarmupLinearScheduleNonZero, _NO_DECAY
from cqa_crct_b200.synthetic import default_params
from tests.helpers import CONFIG_DIR

LANG = os.path.join(CONFIG_DIR, 'language_weights.json')


def reference_style_optimizer(model, lr, image_lr, wd):
    """Restatement of get_optimizer (CRCT/utils.py:228-249) over any module with the reference's parameter names."""
    lang = set(json.load(open(LANG)))
    groups = []
    for key, value in dict(model.named_parameters()).items():
        if value.requires_grad:
            g = {'params': [value], 'lr': lr if key in lang else image_lr,
                 'weight_decay': 0 if any(nd in key for nd in _NO_DECAY) else wd}
            groups.append(g)
    return torch.optim.AdamW(groups, lr=lr)


def build(seed=0):
    torch.manual_seed(seed)
    m = VisualDialogEncoder(default_params(os.path.join(CONFIG_DIR, 'tiny.json')))
    opt = FusedAdamW(m, lr=2e-5, image_lr=4e-5, weight_decay=0.01)
    sched = WarmupLinearScheduleNonZero(opt, warmup_steps=10, t_total=100, min_lr=1.3e-5)
    return m, opt, sched


def test_checkpoint_roundtrip_and_reference_layout(tmp_path):
    m, opt, sched = build()
    g = torch.Generator().manual_seed(3)
    opt.m.copy_(torch.randn(opt.n, generator=g))
    opt.v.copy_(torch.rand(opt.n, generator=g))
    opt.step_count = 7
    for _ in range(7):
        sched.step()
    path = C.save_checkpoint(str(tmp_path), 3, 7, m, opt, sched)
    assert os.path.basename(path) == 'plotqa_encoder_3_7.ckpt'                       # train.py:282
    payload = torch.load(path, weights_only=False)
    assert set(payload) == {'model_state_dict', 'scheduler_state_dict', 'optimizer_state_dict', 'iter_id'}   # train.py:287-289
    assert payload['iter_id'] == 7
    osd = payload['optimizer_state_dict']
    names = [k for k, _ in m.named_parameters()]
    assert len(osd['param_groups']) == len(names)                                    # one group per tensor, utils.py:236-247
    lang = set(json.load(open(LANG)))
    for grp, name in zip(osd['param_groups'], names):
        assert grp['initial_lr'] == (2e-5 if name in lang else 4e-5), name
        assert grp['weight_decay'] == (0 if any(nd in name for nd in _NO_DECAY) else 0.01), name
    live = {('bert_pretrained.' + p.name) for p in m.arena.spec if p.live}
    assert {names[i] for i in osd['state']} == live                                  # dead tensors never get state
    # the file is three flat buffers, not a thousand tensors
    assert os.path.getsize(path) < 4 * (m.arena.total + 2 * opt.n) + (1 << 20)

    # --- it loads into torch.optim.AdamW built the reference's way
    ref_opt = reference_style_optimizer(m, 2e-5, 4e-5, 0.01)
    ref_opt.load_state_dict(osd)
    p0 = dict(m.named_parameters())['bert_pretrained.bert.encoder.layer.0.output.dense.weight']
    o = m.arena.offsets['bert.encoder.layer.0.output.dense.weight']
    assert torch.equal(ref_opt.state[p0]['exp_avg'].reshape(-1), opt.m[o:o + p0.numel()])
    assert float(ref_opt.state[p0]['step']) == 7

    # --- and resumes here (train.py:104-127)
    m2, opt2, sched2 = build(seed=1)
    cont_epoch, start_iter, rest = C.resume(m2, opt2, sched2, path)
    assert (cont_epoch, start_iter, rest) == (4, 7, {})
    assert torch.equal(m2.arena.w32[:m2.arena.live_end], m.arena.w32[:m.arena.live_end])
    for p in m.arena.spec:                       # alignment padding between tensors is not part of the checkpoint
        if p.live:
            o = m.arena.offsets[p.name]
            assert torch.equal(opt2.m[o:o + p.numel], opt.m[o:o + p.numel]) and torch.equal(opt2.v[o:o + p.numel], opt.v[o:o + p.numel])
    assert opt2.step_count == 7
    assert sched2.last_epoch == sched.last_epoch and opt2.current_lrs() == opt.current_lrs()
    assert opt2.base_lr == opt.base_lr


def test_loads_a_torch_adamw_checkpoint():
    """A reference checkpoint's optimizer_state_dict (torch.optim.AdamW, one group per parameter, state only for
    tensors with gradients) fills the flat moment arenas."""
    m, opt, sched = build()
    ref_opt = reference_style_optimizer(m, 2e-5, 4e-5, 0.01)
    live = {('bert_pretrained.' + p.name) for p in m.arena.spec if p.live}
    g = torch.Generator().manual_seed(5)
    for k, p in m.named_parameters():
        p.grad = torch.randn(p.shape, generator=g) if k in live else None
    before = m.arena.w32.clone()
    ref_opt.step()
    ref_opt.step()
    m.arena.w32.copy_(before)
    opt.load_state_dict(ref_opt.state_dict())
    assert opt.step_count == 2
    for k, p in m.named_parameters():
        if k in live:
            name = k[len('bert_pretrained.'):]
            o = m.arena.offsets[name]
            assert torch.equal(opt.m[o:o + p.numel()], ref_opt.state[p]['exp_avg'].reshape(-1)), k
            assert torch.equal(opt.v[o:o + p.numel()], ref_opt.state[p]['exp_avg_sq'].reshape(-1)), k
    assert opt.base_lr == [2e-5, 2e-5, 4e-5, 4e-5]


def test_weights_only_load_filters_keys(tmp_path):
    """train.py:91-103 / evaluation.py:30-42: bare state dict or full payload, foreign keys dropped, >= 1 key needed."""
    m, _, _ = build()
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    sd['bert_pretrained.bert.encoder.layer.0.output.dense.bias'] += 1.0
    sd['something.else'] = torch.zeros(3)
    m2, _, _ = build(seed=2)
    assert C.load_weights(m2, sd) == len(sd) - 1
    assert torch.equal(m2.state_dict()['bert_pretrained.bert.encoder.layer.0.output.dense.bias'],
                       sd['bert_pretrained.bert.encoder.layer.0.output.dense.bias'])
    torch.save({'model_state_dict': sd}, tmp_path / 'crct.ckpt')
    m3, _, _ = build(seed=3)
    assert C.load_weights(m3, str(tmp_path / 'crct.ckpt')) == len(sd) - 1
    with pytest.raises(AssertionError):
        C.load_weights(m3, {'nothing': torch.zeros(1)})
    with pytest.raises(ValueError):
        C.epoch_of('model.ckpt')


def test_scheduler_state_matches_torch_lr_scheduler_layout():
    from torch.optim.lr_scheduler import LambdaLR
    m, opt, sched = build()
    keys = set(sched.state_dict())
    assert {'warmup_steps', 't_total', 'min_lr', 'base_lrs', 'last_epoch', '_step_count', '_last_lr'} <= keys    # utils.py:16-20 + _LRScheduler
    assert len(sched.state_dict()['base_lrs']) == len(list(m.named_parameters()))
    for _ in range(25):
        sched.step()
    sd = sched.state_dict()
    m2, opt2, sched2 = build()
    sched2.load_state_dict(sd)
    assert sched2.last_epoch == 25 and opt2.current_lrs() == opt.current_lrs()
