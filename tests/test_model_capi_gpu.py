"""Whole-model C entry points (include/crct_b200.h: crct_create / crct_bind_params / crct_workspace_bytes / crct_forward,
csrc/model.cu) against the Python host schedule: the inference forward the library schedules itself runs the same kernels in the
same order as `VisualDialogEncoder.forward` in evaluation mode (CRCT/backbone/encoder_decorator.py:73-158 with evaluation=True),
so every output must agree BIT FOR BIT — on the tiny and the full configuration, on the reference's replicated layout and on
de-duplicated question batches (f3), and against the golden vectors produced by the reference itself."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import _lib as L                                                         # noqa: E402
from cqa_crct_b200.capi import CModel                                                       # noqa: E402
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward                         # noqa: E402
from cqa_crct_b200.evaluate import candidate_groups, expand_question_batch                  # noqa: E402
from cqa_crct_b200.spec import ModelConfig, synth_state_dict                                # noqa: E402
from cqa_crct_b200.synthetic import default_params, make_batch, make_question_batch         # noqa: E402
from tests.helpers import CONFIG_DIR, golden_inputs, load_golden                            # noqa: E402


def _model(config, T, R, seed=3, style='trained'):
    cfg_path = os.path.join(CONFIG_DIR, config)
    cfg = ModelConfig(cfg_path)
    params = default_params(cfg_path, device='cuda', max_seq_len=T, max_vis_features=R, L1=True)
    m = VisualDialogEncoder(params)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in synth_state_dict(cfg, 228, seed, style).items()})
    m.to('cuda').eval()
    return m, params, cfg


def _c_inputs(gb):
    seq_len = torch.gather(gb['sep_indices'], 1, gb['hist_len'].view(-1, 1)).squeeze(1) + 1          # encoder_decorator.py:118-119
    amask = torch.arange(gb['tokens'].shape[1], device=seq_len.device).unsqueeze(0) < seq_len.unsqueeze(1)
    return {'tokens': gb['tokens'], 'segments': gb['segments'], 'loc': gb['loc'], 'attention_mask': amask, 'image_feat': gb['image_feat'],
            'image_loc': gb['image_loc'], 'image_target': gb['image_target'], 'image_mask': gb['image_mask'], 'R': gb['R']}


def _same(out, scores, reg):
    assert torch.equal(out['logits'], scores)
    assert torch.equal(out['reg_pred'], reg[0]) and torch.equal(out['reg_loss'], reg[1]) and torch.equal(out['reg_l1'], reg[2])
    assert torch.equal(out['reg_dist'], reg[4])
    assert float(out['scalars'][3]) == float(reg[3][0]) and float(out['scalars'][4]) == float(reg[3][1])


@pytest.mark.parametrize('config,B,T,R', [('tiny.json', 6, 32, 12), ('tiny.json', 1, 16, 4), ('vilbert.json', 8, 124, 44), ('vilbert.json', 80, 124, 44)])
def test_c_forward_equals_python_forward_bit_for_bit(config, B, T, R):
    m, params, cfg = _model(config, T, R)
    gb = {k: v.cuda() for k, v in make_batch(B, T, R, cfg.v_feature_size, seed=17, vocab_size=cfg.vocab_size).items()}
    with torch.no_grad():
        _, _, _, _, scores, reg = glue_forward(m, gb, params, evaluation=True)
    cm = CModel(m)
    out = cm.forward(_c_inputs(gb))
    _same(out, scores, reg)
    out2 = cm.forward(_c_inputs(gb), fill=(0.7, 0.5))             # the fill hints steer tile shapes only
    _same(out2, scores, reg)
    if B <= 8:                                                    # crct_forward is capturable: replayed from a CUDA graph, other batch, same shapes
        gb2 = {k: v.cuda() for k, v in make_batch(B, T, R, cfg.v_feature_size, seed=18, vocab_size=cfg.vocab_size).items()}
        cm.forward_graphed(_c_inputs(gb))
        out3 = cm.forward_graphed(_c_inputs(gb2))
        with torch.no_grad():
            _, _, _, _, scores2, reg2 = glue_forward(m, gb2, params, evaluation=True)
        _same(out3, scores2, reg2)


@pytest.mark.parametrize('config,Q,T,R', [('tiny.json', 9, 32, 12), ('vilbert.json', 4, 124, 44)])
def test_c_forward_question_batches(config, Q, T, R):
    """f3: visual tensors once per question, `group` maps candidates to questions — equal to the replicated layout through Python."""
    m, params, cfg = _model(config, T, R)
    qb = make_question_batch(Q, T, R, cfg.v_feature_size, seed=5, vocab_size=cfg.vocab_size, max_ans=9)
    full = {k: v.cuda() for k, v in expand_question_batch(qb).items()}
    with torch.no_grad():
        _, _, _, _, scores, reg = glue_forward(m, full, params, evaluation=True)
    qd = {k: v.cuda() for k, v in qb.items() if torch.is_tensor(v)}
    group = candidate_groups(qb['num_ans']).cuda()
    inp = _c_inputs(qd)
    inp['R'] = full['R']                      # the regression targets are per candidate ([B,4]); only the IMAGE tensors are per question
    out = CModel(m).forward(inp, group=group)
    _same(out, scores, reg)


def test_c_forward_against_the_reference_goldens():
    """Same bars as the Python host's golden test: the C-scheduled forward against outputs of the UNMODIFIED reference."""
    rec = load_golden('full_eval_b8_mild')
    cfg_path, cfg, sd, batch = golden_inputs(rec)
    params = default_params(cfg_path, device='cuda', max_seq_len=rec['T'], max_vis_features=rec['R'], L1=True)
    m = VisualDialogEncoder(params)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()})
    m.to('cuda').eval()
    gb = {k: v.cuda() for k, v in batch.items()}
    out = CModel(m).forward(_c_inputs(gb))
    ref = rec['logits'].float()
    assert float((out['logits'].cpu() - ref).abs().max() / ref.abs().max()) < 1e-2          # north star: ~1e-2 relative in bf16
    rp = rec['reg_pred'].float()
    assert float((out['reg_pred'].cpu() - rp).abs().max() / rp.abs().max()) < 1e-3              # REG_TOL of tests/test_model_gpu.py


def test_c_forward_follows_weight_updates_and_reports_errors():
    m, params, cfg = _model('tiny.json', 32, 12)
    gb = {k: v.cuda() for k, v in make_batch(4, 32, 12, cfg.v_feature_size, seed=2, vocab_size=cfg.vocab_size).items()}
    cm = CModel(m)
    a = cm.forward(_c_inputs(gb))['logits'].clone()
    with torch.no_grad():
        for p in m.parameters():
            p.mul_(1.01)
    with torch.no_grad():
        _, _, _, _, scores, _ = glue_forward(m, gb, params, evaluation=True)
    b = cm.forward(_c_inputs(gb))['logits']
    assert torch.equal(b, scores) and not torch.equal(a, b)          # the bf16 copy is re-cast before the C forward reads it
    bad = _c_inputs(gb)
    bad['image_feat'] = bad['image_feat'][:2]                         # 2 visual rows for 4 text rows, no group
    with pytest.raises(L.CrctError, match='visual rows'):
        cm.forward(bad)
    import ctypes as C
    ws = torch.empty(1024, dtype=torch.uint8, device='cuda')
    t = _c_inputs(gb)
    ba = L.BatchArgs()
    ba.tokens, ba.segments, ba.loc, ba.attention_mask = (L.ptr(t['tokens']), L.ptr(t['segments']), L.ptr(t['loc']), L.ptr(t['attention_mask']))
    ba.image_feat, ba.image_loc, ba.image_target, ba.image_mask, ba.R4 = (L.ptr(t['image_feat']), L.ptr(t['image_loc']), L.ptr(t['image_target']),
                                                                          L.ptr(t['image_mask']), L.ptr(t['R']))
    ba.B, ba.Bq, ba.T, ba.R = 4, 4, 32, 12
    ba.image_mask_kind = 1
    o = L.OutArgs()
    buf = torch.empty(64, device='cuda')
    o.logits = o.reg_pred = o.reg_loss = o.reg_l1 = o.reg_dist = o.scalars = L.ptr(buf)
    rc = L.lib().crct_forward(cm._h, C.byref(ba), C.byref(o), L.ptr(ws), 1024, L.stream_ptr())
    assert rc == -1 and b'workspace' in L.lib().crct_last_error()
