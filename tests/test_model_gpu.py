"""End-to-end parity on the B200: the CUDA path through the public `VisualDialogEncoder` / `glue_forward` API
against (1) the oracle (fp64) on the same seeded inputs and weights and (2) the golden vectors the reference produced.

The bars come from a YARDSTICK, not from this implementation: `tests/golden/yardstick_bf16.json` (oracle/make_yardstick.py)
holds, per golden case, how far the UNMODIFIED reference moves from its own fp32 run when it is run under
`torch.autocast(bfloat16)` — the reference's own mixed-precision recipe (CRCT/train.py:172 wraps the step in autocast).
A bf16 tensor-core path cannot be asked to be closer to the fp32 reference than the reference's own bf16 run is; it is asked
to be NO WORSE than 1.25x that, and additionally to meet the north star's absolute "about 1e-2" where the model is
well-conditioned (tiny / reference-scale "mild" weights).  All errors relative to the scale (max |.|) of the compared tensor:
  class logits       <= min(1e-2 [tiny, mild], 1.25 x yardstick)     measured 1.4e-3 .. 6.0e-3 (tiny / mild), 1.6e-2 / 2.8e-2 ("trained"
                                                                      2x-wide weights; yardstick 2.6e-2 / 3.7e-2)
  regression output  <= 1e-3                                          measured 1e-6 .. 4e-4 (yardstick 1e-4 .. 4e-3)
  per-row losses     <= 1e-3 absolute, total loss <= 1e-2 absolute    measured <= 5.7e-3
  argmax             identical on every row whose reference margin exceeds twice the logit bar
  gradients          global relative L2 over all tensors <= 1.25 x yardstick     measured 0.004 / 0.010 / 0.017 (tiny; yardstick 0.0056 /
                     0.0101 / 0.0171), 0.102 (mild; 0.0997), 0.114 (trained; 0.144);  per tensor ||got-ref|| <= 0.25 ||ref|| + 5e-3
                     max_t||ref_t||, cosine >= 0.975 for every tensor carrying more than 1 % of the largest gradient norm (measured
                     worst 0.19 / 0.982).
Why the gradients of the full model sit at 0.10 and not 1e-2 (DESIGN.md "Numerical floor"): with the fp32 residual stream the
forward agrees with the reference to 1.4e-3, yet ANY rounding of the saved activations moves the gradients of this 24-block
post-LN network at random weights by 3e-2 .. 1.4e-1 — the reference's own autocast run by 0.0997, the oracle's emulation of
this path's rounding points by 0.03 .. 0.09 depending on the seed (tools/parity_sensitivity.py).  Kernel-level correctness is
pinned separately and tightly in test_kernels_gpu.py / test_varlen_gpu.py, the schedule and the backward derivation to 4e-6 in
test_check_f32_gpu.py."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward   # noqa: E402
from cqa_crct_b200.synthetic import default_params                     # noqa: E402
from oracle import crct_oracle as O                                    # noqa: E402
from tests.helpers import load_golden, golden_inputs, sample_idx, rel_err   # noqa: E402


def build(rec):
    cfg_path, cfg, sd, batch = golden_inputs(rec)
    params = default_params(cfg_path, device='cuda', max_seq_len=rec['T'], max_vis_features=rec['R'], L1=rec['l1'])
    m = VisualDialogEncoder(params)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to('cuda').eval()          # dropout off; branch still chosen by kwargs
    gb = {k: v.to('cuda') for k, v in batch.items()}
    return m, params, cfg, sd, batch, gb


def scale_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


REG_TOL = 1e-3
YARD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'yardstick_bf16.json')))


def logit_bar(name):
    bar = 1.25 * YARD[name]['logits_err']
    return min(bar, 1e-2) if (name.startswith('tiny') or name.endswith('mild')) else bar


def grad_bar(name):
    return 1.25 * YARD[name]['grad_global_rel']


def check_outputs(rec, scores, reg):
    LOGIT_TOL = logit_bar(rec['name'])
    assert scale_err(scores, rec['logits']) < LOGIT_TOL, (scale_err(scores, rec['logits']), LOGIT_TOL)
    assert scale_err(reg[0], rec['reg_pred']) < REG_TOL
    assert float((reg[1].cpu() - rec['reg_loss']).abs().max()) < 1e-3
    assert float((reg[2].cpu() - rec['reg_l1']).abs().max()) < 1e-3
    assert float((reg[4].cpu() - rec['reg_dist']).abs().max()) < 1e-2 * max(1.0, float(rec['reg_dist'].abs().max()))
    margin = (rec['logits'][:, 0] - rec['logits'][:, 1]).abs()
    sure = margin > 2 * LOGIT_TOL * rec['logits'].abs().max()
    assert torch.equal(scores.cpu().argmax(1)[sure], rec['logits'].argmax(1)[sure])


@pytest.mark.parametrize('name', ['tiny_eval', 'full_eval_b8', 'full_eval_b8_mild'])
def test_eval_forward_matches_reference_golden(name):
    rec = load_golden(name)
    m, params, cfg, sd, batch, gb = build(rec)
    with torch.no_grad():
        loss, _, nsp, _, scores, reg = glue_forward(m, gb, params, evaluation=True)
    assert loss is None and nsp is None
    check_outputs(rec, scores, reg)
    out, _ = O.forward(sd, O.Config(cfg.__dict__), batch, train=False, l1=rec['l1'], keep_cache=False)
    assert scale_err(scores, out['logits']) < logit_bar(name)


@pytest.mark.parametrize('name', ['tiny_train_l1', 'tiny_train_smooth', 'tiny_ragged', 'full_train_b4', 'full_train_b4_mild'])
def test_train_forward_backward_matches_reference(name):
    rec = load_golden(name)
    m, params, cfg, sd, batch, gb = build(rec)
    m.zero_grad()
    loss, _, nsp, _, scores, reg, _ = glue_forward(m, gb, params)
    loss.backward()
    torch.cuda.synchronize()
    check_outputs(rec, scores, reg)
    assert abs(float(loss) - rec['loss']) < 1e-2
    assert abs(float(nsp) - rec['nsp_loss']) < 1e-2
    assert (int(reg[3][0]), int(reg[3][1])) == rec['reg_right']
    # full per-tensor gradients against the oracle (fp64), golden summaries against the reference itself
    out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=rec['l1'], dtype=torch.float64)
    g = O.backward(cache)
    named = dict(m.bert_pretrained.named_parameters())
    gnorm = max(float(v.norm()) for v in g.values())
    bad, num, den = [], 0.0, 0.0
    for k, ref in g.items():
        got = named[k].grad
        assert got is not None, k
        got, ref = got.double().cpu(), ref.double()
        err, rn = float((got - ref).norm()), float(ref.norm())
        num, den = num + err * err, den + rn * rn
        if err > 0.25 * rn + 5e-3 * gnorm:
            bad.append((k, err / max(rn, 1e-30), rn / gnorm))
        if rn > 1e-2 * gnorm:
            cos = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.flatten(), dim=0))
            if cos < 0.975:
                bad.append((k, 'cos', cos))
    assert not bad, bad[:10]
    assert (num / den) ** 0.5 < grad_bar(name), ((num / den) ** 0.5, grad_bar(name))
    for k, s in rec['grads'].items():                   # the reference's own gradient norms (golden)
        if s['norm'] > 1e-2 * gnorm:
            got = named[k].grad.double().cpu().flatten()
            assert abs(float(got.norm()) - s['norm']) <= 0.25 * s['norm'], k
    for k, p in named.items():                          # the reference's never-used tensors stay gradient-free
        if k not in g:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k


def test_module_surface_matches_reference_checkpoint_layout():
    rec = load_golden('tiny_eval')
    m, params, cfg, sd, batch, gb = build(rec)
    keys = list(m.state_dict().keys())
    assert keys == ['bert_pretrained.' + k for k in sd.keys() if k != 'cls.predictions.decoder.weight'][:len(keys)] or set(keys) == {'bert_pretrained.' + k for k in sd}
    assert len(keys) == len(sd)
    assert m.state_dict()['bert_pretrained.cls.predictions.decoder.weight'].data_ptr() == \
        m.state_dict()['bert_pretrained.bert.embeddings.word_embeddings.weight'].data_ptr()
    # weights survive a round trip through the flat arena bit-exactly
    for k, v in sd.items():
        assert torch.equal(m.state_dict()['bert_pretrained.' + k].cpu(), v), k
    # train()/eval() only switch dropout; with dropout on, two forwards differ and the loss stays finite
    m.train()
    l1 = glue_forward(m, gb, params, evaluation=True)[4]
    l2 = glue_forward(m, gb, params, evaluation=True)[4]
    assert torch.isfinite(l1).all() and not torch.equal(l1, l2)
    m.eval()
    l3 = glue_forward(m, gb, params, evaluation=True)[4]
    l4 = glue_forward(m, gb, params, evaluation=True)[4]
    assert torch.equal(l3, l4)


def test_gradient_accumulation_and_zero_grad():
    rec = load_golden('tiny_train_l1')
    m, params, cfg, sd, batch, gb = build(rec)
    m.zero_grad()
    glue_forward(m, gb, params)[0].backward()
    g1 = m.arena.g32.clone()
    glue_forward(m, gb, params)[0].backward()          # no zero_grad: gradients accumulate like torch .grad
    assert rel_err(m.arena.g32, 2 * g1) < 1e-3
    m.zero_grad()
    assert float(m.arena.g32.abs().sum()) == 0.0


def test_dropout_training_step_is_finite_and_unbiased():
    rec = load_golden('tiny_train_l1')
    m, params, cfg, sd, batch, gb = build(rec)
    m.eval(); m.zero_grad()
    glue_forward(m, gb, params)[0].backward()
    g_ref = m.arena.g32.clone()
    m.train()
    acc = torch.zeros_like(g_ref)
    n = 24
    for _ in range(n):
        m.zero_grad()
        loss = glue_forward(m, gb, params)[0]
        loss.backward()
        assert torch.isfinite(loss)
        acc += m.arena.g32
    assert torch.isfinite(acc).all()
    # dropout changes the function, so only the direction of the averaged gradient is compared
    cos = torch.nn.functional.cosine_similarity((acc / n).flatten(), g_ref.flatten(), dim=0)
    assert float(cos) > 0.5


def test_stress_shape_forward_backward_runs():
    """BASELINE configs[4]: 2x regions, 2x tokens (T=248, R=88) on the full model, B=2."""
    cfg_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'cqa_crct_b200', 'config', 'vilbert.json')
    from cqa_crct_b200.synthetic import make_batch
    params = default_params(cfg_path, device='cuda', max_seq_len=248, max_vis_features=88)
    torch.manual_seed(0)
    m = VisualDialogEncoder(params).to('cuda').eval()
    gb = {k: v.to('cuda') for k, v in make_batch(2, 248, 88, 1024, seed=3).items()}
    m.zero_grad()
    loss = glue_forward(m, gb, params)[0]
    loss.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and torch.isfinite(m.arena.g32).all()


@pytest.mark.parametrize('overlap_optimizer', [False, True])
def test_graphed_train_step_matches_eager_and_redraws_dropout(overlap_optimizer):
    """cqa_crct_b200.graph.GraphedTrainStep: the captured step updates the weights exactly like the eager calls
    (dropout off), and with dropout on every replay draws new masks (device salt) — losses differ between replays.
    `overlap_optimizer`: AdamW launched range by range under the backward instead of after it — same result."""
    from cqa_crct_b200.graph import GraphedTrainStep
    from cqa_crct_b200.optim import FusedAdamW
    rec = load_golden('tiny_train_l1')
    m1, params, cfg, sd, batch, gb = build(rec)
    m2, *_ = build(rec)
    params = dict(params, overlap_optimizer=overlap_optimizer, optimizer_chunk=1 << 16)     # several optimizer launches on the tiny model
    o1, o2 = FusedAdamW(m1, lr=2e-5, image_lr=2e-5), FusedAdamW(m2, lr=2e-5, image_lr=2e-5)      # the reference's lr (options.py:21)
    g = GraphedTrainStep(m2, o2, params, gb, warmup_steps=1)          # 1 eager warm-up step applied; capture itself runs nothing
    o1.zero_grad()
    glue_forward(m1, gb, params)[0].backward()
    o1.step()
    e0 = rel_err(m2.arena.w32, m1.arena.w32)
    assert e0 < 1e-5, e0               # fp32 split-K atomics: not bit-identical
    l_eager = None
    for _ in range(3):
        o1.zero_grad()
        l_eager = glue_forward(m1, gb, params)[0]
        l_eager.backward()
        o1.step()
        l_graph = g.step(gb)
    torch.cuda.synchronize()
    assert abs(float(l_eager) - float(l_graph)) < 1e-3, (float(l_eager), float(l_graph))
    e1 = rel_err(m2.arena.w32, m1.arena.w32)
    assert e1 < 1e-4, e1
    # input pipeline: a prefetched pinned host batch + asynchronous loss read give the same step
    pinned = {k: v.cpu().pin_memory() for k, v in gb.items()}
    g.prefetch(pinned)
    h = g.step_async()
    o1.zero_grad()
    l_eager = glue_forward(m1, gb, params)[0]
    l_eager.backward()
    o1.step()
    assert abs(float(l_eager) - h.item()) < 1e-3
    assert rel_err(m2.arena.w32, m1.arena.w32) < 1e-4
    m2.train()                                                       # dropout on: a new graph, masks must change per replay
    g2 = GraphedTrainStep(m2, o2, params, gb, warmup_steps=1)
    losses = [float(g2.step(gb)) for _ in range(4)]
    assert len(set(round(x, 6) for x in losses)) > 1 and all(x == x for x in losses)


def _full_model(style='mild', seed=1):
    from cqa_crct_b200.spec import ModelConfig, synth_state_dict
    cfg_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'cqa_crct_b200', 'config', 'vilbert.json')
    cfg = ModelConfig(cfg_path)
    params = default_params(cfg_path, device='cuda', max_seq_len=124, max_vis_features=44, L1=True)
    m = VisualDialogEncoder(params)
    sd = synth_state_dict(cfg, 228, seed, style)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to('cuda').eval()
    return m, params, cfg, sd


def test_full_model_per_question_argmax_matches_the_oracle():
    """The evaluation outcome itself (CRCT/evaluation.py:254-258,287-296) on the FULL model: 16 questions x 32 candidate
    sequences through `evaluate_batch` (question-level visual rows, packed tokens) against the fp32 oracle on the replicated
    layout.  Stated in probabilities p = softmax(nsp)[:,0] (what the argmax runs over):
      * every candidate's p within 1e-3 of the oracle's (measured 5.7e-4; logits <= 1e-2 of their scale, measured 6.6e-3);
      * REGRET: the oracle probability of the candidate the CUDA path selects is within 6e-4 of the oracle's best, for every
        question (measured 1.2e-4 on the one question — margin 1.25e-4 — where the selection differs);
      * IDENTITY: the selected candidate is the oracle's wherever the oracle's best-vs-second margin exceeds 1.2e-3 (twice the
        largest within-question differential error measured, 5.5e-4), which covers at least half of the questions
        (independent candidate sequences, `distinct=True`: margins 6e-5 .. 9e-3 at random-init weights)."""
    from cqa_crct_b200.evaluate import evaluate_batch, expand_question_batch
    from cqa_crct_b200.synthetic import make_question_batch
    m, params, cfg, sd = _full_model()
    qb = make_question_batch(16, 124, 44, cfg.v_feature_size, seed=77, total=512, distinct=True)
    out = evaluate_batch(m, qb, params, eval_batch_size=512)
    full = expand_question_batch(qb)
    with torch.no_grad():
        o, _ = O.forward(sd, O.Config(cfg.__dict__), full, train=False, l1=True, keep_cache=False)
    assert scale_err(out['logits'], o['logits']) < 1e-2
    p = torch.softmax(o['logits'], 1)[:, 0]
    pc = out['prob'].cpu()
    assert float((pc - p).abs().max()) < 1e-3
    off, sure = 0, 0
    for q, n in enumerate(qb['num_ans'].tolist()):
        top = torch.sort(p[off:off + n], descending=True)
        chosen = int(out['answers'][q])
        assert float(top.values[0] - p[off + chosen]) <= 6e-4, (q, float(top.values[0] - p[off + chosen]))
        if n == 1 or float(top.values[0] - top.values[1]) > 1.2e-3:
            sure += 1
            assert chosen == int(top.indices[0]), (q, float(top.values[0] - top.values[1]))
        off += n
    assert sure >= 8, sure                      # the margin rule must not make the test vacuous


def test_full_size_b80_train_gradients_against_the_oracle():
    """BASELINE configs[1] at its full size (B = 80, T = 124, R = 44, dropout off): loss, logits and the gradients of a spread of
    tensors (embeddings, first / middle / last text, visual and co-attention blocks, poolers, classifier, regressor) against
    the fp32 oracle run on the host cores."""
    from cqa_crct_b200.synthetic import make_batch
    m, params, cfg, sd = _full_model()
    batch = make_batch(80, 124, 44, cfg.v_feature_size, seed=4242)
    gb = {k: v.to('cuda') for k, v in batch.items()}
    m.zero_grad()
    loss, _, nsp, _, scores, reg, _ = glue_forward(m, gb, params)
    loss.backward()
    torch.cuda.synchronize()
    out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=True)
    g = O.backward(cache)
    assert abs(float(loss) - float(out['loss'])) < 1e-2
    assert scale_err(scores, out['logits']) < 1e-2
    assert scale_err(reg[0], out['reg_pred']) < REG_TOL
    named = dict(m.bert_pretrained.named_parameters())
    picks = ['bert.embeddings.word_embeddings.weight', 'bert.embeddings.LayerNorm.weight', 'bert.v_embeddings.new_image_embeddings.weight',
             'bert.encoder.layer.0.attention.self.query.weight', 'bert.encoder.layer.5.intermediate.dense.weight',
             'bert.encoder.layer.11.output.dense.weight', 'bert.encoder.v_layer.0.attention.self.value.weight',
             'bert.encoder.v_layer.5.output.dense.weight', 'bert.encoder.c_layer.0.biattention.key2.weight',
             'bert.encoder.c_layer.3.biOutput.dense1.weight', 'bert.encoder.c_layer.5.t_output.dense.weight', 'bert.t_pooler.dense.weight',
             'bert.v_pooler.dense.weight', 'cls.bi_seq_relationship.weight', 'regressor.fusion.0.weight', 'regressor.txt_pipe.0.weight']
    gnorm = max(float(v.norm()) for v in g.values())
    num = den = 0.0
    for k, ref in g.items():
        got = named[k].grad.double().cpu()
        num, den = num + float((got - ref.double()).norm() ** 2), den + float(ref.double().norm() ** 2)
    glob = (num / den) ** 0.5
    assert glob < 1.25 * YARD['full_train_b4_mild']['grad_global_rel'], glob
    for k in picks:
        got, ref = named[k].grad.double().cpu(), g[k].double()
        err, rn = float((got - ref).norm()), float(ref.norm())
        assert err <= 0.25 * rn + 5e-3 * gnorm, (k, err / rn)


def test_weights_changed_after_the_first_forward_are_picked_up():
    """after `.to('cuda')` the nn.Parameters no longer share the arena's version counter; load_state_dict /
    optimizers / manual copy_ into a parameter must still refresh the bf16 operand copy (several checkpoints evaluated with
    one model object)."""
    from cqa_crct_b200.spec import synth_state_dict
    rec = load_golden('tiny_eval')
    m, params, cfg, sd, batch, gb = build(rec)
    with torch.no_grad():
        s1 = glue_forward(m, gb, params, evaluation=True)[4].clone()
    sd2 = synth_state_dict(cfg, 228, 7, 'mild')
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd2.items()}, strict=True)
    with torch.no_grad():
        s2 = glue_forward(m, gb, params, evaluation=True)[4].clone()
    fresh = VisualDialogEncoder(params)
    fresh.load_state_dict({'bert_pretrained.' + k: v for k, v in sd2.items()}, strict=True)
    fresh.to('cuda').eval()
    with torch.no_grad():
        s3 = glue_forward(fresh, gb, params, evaluation=True)[4]
    assert not torch.equal(s1, s2) and torch.equal(s2, s3)
    with torch.no_grad():                                   # an in-place edit of one parameter
        m.bert_pretrained.cls.bi_seq_relationship.bias.add_(1.0)
        s4 = glue_forward(m, gb, params, evaluation=True)[4]
    assert float((s4 - s2 - 1.0).abs().max()) < 1e-5


def test_two_forwards_before_the_first_backward_keep_their_own_dropout_masks():
    """l1 = model(a); l2 = model(b); (l1 + l2).backward() — each pass recomputes the masks of ITS forward (per-pass
    salt snapshot), so the summed gradient equals the sum of the two separately back-propagated passes with the same salts."""
    rec = load_golden('tiny_train_l1')
    m, params, cfg, sd, batch, gb = build(rec)
    m.train()
    gb2 = {k: v.clone() for k, v in gb.items()}
    gb2['tokens'] = gb['tokens'].roll(1, 0)

    def salt_reset():
        m._salt = None                                       # re-seed the live counter: same sequence of per-pass salts
    salt_reset()
    m.zero_grad()
    l1 = glue_forward(m, gb, params)[0]
    l2 = glue_forward(m, gb2, params)[0]
    (l1 + l2).backward()
    both = m.arena.g32.clone()
    salt_reset()
    m.zero_grad()
    l1b = glue_forward(m, gb, params)[0]
    l1b.backward()
    l2b = glue_forward(m, gb2, params)[0]
    l2b.backward()
    torch.cuda.synchronize()
    assert float(l1) == float(l1b) and float(l2) == float(l2b)
    assert rel_err(both, m.arena.g32) < 1e-4
