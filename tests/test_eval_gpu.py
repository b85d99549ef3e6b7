"""f3 on the B200: candidate expansion, answer selection and the accuracy table against the oracle restatement of
CRCT/evaluation.py:254-313,494-525 (oracle/eval_oracle.py), and the de-duplicated question batch against the reference's
replicated layout (CRCT/fig_dataloader.py:690-703).

Bars: indices, flags and counts bit-exact; the selected regression values bit-exact (they are copies);
softmax probability <= 2e-7 absolute (CUDA expf vs the host libm); logits of the two layouts bit-identical."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import _lib as L                                             # noqa: E402
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward             # noqa: E402
from cqa_crct_b200.evaluate import evaluate_batch, expand_question_batch        # noqa: E402
from cqa_crct_b200.spec import ModelConfig, synth_state_dict                    # noqa: E402
from cqa_crct_b200.synthetic import default_params, make_question_batch         # noqa: E402
from oracle import crct_oracle as O                                             # noqa: E402
from oracle.eval_oracle import select_and_score                                 # noqa: E402
from tests.helpers import CONFIG_DIR                                            # noqa: E402


@pytest.mark.parametrize('n,shape,dtype', [(37, (12, 128), torch.bfloat16), (5, (44,), torch.float32), (9, (3,), torch.float32),
                                           (0, (8,), torch.float32), (300, (44, 1024), torch.bfloat16)])
def test_expand_blocks(n, shape, dtype):
    g = torch.Generator().manual_seed(n)
    src = torch.randn((11,) + shape, generator=g).to(dtype).cuda()
    group = torch.randint(0, 11, (n,), generator=g).cuda()
    dst = torch.full((n,) + shape, 7.0, dtype=dtype, device='cuda')
    L.expand_blocks(src, group, dst)
    assert torch.equal(dst, src.index_select(0, group))


@pytest.mark.parametrize('Q,max_ans,force', [(1, 1, False), (23, 40, False), (300, 120, False), (64, 7, True)])
def test_select_and_score_kernels_match_the_oracle(Q, max_ans, force):
    qb = make_question_batch(Q, 16, 4, 8, seed=Q, vocab_size=2048, min_ans=1, max_ans=max_ans)
    if force:
        qb['gt_id'] = qb['gt_id'].clamp(min=0)
    N = int(qb['num_ans'].sum())
    g = torch.Generator().manual_seed(Q)
    scores = torch.randn(N, 2, generator=g) * 3
    if N > 5:
        scores[N // 2] = scores[N // 2 - 1]                       # exact tie (same question or not): first maximum wins
    reg_pred, reg_dist, reg_l1 = torch.randn(N, generator=g), torch.rand(N, generator=g) * 0.1, torch.rand(N, generator=g) * 0.02
    ref = select_and_score(scores, reg_pred, reg_dist, reg_l1, qb['num_ans'], qb['gt_id'], qb['needs_reg'], qb['tolerance_margin'], force)
    dev = 'cuda'
    off = torch.zeros(Q + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(qb['num_ans'], 0)
    answers = torch.empty(Q, dtype=torch.int64, device=dev)
    prob = torch.empty(N, device=dev)
    sel = [torch.empty(Q, device=dev) for _ in range(3)]
    L.select_answers(scores.to(dev), reg_pred.to(dev), reg_dist.to(dev), reg_l1.to(dev), off.to(dev), answers, *sel, prob=prob,
                     forced=qb['gt_id'].to(dev) if force else None)
    flags = torch.empty(Q, 5, dtype=torch.uint8, device=dev)
    total = torch.zeros(6, 2, dtype=torch.float64, device=dev)
    for _ in range(2):                                            # the table accumulates
        L.score_answers(answers, qb['gt_id'].to(dev), qb['needs_reg'].to(torch.uint8).to(dev), sel[1], sel[2],
                        qb['tolerance_margin'].to(dev), total, flags=flags)
    assert float((prob.cpu() - ref['prob']).abs().max()) <= 2e-7
    assert torch.equal(answers.cpu(), ref['answers'])
    assert torch.equal(sel[0].cpu(), ref['reg_output']) and torch.equal(sel[1].cpu(), ref['reg_loss']) and torch.equal(sel[2].cpu(), ref['reg_t_loss'])
    assert torch.equal(flags.cpu(), ref['flags'])
    assert torch.equal(total.cpu(), 2 * ref['total_correct'])


def _tiny_model():
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    cfg = ModelConfig(cfg_path)
    params = default_params(cfg_path, device='cuda', max_seq_len=32, max_vis_features=12)
    m = VisualDialogEncoder(params)
    sd = synth_state_dict(cfg, 228, 3, 'trained')
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()})
    m.to('cuda').eval()
    return m, params, cfg, sd


@pytest.mark.parametrize('chunk', [512, 16, 7])
def test_question_batch_equals_the_replicated_layout_and_the_oracle(chunk):
    m, params, cfg, sd = _tiny_model()
    qb = make_question_batch(9, 32, 12, cfg.v_feature_size, seed=5, vocab_size=cfg.vocab_size, max_ans=9)
    out = evaluate_batch(m, qb, params, eval_batch_size=chunk)
    # (1) replicated layout through the reference-facing glue: same logits bit for bit
    full = expand_question_batch(qb)
    gb = {k: v.cuda() for k, v in full.items()}
    with torch.no_grad():
        _, _, _, _, scores, reg = glue_forward(m, gb, params, evaluation=True)
    assert torch.equal(scores, out['logits'])
    # (2) selection on those logits = the oracle's
    ref = select_and_score(scores.cpu(), reg[0].cpu(), reg[4].cpu(), reg[2].cpu(), qb['num_ans'], qb['gt_id'], qb['needs_reg'],
                           qb['tolerance_margin'])
    assert torch.equal(out['answers'].cpu(), ref['answers']) and torch.equal(out['flags'].cpu(), ref['flags'])
    assert torch.equal(out['total_correct'].cpu(), ref['total_correct'])
    assert torch.equal(out['reg_output'].cpu(), ref['reg_output'])
    # (3) whole path against the fp64 oracle model: same answer wherever the oracle's margin is not a rounding matter
    o, _ = O.forward(sd, O.Config(cfg.__dict__), full, train=False, l1=True, keep_cache=False, dtype=torch.float64)
    oref = select_and_score(o['logits'].float(), o['reg_pred'].float(), o['reg_dist'].float(), o['reg_l1'].float(), qb['num_ans'],
                            qb['gt_id'], qb['needs_reg'], qb['tolerance_margin'])
    p = oref['prob']
    off = 0
    for q, n in enumerate(qb['num_ans'].tolist()):
        top = torch.sort(p[off:off + n], descending=True).values
        if n == 1 or float(top[0] - top[1]) > 0.05:
            assert int(out['answers'][q]) == int(oref['answers'][q]), q
        off += n


def test_training_rejects_shared_visual_rows():
    m, params, cfg, sd = _tiny_model()
    qb = make_question_batch(3, 32, 12, cfg.v_feature_size, seed=1, vocab_size=cfg.vocab_size, max_ans=4)
    gb = {k: v.cuda() for k, v in qb.items()}
    with pytest.raises(ValueError):
        glue_forward(m, gb, params)                              # 3 visual rows for sum(num_ans) text rows, no image_group


def test_eval_pipeline_equals_batch_by_batch_evaluation():
    """evaluate.EvalPipeline (H2D of the next batch under the current batch's kernels, results read one batch late) returns what
    `evaluate_batch` returns batch by batch, and accumulates the same accuracy table."""
    from cqa_crct_b200.evaluate import EvalPipeline
    m, params, cfg, sd = _tiny_model()
    batches = [make_question_batch(5 + i, 32, 12, cfg.v_feature_size, seed=40 + i, vocab_size=cfg.vocab_size, max_ans=7) for i in range(5)]
    pinned = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in b.items()} for b in batches]
    total = None
    want = []
    for b in batches:
        out = evaluate_batch(m, b, params, eval_batch_size=16, total_correct=total)
        total = out['total_correct']
        want.append((out['answers'].cpu(), out['reg_output'].cpu()))
    pipe = EvalPipeline(m, params, eval_batch_size=16, depth=2)
    got, pending = [], None
    for b in pinned:
        h = pipe.submit(b)
        if pending is not None:
            got.append(pending.result())
        pending = h
    got.append(pending.result())
    assert len(got) == len(want)
    for (a, r), (a2, r2) in zip(got, want):
        assert torch.equal(a, a2) and torch.equal(r, r2)
    assert torch.equal(pipe.total_correct.cpu(), total.cpu())
