"""Seeded synthetic PlotQA-shaped batches (SURVEY.md §8d "Synthetic inputs").

The detection stage (Mask-RCNN / Detectron2) is out of scope, so every test and
benchmark feeds the question-answering stage with tensors that have the shape,
dtype and value ranges of what `PlotQA_Dataset.__getitem__` produces
(reference: CRCT/fig_dataloader.py:308-350, 524-690; CRCT/utils.py:105-225).

Everything is generated with a CPU `torch.Generator`, so the same seed gives
bit-identical batches in the build container and on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

CLS_ID, SEP_ID = 101, 102
IMG_CLASS = 228          # `<IMG>` whole-figure box, CRCT/fig_dataloader.py:338


def make_batch(B: int, T: int = 124, R: int = 44, feat_dim: int = 1024, seed: int = 1234,
               vocab_size: int = 30522, categories: int = 228, min_answer: int = 1) -> Dict[str, torch.Tensor]:
    """One batch in the layout `encoder_decorator.forward` consumes
    (reference: CRCT/backbone/encoder_decorator.py:81-116).

    Returns CPU tensors:
      tokens[B,T] i64, loc[B,T,4] f32, segments[B,T] i64 in {-1,0..11}, mask[B,T] i64 (ignored),
      sep_indices[B,1] i64 + hist_len[B] i64 (sequence length = sep_indices[hist_len]+1),
      image_feat[B,R,F] f32, image_loc[B,R,4] f32, image_mask[B,R] i64, image_target[B,R] i64,
      image_label[B,R] i64 (ignored), next_sentence_labels[B,1] i64, R[B,4] f32, needs_reg[B] bool.
    """
    g = torch.Generator().manual_seed(seed)

    def randint(lo, hi, size=()):
        return torch.randint(lo, hi + 1, size, generator=g)

    tokens = torch.zeros(B, T, dtype=torch.int64)
    segments = torch.zeros(B, T, dtype=torch.int64)
    loc = torch.zeros(B, T, 4, dtype=torch.float32)
    seq_len = torch.zeros(B, dtype=torch.int64)
    lo_len = min(48, max(8, T // 2))
    for b in range(B):
        lt = int(randint(lo_len, T))
        lq_hi = max(1, min(40, lt - 6))
        lq = int(randint(min(8, lq_hi), lq_hi))
        la = int(randint(min_answer, 4))                 # answer tokens incl. the closing [SEP]
        n_chart = max(0, lt - 1 - lq - la - 2)          # [CLS] chart.. [SEP] question [SEP] answer.. (last = [SEP])
        lt = 1 + n_chart + 1 + lq + la                   # exact valid length
        ids = randint(1000, vocab_size - 1, (lt,))
        ids[0] = CLS_ID
        ty = torch.zeros(lt, dtype=torch.int64)
        bx = torch.zeros(lt, 4)
        # chart text: types 2..11, boxes U[0,1]^4 with ~25% all-zero boxes
        if n_chart > 0:
            ty[1:1 + n_chart] = randint(2, 11, (n_chart,))
            cb = torch.rand(n_chart, 4, generator=g)
            cb[torch.rand(n_chart, generator=g) < 0.25] = 0
            bx[1:1 + n_chart] = cb
        p = 1 + n_chart
        ids[p] = SEP_ID
        ty[p] = ty[p - 1] if n_chart > 0 else 0
        # question (type -1) and answer (type 1); zero boxes
        ty[p + 1:p + 1 + lq] = -1
        ids[p + lq] = SEP_ID
        ty[p + 1 + lq:] = 1
        ids[lt - 1] = SEP_ID
        tokens[b, :lt] = ids
        segments[b, :lt] = ty
        loc[b, :lt] = bx
        seq_len[b] = lt

    image_feat = torch.zeros(B, R, feat_dim)
    image_loc = torch.zeros(B, R, 4)
    image_mask = torch.zeros(B, R, dtype=torch.int64)
    image_target = torch.zeros(B, R, dtype=torch.int64)
    for b in range(B):
        lr = int(randint(min(4, R), R))
        f = torch.relu(torch.randn(lr, feat_dim, generator=g)) * 1.5     # post-ReLU box_head output
        image_feat[b, :lr] = f
        bx = torch.rand(lr, 4, generator=g) * 1.2 - 0.1
        bx[0] = 0
        image_loc[b, :lr] = bx
        cls = randint(8, categories - 1, (lr,))
        cls[0] = IMG_CLASS if categories >= IMG_CLASS else categories
        image_target[b, :lr] = cls
        image_mask[b, :lr] = 1

    nsl = randint(0, 1, (B, 1))
    needs_reg = torch.rand(B, generator=g) < 0.3
    if B > 1:
        needs_reg[0] = True
    scale = torch.exp(torch.rand(B, generator=g) * math.log(1e4))
    ratio = torch.rand(B, generator=g) * 2.4 - 1.2
    Rt = torch.stack([ratio * scale, needs_reg.float(), torch.full((B,), 0.01), scale], dim=1)

    return {
        'tokens': tokens, 'loc': loc, 'segments': segments,
        'mask': torch.zeros(B, T, dtype=torch.int64),
        'sep_indices': (seq_len - 1).view(B, 1), 'hist_len': torch.zeros(B, dtype=torch.int64),
        'image_feat': image_feat, 'image_loc': image_loc, 'image_mask': image_mask,
        'image_target': image_target, 'image_label': torch.zeros(B, R, dtype=torch.int64),
        'next_sentence_labels': nsl, 'R': Rt.float(), 'needs_reg': needs_reg,
    }


def default_params(model_config: str, device: str = 'cpu', **over) -> dict:
    """The subset of `options.read_command_line` keys the model reads
    (reference: CRCT/options.py:9-124, CRCT/backbone/vilbert.py:1461-1467,1518-1533,1583-1637)."""
    p = {
        'model_config': model_config, 'categories': 228, 'dataset': 'plotqa', 'mask_prob_img': 0,
        'binary_answers': False, 'qa_file': 'synthetic', 'CE_REG': False, 'L1': True, 'rank': 0,
        'rank_from': 1, 'BOT_MODE': False, 'max_seq_len': 124, 'max_vis_features': 44,
        'device': torch.device(device), 'tol_margin': 0.01, 'dvqa_floats': [],
        'nsp_loss_coeff': 1.0, 'reg_loss_coeff': 1.0,
    }
    p.update(over)
    return p


def make_question_batch(Q: int, T: int = 124, R: int = 44, feat_dim: int = 1024, seed: int = 1234, vocab_size: int = 30522,
                        min_ans: int = 2, max_ans: int = 48, total: int = None, distinct: bool = False) -> Dict[str, torch.Tensor]:
    """One EVALUATION batch in the de-duplicated layout of `cqa_crct_b200.evaluate` (f3): Q questions, question q with
    `num_ans[q]` candidate answers (reference: one dataset item per question with all its candidates,
    CRCT/fig_dataloader.py:584-587,648,690-693; up to EVAL_PADDED_SIZE = 120 candidates, fixed vocabulary alone = 35).
    Text tensors hold one row per candidate (N = sum(num_ans)), visual tensors and R one row per question.
    `total` forces N (the last question absorbs the difference).  `distinct`: the candidates of a question are independent random
    sequences instead of sharing everything but the answer span — at random-init weights candidates that differ in 1-3 answer
    tokens are separated by 1e-5 .. 1e-3 in probability (no usable argmax margin); independent sequences are separated like
    trained candidates are."""
    g = torch.Generator().manual_seed(seed ^ 0x5EED)
    num_ans = torch.randint(min_ans, max_ans + 1, (Q,), generator=g)
    if total is not None:
        base = max(min_ans, total // Q)
        num_ans = torch.full((Q,), base, dtype=torch.int64)
        num_ans[-1] = total - base * (Q - 1)
        assert int(num_ans[-1]) >= 1
    N = int(num_ans.sum())
    txt = make_batch(N, T, R=1, feat_dim=8, seed=seed + 1, vocab_size=vocab_size, min_answer=2)     # candidates differ in >= 1 token
    vis = make_batch(Q, T=8, R=R, feat_dim=feat_dim, seed=seed + 2, vocab_size=vocab_size)
    # candidates of one question share everything but the answer tokens: copy the first candidate's row, then re-draw the
    # answer span (type 1) of the others
    off = 0
    for q in range(Q):
        n = int(num_ans[q])
        if distinct:
            off += n
            continue
        for k in ('tokens', 'loc', 'segments', 'sep_indices', 'hist_len'):
            txt[k][off + 1:off + n] = txt[k][off]
        ans = (txt['segments'][off] == 1).nonzero().view(-1)
        if ans.numel() > 1:
            span = ans[:-1]                                                     # keep the closing [SEP]
            txt['tokens'][off + 1:off + n, span] = torch.randint(1000, vocab_size, (n - 1, span.numel()), generator=g)
        off += n
    gt_id = (torch.rand(Q, generator=g) * num_ans).long()
    gt_id[torch.rand(Q, generator=g) < 0.05] = -1                               # ground truth not among the candidates (fig_dataloader.py:590-599)
    out = {k: txt[k] for k in ('tokens', 'loc', 'segments', 'mask', 'sep_indices', 'hist_len', 'next_sentence_labels')}
    out.update({k: vis[k] for k in ('image_feat', 'image_loc', 'image_mask', 'image_target', 'image_label', 'R')})
    out.update({'num_ans': num_ans, 'gt_id': gt_id, 'needs_reg': vis['needs_reg'], 'tolerance_margin': vis['R'][:, 2].clone(),
                'id': torch.arange(Q)})
    return out
