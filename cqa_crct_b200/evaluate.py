"""f3 — evaluation path: de-duplicated question batches, device-side answer selection, accuracy table.

Reference behaviour being replaced (CRCT/evaluation.py:231-317, CRCT/fig_dataloader.py:584-587,690-703):
every evaluation item is ONE question with `num_ans` candidate answers; the text tensors differ per candidate, the
visual tensors (`PADDING_VIS`: image_feat, image_loc, image_mask, image_target, image_label, R) are `expand`ed to
`num_ans` identical copies on the host and padded to 120; `cut_batch_padding` concatenates the real rows, the loop runs
the model over chunks of `eval_batch_size` candidate sequences, and a Python loop over questions takes
`argmax(softmax(nsp_scores)[:, 0])` inside each question and picks that candidate's regression outputs (several device
syncs per question), then `reduce_total_acc` updates a 6x2 table.

Here a batch keeps the visual tensors once per question (`QuestionBatch`): 180 KB of fp32 RoI features cross PCIe once
per question instead of once per candidate, the visual embedding (softmax + 1024x1024 projection + LayerNorm) runs once
per question and is fanned out on the device (`VisualDialogEncoder.forward(..., image_group=)`), and the selection /
flags / table are two kernels (`crct_select_answers`, `crct_score_answers`) with no host read-back.
`expand_question_batch` rebuilds the reference's replicated layout (used by the parity tests: both layouts give
bit-identical logits).

Out of scope: the per-template breakdown tensor, the error histogram and the CSV log of CRCT/evaluation.py:319-420
(reporting, fed by the same per-question outputs returned here).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist

from . import _lib as L

TEXT_KEYS = ('tokens', 'segments', 'sep_indices', 'mask', 'next_sentence_labels', 'hist_len', 'loc')     # PADDING_TXT, fig_dataloader.py:27-28
VIS_KEYS = ('image_feat', 'image_loc', 'image_mask', 'image_target', 'image_label', 'R')                 # PADDING_VIS, fig_dataloader.py:30-32
QUESTION_KEYS = ('num_ans', 'gt_id', 'needs_reg', 'tolerance_margin', 'id')


def candidate_groups(num_ans: torch.Tensor) -> torch.Tensor:
    """[N] question index of every candidate sequence (N = sum(num_ans))."""
    return torch.repeat_interleave(torch.arange(num_ans.numel(), device=num_ans.device), num_ans.view(-1))


def expand_question_batch(qb: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """The reference's layout after `cut_batch_padding` (fig_dataloader.py:697-703): visual tensors replicated per candidate."""
    grp = candidate_groups(qb['num_ans'])
    out = dict(qb)
    for k in VIS_KEYS:
        if k in qb:
            out[k] = qb[k].index_select(0, grp.to(qb[k].device))
    return out


C_FORWARD_MAX = 256        # candidate sequences per chunk up to which the C-scheduled forward is used (above: GPU-bound either way)


def _use_c_forward(enc, params, n):
    return (bool(params.get('c_forward', True)) and n <= C_FORWARD_MAX and enc.varlen and not enc.fp32 and not enc.training
            and enc.cfg.max_position_embeddings >= 1)


def _c_model(enc):
    cm = enc.__dict__.get('_c_model')
    if cm is None:
        from .capi import CModel
        cm = CModel(enc)
        enc.__dict__['_c_model'] = cm           # plain attribute (not a submodule / parameter)
    return cm


def _chunk_forward(model, qb_dev, params, c0, c1, grp):
    """One chunk of candidate sequences [c0, c1) through the model (the body of encoder_decorator.forward:73-158 in its
    evaluation branch, with the visual tensors taken once per question)."""
    sl = slice(c0, c1)
    q0, q1 = int(grp[c0]), int(grp[c1 - 1]) + 1          # host copy of the grouping: no device read-back
    tokens, sep_indices, hist_len = qb_dev['tokens'][sl], qb_dev['sep_indices'][sl], qb_dev['hist_len'][sl]
    seq_len = torch.gather(sep_indices, 1, hist_len.view(-1, 1)).squeeze(1) + 1                          # :118-119
    attention_mask = torch.arange(tokens.shape[1], device=tokens.device).unsqueeze(0) < seq_len.unsqueeze(1)
    group = qb_dev['_group'][sl] - q0
    Rq = qb_dev['R'][q0:q1].contiguous()
    Rc = torch.empty(c1 - c0, 4, dtype=torch.float32, device=tokens.device)
    L.expand_blocks(Rq, group, Rc)
    if _use_c_forward(model, params, c1 - c0):
        # small chunks (one question with its candidates, interactive use): the Python schedule is host-bound at ~8 ms per forward
        # whatever the batch; the library-scheduled forward (crct_forward, csrc/model.cu) enqueues the same kernels in the same
        # order from C++ — bit-identical outputs, 2.3 ms at B = 1, 3.0 ms at B = 32, 6.5 ms at B = 128 (tools/capi_latency.py)
        out = _c_model(model).forward(
            {'tokens': tokens, 'segments': qb_dev['segments'][sl], 'loc': qb_dev['loc'][sl], 'attention_mask': attention_mask,
             'image_feat': qb_dev['image_feat'][q0:q1], 'image_loc': qb_dev['image_loc'][q0:q1], 'image_target': qb_dev['image_target'][q0:q1],
             'image_mask': qb_dev['image_mask'][q0:q1], 'R': Rc}, group=group, fill=model.row_fill_hint or (0.0, 0.0))
        return out['logits'], [out['reg_pred'], out['reg_loss'], out['reg_l1'], (out['scalars'][3], out['scalars'][4]), out['reg_dist']]
    _, _, _, scores, reg, _ = model(
        tokens, qb_dev['loc'][sl], qb_dev['image_feat'][q0:q1], qb_dev['image_loc'][q0:q1], sep_indices=sep_indices,
        sep_len=hist_len + 1, token_type_ids=qb_dev['segments'][sl], masked_lm_labels=qb_dev['mask'][sl],
        attention_mask=attention_mask, next_sentence_label=None, output_nsp_scores=True,
        image_attention_mask=qb_dev['image_mask'][q0:q1], image_label=None, image_target=qb_dev['image_target'][q0:q1],
        gt_reg=[Rc, 'L1'], image_group=group)                                                            # :106 kind 'L1' at evaluation
    return scores, reg


def to_device(qb: Dict[str, torch.Tensor], device) -> Dict[str, torch.Tensor]:
    """Host -> device copies of one question batch (pinned sources copy asynchronously); adds the candidate grouping."""
    out = {k: (v.to(device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in qb.items()}
    grp = candidate_groups(qb['num_ans'].cpu())
    out['_group_host'] = grp
    out['_group'] = grp.to(device, non_blocking=True)
    off = torch.zeros(qb['num_ans'].numel() + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(qb['num_ans'].view(-1).cpu(), 0)
    out['_offsets'] = off.to(device, non_blocking=True)
    # expected fraction of valid rows (host-side data, no device read-back): steers the GEMM tile shapes only
    seq_len = torch.gather(qb['sep_indices'].cpu(), 1, qb['hist_len'].cpu().view(-1, 1)).squeeze(1) + 1
    T = qb['tokens'].shape[1]
    fv_q = (qb['image_mask'].cpu() != 0).float().mean(1)                    # per question; candidates inherit their question's rows
    out['_fill'] = (max(float(seq_len.clamp(max=T).sum()) / float(seq_len.numel() * T), 1e-3),
                    max(float((fv_q * qb['num_ans'].cpu().view(-1).float()).sum() / qb['num_ans'].sum()), 1e-3))
    return out


def evaluate_batch(model, qb: Dict[str, torch.Tensor], params: dict, eval_batch_size: int = 512,
                   total_correct: Optional[torch.Tensor] = None, dist_group=None, force_gt: bool = False,
                   reduce: bool = True) -> Dict[str, torch.Tensor]:
    """CRCT/evaluation.py:231-317 for one dataloader batch.  Returns device tensors:
    `answers [Q]` (index within the question), `prob [N]` (softmax(nsp)[:,0] per candidate), `reg_output / reg_loss /
    reg_t_loss [Q]` (the selected candidate's regression[0] / [4] / [2]), `flags [Q,5]` uint8 = (nsp_right, reg_right,
    reg_t_right, correct +-5 %, correct within tolerance) and `total_correct [6,2]` float64 (the table of
    `reduce_total_acc`, summed over ranks when a process group is initialised), accumulated into the tensor passed in.
    `force_gt` is the '_REGS' branch of evaluation.py:288-289 (answer = gt_id); `reduce=False` leaves the cross-rank sum to
    the caller (one all-reduce of the final table instead of one per batch)."""
    enc = getattr(model, 'module', model)
    dev = enc.arena.w32.device
    if dev.type != 'cuda':
        raise L.CrctError('cqa_crct_b200 has no CPU path: move the model to a B200 with .to("cuda")')
    if '_group' not in qb:
        qb = to_device(qb, dev)
    grp_host = qb['_group_host']
    if enc.varlen and params.get('row_fill_hint') is None and '_fill' in qb:
        enc.row_fill_hint = qb['_fill']
    N, Q = qb['tokens'].shape[0], qb['num_ans'].numel()
    if N != grp_host.numel():
        raise ValueError(f'{N} candidate sequences but sum(num_ans) = {grp_host.numel()}')
    logits = torch.empty(N, 2, dtype=torch.float32, device=dev)
    reg_pred = torch.empty(N, dtype=torch.float32, device=dev)
    reg_dist = torch.empty(N, dtype=torch.float32, device=dev)
    reg_l1 = torch.empty(N, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for c0 in range(0, N, eval_batch_size):                                    # evaluation.py:242-263
            c1 = min(c0 + eval_batch_size, N)
            scores, reg = _chunk_forward(enc, qb, params, c0, c1, grp_host)
            logits[c0:c1].copy_(scores)
            reg_pred[c0:c1].copy_(reg[0])
            reg_dist[c0:c1].copy_(reg[4])
            reg_l1[c0:c1].copy_(reg[2])
    answers = torch.empty(Q, dtype=torch.int64, device=dev)
    prob = torch.empty(N, dtype=torch.float32, device=dev)
    sel = [torch.empty(Q, dtype=torch.float32, device=dev) for _ in range(3)]
    forced = qb['gt_id'].view(-1).to(torch.int64).contiguous() if force_gt else None
    L.select_answers(logits, reg_pred, reg_dist, reg_l1, qb['_offsets'], answers, *sel, prob=prob, forced=forced)
    flags = torch.empty(Q, 5, dtype=torch.uint8, device=dev)
    batch_total = torch.zeros(6, 2, dtype=torch.float64, device=dev)
    needs = qb['needs_reg'].view(-1).to(torch.uint8).contiguous()
    L.score_answers(answers, qb['gt_id'].view(-1).to(torch.int64).contiguous(), needs, sel[1], sel[2],
                    qb['tolerance_margin'].view(-1).float().contiguous(), batch_total, flags=flags)
    if reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(dist_group) > 1:
        dist.all_reduce(batch_total, op=dist.ReduceOp.SUM, group=dist_group)      # evaluation.py:519-521
    if total_correct is None:
        total_correct = torch.zeros(6, 2, dtype=torch.float64, device=dev)
    total_correct += batch_total
    return {'answers': answers, 'prob': prob, 'reg_output': sel[0], 'reg_loss': sel[1], 'reg_t_loss': sel[2], 'flags': flags,
            'total_correct': total_correct, 'logits': logits}


class EvalPipeline:
    """Overlapped evaluation loop (the reference's loop is synchronous per batch: copy, forward, `.item()`s — CRCT/evaluation.py:231-317).

        pipe = EvalPipeline(model, params, eval_batch_size)
        pending = None
        for host_batch in loader:                      # pinned host tensors (question batches)
            h = pipe.submit(host_batch)                # H2D on a copy stream, forward + selection enqueued behind it, results -> pinned host
            if pending is not None:
                answers, reg_output = pending.result() # waits only for THAT batch
            pending = h

    The host->device copies of batch i+1 run under the kernels of batch i, and the host reads batch i's answers after batch i+1 has
    been enqueued, so the GPU never idles between batches.  Same arithmetic as `evaluate_batch` (it is what runs); `total_correct`
    accumulates on the device across submits."""

    def __init__(self, model, params: dict, eval_batch_size: int = 512, depth: int = 3):
        self.model, self.params, self.ebs = model, params, eval_batch_size
        enc = getattr(model, 'module', model)
        self.dev = enc.arena.w32.device
        self._copy = torch.cuda.Stream(device=self.dev)
        self.total_correct = torch.zeros(6, 2, dtype=torch.float64, device=self.dev)
        self._slots = [None] * depth            # pinned result buffers, reused round-robin
        self._i = 0

    def submit(self, qb: Dict[str, torch.Tensor], force_gt: bool = False):
        cur = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self._copy):
            dev_qb = to_device(qb, self.dev)
            ready = torch.cuda.Event()
            ready.record(self._copy)
        cur.wait_event(ready)
        for v in dev_qb.values():
            if torch.is_tensor(v) and v.is_cuda:
                v.record_stream(cur)            # allocated on the copy stream, consumed on the compute stream
        out = evaluate_batch(self.model, dev_qb, self.params, self.ebs, total_correct=self.total_correct, force_gt=force_gt, reduce=False)
        Q = out['answers'].numel()
        slot = self._i % len(self._slots)
        self._i += 1
        buf = self._slots[slot]
        if buf is None or buf[0].numel() < Q:
            buf = (torch.empty(Q, dtype=torch.int64).pin_memory(), torch.empty(Q, dtype=torch.float32).pin_memory(), None)
        if buf[2] is not None:
            buf[2].synchronize()                # the result that used this slot `depth` submits ago has been read
        buf[0][:Q].copy_(out['answers'], non_blocking=True)
        buf[1][:Q].copy_(out['reg_output'], non_blocking=True)
        done = torch.cuda.Event()
        done.record(cur)
        self._slots[slot] = (buf[0], buf[1], done)
        return _PendingAnswers(buf[0], buf[1], Q, done, out)


class _PendingAnswers:
    def __init__(self, answers, reg, Q, event, out):
        self._a, self._r, self._q, self._e, self.device_out = answers, reg, Q, event, out

    def result(self):
        """(answers [Q] int64, reg_output [Q] fp32) on the host; blocks until this batch (not the queue behind it) has finished."""
        self._e.synchronize()
        return self._a[:self._q].clone(), self._r[:self._q].clone()
