"""Evaluation entry point (CRCT/evaluation.py:21-66, 126-222, 225-317, 423-491).

    python -m cqa_crct_b200.evaluation -model_config cqa_crct_b200/config/vilbert.json -start_checkpoint run/plotqa_encoder_0_50.ckpt \\
        -eval_batch_size 512 -eval_questions 256

`get_encoder` builds the model and loads a checkpoint exactly as the reference does (weights-only by key intersection,
or the `model_state_dict` of a `-continue` checkpoint).  `plotqa_evaluate` is the reference's batch loop with the model
calls, the per-question selection and the accuracy table on the device (`cqa_crct_b200.evaluate`).  The PlotQA reader is
out of scope (detection stage): question batches come from the seeded synthetic generator in the dataset's layout
(one item per question with all its candidate answers, visual tensors stored once per question).
"""
from __future__ import annotations

import os
import sys
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

from . import checkpoint as ckpt
from .encoder import VisualDialogEncoder
from .evaluate import EvalPipeline, evaluate_batch
from .parallel import DistributedDataParallel
from .synthetic import make_question_batch

ROWS = ('nsp', 'reg_cls', 'reg', 'reg_t', 'total', 'total_t')          # reduce_total_acc row order, evaluation.py:498-517


def get_encoder(params: dict, ckpt_path: str):
    """CRCT/evaluation.py:21-66."""
    device = params['device']
    dialog_encoder = VisualDialogEncoder(params).to(device)
    if ckpt_path:
        ckpt.load_weights(dialog_encoder, ckpt_path, device)            # both branches of :30-53 load the model weights only
    if params.get('ddp') and params.get('world_size', 1) > 1:
        dialog_encoder = DistributedDataParallel(dialog_encoder)        # :56-62 (no gradients flow at evaluation)
    return dialog_encoder


def plotqa_evaluate(batches: Iterable[dict], params: dict, eval_batch_size: int, dialog_encoder, collect: bool = False):
    """CRCT/evaluation.py:195-317 `plotqa_evaluate_DDP` without its reporting side (CSV / histogram / breakdown).
    Returns (total_correct [6,2] float64 on the device, list of per-batch outputs if `collect`)."""
    enc = getattr(dialog_encoder, 'module', dialog_encoder)
    enc.eval()                                                           # :197
    dev = enc.arena.w32.device
    total = torch.zeros(6, 2, dtype=torch.float64, device=dev)          # :211
    outs = []
    force = '_REGS' in params.get('qa_file', '')                         # :288-289
    pipe = EvalPipeline(dialog_encoder, params, eval_batch_size)         # next batch's copies under this batch's kernels; no host read-back
    pipe.total_correct = total
    for qb in batches:
        if qb['id'].shape[0] == 0:                                       # :233-234
            continue
        h = pipe.submit(qb, force_gt=force)
        if collect:
            outs.append(h.device_out)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)                     # :519-521, once for the whole run (the sum is associative)
    return total, outs


def format_table(total: torch.Tensor) -> str:
    t = total.cpu()
    return ' | '.join('{} {:.2f}% ({:.0f}/{:.0f})'.format(n, 100.0 * float(t[i, 0]) / max(float(t[i, 1]), 1.0), float(t[i, 0]), float(t[i, 1]))
                      for i, n in enumerate(ROWS))


def evaluate_synthetic(dialog_encoder, params: dict, n_questions: int = 64, questions_per_batch: int = 16, seed: int = 4321):
    """The post-epoch evaluation of train.py:293-300 on synthetic questions; ranks take disjoint batches
    (DistributedSampler, evaluation.py:150) and the table is summed across them."""
    enc = getattr(dialog_encoder, 'module', dialog_encoder)
    cfg = enc.cfg
    rank, world = params.get('rank', 0), max(1, params.get('world_size', 1))
    n_batches = max(1, n_questions // questions_per_batch)

    def batches():
        for i in range(rank, n_batches, world):
            yield make_question_batch(questions_per_batch, params['max_seq_len'], params['max_vis_features'], cfg.v_feature_size,
                                      seed=seed + i, vocab_size=cfg.vocab_size, max_ans=48)

    total, _ = plotqa_evaluate(batches(), params, params.get('eval_batch_size', 512), dialog_encoder)
    return total


def main(argv: Optional[List[str]] = None):
    from .train import read_command_line, log_line
    params = read_command_line(argv)
    if not torch.cuda.is_available():
        raise SystemExit('cqa_crct_b200.evaluation needs a B200; there is no CPU fallback')
    gpu = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(gpu)
    params['device'] = torch.device('cuda', gpu)
    if params['ddp'] and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=params['device'])       # :127-132
    encoder = get_encoder(params, params['start_checkpoint'])
    log_line(params, "Model's parameters: {}".format(sum(p.numel() for p in encoder.parameters())))   # :141
    total = evaluate_synthetic(encoder, params, n_questions=params['eval_questions'])
    log_line(params, format_table(total))
    return total


if __name__ == '__main__':
    main(sys.argv[1:])
