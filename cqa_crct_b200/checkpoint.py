"""f4 — checkpoint I/O in the reference's on-disk layout.

The reference writes, once per epoch (CRCT/train.py:284-291),

    torch.save({'model_state_dict': crct_model.module.state_dict(), 'scheduler_state_dict': scheduler.state_dict(),
                'optimizer_state_dict': optimizer.state_dict(), 'iter_id': step_iter_id + 1},
               save_path/'plotqa_encoder_<epoch>_<iter>.ckpt')

and reads it back either for fine-tuning / evaluation (weights only, keys filtered against the model,
CRCT/train.py:91-103, CRCT/evaluation.py:30-42; a bare state dict such as the published `crct.ckpt` is accepted too) or
to continue a run (`--continue`: weights + optimizer + scheduler + iteration counter, the epoch parsed from the file
name, CRCT/train.py:104-127).  This module keeps exactly that dict layout, key set and file-name convention, so
checkpoints move between the two implementations unchanged.

Nothing is staged through Python copies: the model's state dict entries are views of the flat fp32 arena and the
optimizer's `exp_avg` / `exp_avg_sq` entries are views of the two flat moment arenas (`FusedAdamW.state_dict`);
`torch.save` streams those storages.  (torch serialises a shared storage once, so the file holds three flat
buffers plus the key table, not 561 + 2 x 525 separate tensors.)
"""
from __future__ import annotations

import os
import re
from typing import Dict, Optional, Tuple

import torch


def checkpoint_name(epoch: int, iter_id: int) -> str:
    """CRCT/train.py:282."""
    return 'plotqa_encoder_%d_%d.ckpt' % (epoch, iter_id)


def epoch_of(path: str) -> int:
    """`--continue` derives the epoch to resume at from the file name (CRCT/train.py:105)."""
    m = re.match(r'plotqa_encoder_(\d+)_(\d+)\.ckpt$', os.path.basename(path))
    if not m:
        raise ValueError(f'{path}: not a plotqa_encoder_<epoch>_<iter>.ckpt name (CRCT/train.py:105 parses it)')
    return int(m.group(1))


def _unwrap(model):
    return getattr(model, 'module', model)


def save_checkpoint(save_path: str, epoch: int, iter_id: int, model, optimizer, scheduler, extra: Optional[dict] = None) -> str:
    """Writes save_path/plotqa_encoder_<epoch>_<iter_id>.ckpt (CRCT/train.py:282-291); call on rank 0 only."""
    os.makedirs(save_path, exist_ok=True)
    path = os.path.join(save_path, checkpoint_name(epoch, iter_id))
    payload = {'model_state_dict': _unwrap(model).state_dict(), 'scheduler_state_dict': scheduler.state_dict(),
               'optimizer_state_dict': optimizer.state_dict(), 'iter_id': iter_id}
    if extra:
        payload.update(extra)
    tmp = path + '.tmp'
    torch.save(payload, tmp)
    os.replace(tmp, path)
    return path


def load_weights(model, ckpt, device='cpu') -> int:
    """Weights-only load (CRCT/train.py:91-103, CRCT/evaluation.py:30-42): accepts a path or a loaded object, a full
    checkpoint dict or a bare state dict, keeps only the keys the model has, asserts at least one was transferred.
    Returns the number of keys transferred."""
    pretrained = torch.load(ckpt, map_location=device, weights_only=False) if isinstance(ckpt, (str, os.PathLike)) else ckpt
    if 'model_state_dict' in pretrained:
        pretrained = pretrained['model_state_dict']
    enc = _unwrap(model)
    model_dict = enc.state_dict()
    picked = {k: v for k, v in pretrained.items() if k in model_dict}
    assert len(picked) > 0, 'no checkpoint key matches the model (CRCT/train.py:99)'
    for k, v in picked.items():
        if tuple(v.shape) != tuple(model_dict[k].shape):
            raise ValueError(f'{k}: checkpoint shape {tuple(v.shape)} != model shape {tuple(model_dict[k].shape)}')
    model_dict.update(picked)
    enc.load_state_dict(model_dict)
    return len(picked)


def resume(model, optimizer, scheduler, path: str, device='cpu') -> Tuple[int, int, Dict]:
    """`--continue` (CRCT/train.py:104-127): weights + optimizer moments + scheduler position.  Returns
    (cont_epoch, start_iter_id, the remaining payload entries such as 'loss_avg')."""
    payload = torch.load(path, map_location=device, weights_only=False)
    load_weights(model, {'model_state_dict': payload['model_state_dict']}, device)
    optimizer.load_state_dict(payload['optimizer_state_dict'])
    sched_sd = dict(payload['scheduler_state_dict'])
    sched_sd.setdefault('last_epoch', payload['iter_id'])          # train.py:121-123: last_epoch = iter_id, then load_state_dict
    scheduler.load_state_dict(sched_sd)
    rest = {k: v for k, v in payload.items()
            if k not in ('model_state_dict', 'optimizer_state_dict', 'scheduler_state_dict', 'iter_id')}
    return epoch_of(path) + 1, int(payload['iter_id']), rest
