"""f1 — optimizer step of the training path: one fused multi-tensor AdamW kernel over the flat arena.

Mirrors `get_optimizer` (CRCT/utils.py:228-249: one param group per tensor, `-lr` for the names listed in
config/language_weights.json and `-image_lr` for the rest, weight decay 0 for names containing 'bias',
'LayerNorm.bias' or 'LayerNorm.weight') with torch.optim.AdamW defaults (betas 0.9/0.999, eps 1e-8), and
`WarmupLinearScheduleNonZero` (CRCT/utils.py:11-29).  The same kernel rewrites the bf16 operand copy, so the next
forward needs no cast pass; the GradScaler of CRCT/train.py:157,212-214 has no bf16 counterpart (fp32 exponent range).
"""
from __future__ import annotations

import json
import os
from typing import Optional

import torch

from . import _lib as L

_NO_DECAY = ('bias', 'LayerNorm.bias', 'LayerNorm.weight')          # CRCT/utils.py:229


class FusedAdamW:
    _DYN_SLOTS = 8

    def __init__(self, model, lr: float = 2e-5, image_lr: float = 2e-5, weight_decay: float = 0.01,
                 betas=(0.9, 0.999), eps: float = 1e-8, language_weights: Optional[str] = None):
        enc = getattr(model, 'module', model)
        self.enc = enc
        arena = enc.arena
        if language_weights is None:
            language_weights = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'config', 'language_weights.json')
        with open(language_weights) as f:
            lang = set(json.load(f))
        self.base_lr = [lr, lr, image_lr, image_lr]            # groups: lang+wd, lang no-wd, image+wd, image no-wd
        self.wd = [weight_decay, 0.0, weight_decay, 0.0]
        self.betas, self.eps = betas, eps
        self.n = arena.live_end
        group = torch.zeros(self.n // 64, dtype=torch.uint8)
        for p in arena.order:
            if not p.live:
                continue
            key = 'bert_pretrained.' + p.name
            gid = (0 if key in lang else 2) + (1 if any(nd in key for nd in _NO_DECAY) else 0)
            o = arena.offsets[p.name]
            group[o // 64:(o + p.numel + 63) // 64] = gid
        dev = arena.w32.device
        self._lang = lang
        self._group_host = group
        self.group = group.to(dev)
        self.m = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.step_count = 0
        self.lr_factor = 1.0
        self.min_lr = 0.0
        self.moments_partial = False    # set by graph.GraphedTrainStep (sharded optimizer): m / v are only current on their owner rank
        self.dyn = None            # device {lr[4], 1-beta1^t, sqrt(1-beta2^t)} for CUDA-graph replay (see graph.py)
        self._dyn_host = None

    def current_lrs(self):
        return [max(b * self.lr_factor, self.min_lr) if b * self.lr_factor > self.min_lr else self.min_lr for b in self.base_lr]

    def base_lr_per_param(self):
        return [self.base_lr[self.group_of(p) if p.live else 0] for _, _, p in self._named()]

    def lr_per_param(self):
        lrs = self.current_lrs()
        return [lrs[self.group_of(p) if p.live else 0] for _, _, p in self._named()]

    def zero_grad(self, set_to_none: bool = False):
        self.enc.zero_grad()

    def step(self, grad_scale: float = 1.0):
        arena = self.enc.arena
        self.step_count += 1
        arena.ensure_device_buffers()
        L.adamw(arena.w32, arena.g32, self.m, self.v, arena.w16, self.group, self.n, self.current_lrs(), self.wd,
                self.betas[0], self.betas[1], self.eps, self.step_count, grad_scale)
        arena.mark_bf16_fresh()

    # ---- CUDA-graph support: the kernel reads lr / bias corrections from a device array refreshed before each replay
    def enable_device_scalars(self):
        dev = self.enc.arena.w32.device
        self.dyn = torch.zeros(6, dtype=torch.float32, device=dev)
        # ring of pinned slots, one event each: the host never rewrites a slot whose host->device copy may still be queued
        # (a loop that enqueues steps without synchronising would otherwise hand step n the scalars of step n + k)
        self._dyn_host = torch.zeros(self._DYN_SLOTS, 6, dtype=torch.float32).pin_memory()
        self._dyn_events = [None] * self._DYN_SLOTS
        self._dyn_slot = 0

    def push_device_scalars(self):
        """Advance the step counter and upload {lr[4], bias corrections} (call right before replaying a captured step)."""
        self.step_count += 1
        lrs = self.current_lrs()
        slot = self._dyn_slot
        self._dyn_slot = (slot + 1) % self._DYN_SLOTS
        if self._dyn_events[slot] is not None:
            self._dyn_events[slot].synchronize()         # the copy issued _DYN_SLOTS pushes ago has left this slot
        h = self._dyn_host[slot]
        h[0], h[1], h[2], h[3] = lrs
        h[4] = 1.0 - self.betas[0] ** self.step_count
        h[5] = (1.0 - self.betas[1] ** self.step_count) ** 0.5
        self.dyn.copy_(h, non_blocking=True)
        ev = self._dyn_events[slot] or torch.cuda.Event()
        ev.record()
        self._dyn_events[slot] = ev

    def step_captured(self, grad_scale: float = 1.0):
        """The launch to put INSIDE a CUDA graph: identical kernel, scalars taken from `self.dyn`."""
        arena = self.enc.arena
        L.adamw(arena.w32, arena.g32, self.m, self.v, arena.w16, self.group, self.n, self.base_lr, self.wd,
                self.betas[0], self.betas[1], self.eps, 1, grad_scale, dyn=self.dyn)
        arena.mark_bf16_fresh()

    def step_range_captured(self, lo: int, hi: int, grad_scale: float = 1.0, refresh_bf16: bool = True):
        """The same update restricted to arena elements [lo, hi), scalars from `self.dyn`: lets the step run the optimizer for the
        part of the arena whose gradients are final while the backward is still producing the rest, and lets a data-parallel rank
        update only ITS SHARD of a bucket (graph.py).  `lo` need not be a multiple of 64: the (lr, weight decay) table is indexed
        through `group_offset`."""
        arena = self.enc.arena
        hi = min(hi, self.n)
        if lo < 0 or hi <= lo:
            raise ValueError(f'optimizer range [{lo}, {hi}) is empty')
        b0 = lo // 64
        L.adamw(arena.w32[lo:hi], arena.g32[lo:hi], self.m[lo:hi], self.v[lo:hi], arena.w16[lo:hi] if refresh_bf16 else None,
                self.group[b0:(hi + 63) // 64], hi - lo, self.base_lr, self.wd, self.betas[0], self.betas[1], self.eps, 1, grad_scale,
                dyn=self.dyn, group_offset=lo - b0 * 64)

    # ---- checkpoint layout (f4): what `torch.optim.AdamW(...).state_dict()` gives for `get_optimizer`'s grouping
    def _named(self):
        """[(index, reference parameter name, spec entry)] in `named_parameters()` order = the reference's param-group
        order (CRCT/utils.py:236-247: one group per tensor; the tied decoder weight appears once)."""
        return [(i, 'bert_pretrained.' + p.name, p) for i, p in enumerate(self.enc.arena.spec)]

    def group_of(self, p):
        o = self.enc.arena.offsets[p.name]
        return int(self._group_host[o // 64])

    def state_dict(self):
        """torch.optim.AdamW layout: `param_groups` = one group per parameter (lr / weight_decay / betas / eps /
        initial_lr, `params: [index]`), `state[index] = {step, exp_avg, exp_avg_sq}` for every tensor that has received
        a gradient.  The reference's 36 never-used tensors have no state entry there either (their .grad stays None
        under `find_unused_parameters=True`, CRCT/train.py:141).  The moment tensors are VIEWS of the flat arenas —
        `torch.save` writes them without a copy through Python."""
        self._require_complete_moments()
        arena = self.enc.arena
        lrs = self.current_lrs()
        groups, state = [], {}
        for i, _, p in self._named():
            gid = self.group_of(p) if p.live else (2 if 'bert_pretrained.' + p.name not in self._lang else 0) + \
                (1 if any(nd in 'bert_pretrained.' + p.name for nd in _NO_DECAY) else 0)
            groups.append({'lr': lrs[gid], 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': self.wd[gid], 'amsgrad': False,
                           'initial_lr': self.base_lr[gid], 'params': [i]})
            if p.live and self.step_count > 0:
                o = arena.offsets[p.name]
                state[i] = {'step': torch.tensor(float(self.step_count)), 'exp_avg': self.m[o:o + p.numel].view(p.shape),
                            'exp_avg_sq': self.v[o:o + p.numel].view(p.shape)}
        return {'state': state, 'param_groups': groups}

    def load_state_dict(self, sd):
        """Accepts the layout above — i.e. also the `optimizer_state_dict` of a reference checkpoint
        (CRCT/train.py:110-119) — and the flat layout of `flat_state_dict()`."""
        if 'exp_avg' in sd:                              # flat
            self.step_count, self.lr_factor, self.min_lr = int(sd['step']), sd['lr_factor'], sd['min_lr']
            self.m.copy_(sd['exp_avg'])
            self.v.copy_(sd['exp_avg_sq'])
            return
        arena = self.enc.arena
        named = self._named()
        groups = sd['param_groups']
        index_of = {}                                    # saved parameter index -> position in named_parameters() order
        pos = 0
        for g in groups:
            for idx in g['params']:
                index_of[idx] = pos
                pos += 1
        if pos != len(named):
            raise ValueError(f'optimizer state has {pos} parameters, the model has {len(named)}')
        steps = set()
        self.m.zero_()
        self.v.zero_()
        for idx, st in sd['state'].items():
            _, _, p = named[index_of[int(idx)]]
            if not p.live:
                continue
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError(f'optimizer state of {p.name}: shape {tuple(st["exp_avg"].shape)} != {tuple(p.shape)}')
            o = arena.offsets[p.name]
            self.m[o:o + p.numel].copy_(st['exp_avg'].reshape(-1))
            self.v[o:o + p.numel].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(float(st['step'])))
        if len(steps) > 1:
            raise ValueError(f'per-parameter step counts differ ({sorted(steps)}): one fused step count is kept')
        self.step_count = steps.pop() if steps else 0
        base = [None] * 4
        for g, (_, _, p) in zip(groups, named):          # base learning rates per group, from initial_lr when present
            if p.live:
                base[self.group_of(p)] = float(g.get('initial_lr', g['lr']))
        self.base_lr = [b if b is not None else old for b, old in zip(base, self.base_lr)]

    def _require_complete_moments(self):
        if self.moments_partial:
            raise RuntimeError('the Adam moments are sharded over the data-parallel ranks (graph.GraphedTrainStep, sharded optimizer): call '
                               'GraphedTrainStep.consolidate_optimizer_state() on EVERY rank before reading the optimizer state')

    def flat_state_dict(self):
        self._require_complete_moments()
        return {'step': self.step_count, 'exp_avg': self.m, 'exp_avg_sq': self.v, 'lr_factor': self.lr_factor, 'min_lr': self.min_lr}


class WarmupLinearScheduleNonZero:
    """CRCT/utils.py:11-29: linear warm-up to the base lr over `warmup_steps`, linear decay to zero at `t_total`,
    clamped from below at `min_lr`."""

    def __init__(self, optimizer: FusedAdamW, warmup_steps: int, t_total: int, min_lr: float = 1.3e-5, last_epoch: int = -1):
        self.opt, self.warmup_steps, self.t_total = optimizer, warmup_steps, t_total
        optimizer.min_lr = min_lr
        self.last_epoch = last_epoch
        self.step()

    def factor(self, step: int) -> float:
        if step < self.warmup_steps:
            return float(step) / float(max(1, self.warmup_steps))
        return max(0.0, float(self.t_total - step) / float(max(1.0, self.t_total - self.warmup_steps)))

    def step(self):
        self.last_epoch += 1
        self.opt.lr_factor = self.factor(self.last_epoch)

    def get_last_lr(self):
        return self.opt.current_lrs()

    def state_dict(self):
        """`_LRScheduler.state_dict()` layout (every attribute but the optimizer), so the `scheduler_state_dict` of a
        reference checkpoint (CRCT/train.py:287) and ours are interchangeable."""
        return {'warmup_steps': self.warmup_steps, 't_total': self.t_total, 'min_lr': self.opt.min_lr,
                'base_lrs': list(self.opt.base_lr_per_param()), 'last_epoch': self.last_epoch, '_step_count': self.last_epoch + 1,
                'verbose': False, '_get_lr_called_within_step': False, '_last_lr': list(self.opt.lr_per_param())}

    def load_state_dict(self, sd):
        self.last_epoch = int(sd['last_epoch'])
        self.warmup_steps = sd.get('warmup_steps', self.warmup_steps)
        self.t_total = sd.get('t_total', self.t_total)
        if 'min_lr' in sd:
            self.opt.min_lr = sd['min_lr']
        self.opt.lr_factor = self.factor(self.last_epoch)
