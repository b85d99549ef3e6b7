"""K10 — data-parallel gradient exchange, overlapped with the hand-written backward.

Replaces `torch.nn.parallel.DistributedDataParallel(model, device_ids=[gpu], find_unused_parameters=True)`
(CRCT/train.py:139-142, CRCT/evaluation.py:59-61).  The reference relies on c10d's Reducer hooking autograd's
per-parameter AccumulateGrad nodes; here the backward is one hand-written pass that fills a flat fp32 gradient
arena from its tail (heads) to its head (embeddings), so the exchange is simply: every time the backward finishes
a layer block it reports the finished arena range; ranges are merged into buckets of >= `bucket_cap_mb` and
all-reduced (NCCL over NVLink/NVSwitch, average) asynchronously on NCCL's stream while the next block's kernels run.
The 36 never-used tensors sit after `live_end` and are never exchanged (what `find_unused_parameters=True` discovered
by walking the autograd graph every iteration).  One process per GPU; no collective on the forward path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import nn


class DistributedDataParallel(nn.Module):
    def __init__(self, module, device_ids=None, find_unused_parameters=False, bucket_cap_mb: float = 25.0,
                 process_group=None, broadcast_parameters: bool = True):
        super().__init__()
        self.module = module
        self.pg = process_group
        self.bucket_elems = int(bucket_cap_mb * (1 << 20) / 4)
        self.world = dist.get_world_size(self.pg) if dist.is_initialized() else 1
        self._avg = dist.is_initialized() and dist.get_backend(self.pg) == 'nccl'
        self._works, self._pend, self._expect = [], None, None
        self.require_sync = True
        self.buckets_last_step = []
        module.grad_ready_hook = self._on_ready
        module.report_min_elems = self.bucket_elems          # the backward joins its streams once per bucket
        if broadcast_parameters and self.world > 1:
            dist.broadcast(module.arena.w32, src=0, group=self.pg)        # DDP construction semantics: rank 0's weights win

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def state_dict(self, *a, **k):
        return self.module.state_dict(*a, **k)

    def _launch(self, lo, hi):
        g = self.module.arena.g32[lo:hi]
        self.buckets_last_step.append((lo, hi))
        if self._avg:
            self._works.append((dist.all_reduce(g, op=dist.ReduceOp.AVG, group=self.pg, async_op=True), None))
        else:
            self._works.append((dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg, async_op=True), g))

    def _on_ready(self, lo, hi):
        """Called by the encoder's backward with arena ranges in descending order; (None, None) = backward finished."""
        if self.world == 1 or not self.require_sync:
            return
        if lo is None:
            if self._pend is not None:
                self._launch(*self._pend)
                self._pend = None
            for work, g in self._works:
                work.wait()                      # current stream waits for the NCCL stream; no host block for NCCL
                if g is not None:
                    g.div_(self.world)
            self._works, self._expect = [], None
            return
        if self._expect is None:
            self.buckets_last_step = []
        elif hi != self._expect:
            raise RuntimeError(f'gradient ranges must arrive contiguously from the tail: got [{lo},{hi}) after {self._expect}')
        self._expect = lo
        self._pend = (lo, hi) if self._pend is None else (lo, self._pend[1])
        if self._pend[1] - self._pend[0] >= self.bucket_elems:
            self._launch(*self._pend)
            self._pend = None
