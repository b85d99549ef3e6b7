"""Whole-step CUDA-graph capture (f2: step glue without per-launch host cost).

One training step of the reference loop (CRCT/train.py:167-215: forward glue, `loss.backward()`, optimizer step,
`zero_grad`) is ~640 kernel launches here; enqueuing them from Python costs ~16 ms per step, and the reference
additionally syncs the host six times per step for its statistics (train.py:178-183).  `GraphedTrainStep` captures the
same calls (`glue_forward` -> `loss.backward()` -> `FusedAdamW`) once into a CUDA graph and replays it:

  * inputs live in static device buffers; `step(batch)` copies the batch (host-pinned or device) into them;
  * dropout stays random: every replay runs `crct_bump_salt`, and all dropout kernels XOR that device word into their
    call-site seed (include/crct_b200.h), so masks change per step while forward/backward agree;
  * the optimizer's per-step scalars (lr schedule, bias corrections) are read from a 6-float device array that is
    refreshed before the replay;
  * with a `cqa_crct_b200.parallel.DistributedDataParallel` model the bucketed NCCL all-reduces are captured too.

Outputs (`loss`, `nsp_scores`, regression list) are static tensors, overwritten by the next `step`: read or clone
them before stepping again.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .encoder import glue_forward


class GraphedTrainStep:
    def __init__(self, model, optimizer, params: dict, example_batch: Dict[str, torch.Tensor], scheduler=None,
                 warmup_steps: int = 2):
        self.model, self.opt, self.params, self.sched = model, optimizer, params, scheduler
        enc = getattr(model, 'module', model)
        dev = enc.arena.w32.device
        self.static = {k: v.to(dev).clone() for k, v in example_batch.items()}
        optimizer.enable_device_scalars()
        self.stream = torch.cuda.Stream(device=dev)
        self.graph = torch.cuda.CUDAGraph()
        self.stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup_steps)):       # eager warm-up on the side stream: allocator pools, lazy init, NCCL
                optimizer.push_device_scalars()
                self._one_step()
            torch.cuda.synchronize(dev)
            optimizer.push_device_scalars()
            with torch.cuda.graph(self.graph, stream=self.stream):
                self.loss, self.out = self._one_step()
            optimizer.step_count -= 1                    # capture records the step, it does not run it
        torch.cuda.current_stream(dev).wait_stream(self.stream)
        if scheduler is not None:
            for _ in range(max(1, warmup_steps)):
                scheduler.step()

    def _one_step(self):
        self.opt.zero_grad()
        out = glue_forward(self.model, self.static, self.params)
        loss = out[0]
        loss.backward()
        self.opt.step_captured()
        return loss.detach(), out

    def step(self, batch: Optional[Dict[str, torch.Tensor]] = None):
        """Copy `batch` into the static inputs (if given), replay the captured step, return the (static) loss tensor."""
        if batch is not None:
            for k, dst in self.static.items():
                dst.copy_(batch[k], non_blocking=True)
        self.opt.push_device_scalars()
        self.graph.replay()
        if self.sched is not None:
            self.sched.step()
        return self.loss
