"""Whole-step CUDA-graph capture (f2: step glue without per-launch host cost).

One training step of the reference loop (CRCT/train.py:167-215: forward glue, `loss.backward()`, optimizer step,
`zero_grad`) is ~640 kernel launches here; enqueuing them from Python costs ~16 ms per step, and the reference
additionally syncs the host six times per step for its statistics (train.py:178-183).  `GraphedTrainStep` captures the
same kernels in the same order (`VisualDialogEncoder.train_step_stages` = the body of `glue_forward` + the hand-written
backward, then `FusedAdamW`) into CUDA graphs and replays them:

  * inputs live in static device buffers; `step(batch)` copies the batch (host-pinned or device) into them;
  * dropout stays random: every replay runs `crct_bump_salt`, and all dropout kernels XOR that device word into their
    call-site seed (include/crct_b200.h), so masks change per step while forward/backward agree;
  * the optimizer's per-step scalars (lr schedule, bias corrections) are read from a 6-float device array that is
    refreshed before the replay;
  * data parallel (model wrapped in `cqa_crct_b200.parallel.DistributedDataParallel`): the step is cut into one graph
    per gradient bucket; after replaying segment i the bucket it finished is all-reduced (NCCL, average) asynchronously
    while segment i+1 replays, so the exchange still overlaps the backward without capturing NCCL inside a graph.

`step()` returns the static loss tensor (overwritten by the next step): read or clone it before stepping again.

Input pipeline (the reference's loop copies the batch and `.item()`s its losses synchronously, train.py:167-183):
`prefetch(host_batch)` stages the NEXT batch's host->device copies on a copy stream while the current step runs, and
`step_async()` returns a handle whose `.item()` waits only for that step's own loss (copied to pinned host memory
behind the step), so a loop that reads the loss of step i after enqueuing step i+1 never drains the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist


def fill_fractions(batch) -> tuple:
    """(text, visual) fraction of valid rows of a batch (encoder_decorator.py:118-120 lengths; image_mask), computed with torch ops
    on whatever device the batch lives on — one read-back, done once before capture."""
    seq_len = torch.gather(batch['sep_indices'], 1, batch['hist_len'].view(-1, 1)).squeeze(1) + 1
    T = batch['tokens'].shape[1]
    ft = float(seq_len.clamp(max=T).sum()) / float(seq_len.numel() * T)
    fv = float((batch['image_mask'] != 0).sum()) / float(batch['image_mask'].numel())
    return max(ft, 1e-3), max(fv, 1e-3)


class GraphedTrainStep:
    def __init__(self, model, optimizer, params: dict, example_batch: Dict[str, torch.Tensor], scheduler=None,
                 warmup_steps: int = 2):
        self.opt, self.params, self.sched = optimizer, params, scheduler
        self.ddp = model if hasattr(model, 'module') else None
        self.enc = enc = getattr(model, 'module', model)
        self.world = self.ddp.world if self.ddp is not None else 1
        dev = enc.arena.w32.device
        self.static = {k: v.to(dev).clone() for k, v in example_batch.items() if k != 'needs_reg'}
        self.nsp_coeff, self.reg_coeff = float(params.get('nsp_loss_coeff', 1.0)), float(params.get('reg_loss_coeff', 1.0))
        optimizer.enable_device_scalars()
        if enc.row_fill_hint is None and enc.varlen:
            enc.row_fill_hint = fill_fractions(example_batch)       # tile shapes are frozen at capture: chosen for this fill
        enc.segment_ranges = self.world > 1      # per-bucket ranges only when the step is cut for the exchange
        self.segments = []                # [(graph, (lo, hi) bucket finished by this segment or None)]
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._staged = None               # device staging copy of the prefetched batch
        self._staged_ready = None         # event: staging copy complete
        self._staged_free = None          # event: staging buffers consumed by the step that used them
        self._loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
        self._loss_slot = 0
        self.launches_per_step = 0
        # optional (single GPU): run the optimizer range by range on its own stream under the rest of the backward.  Measured on
        # B200 (profiles/r01_ab_overlap_prio_s7.txt): 0.10-0.25 ms per step SLOWER than AdamW after the backward — the HBM-bound
        # update takes bandwidth and SM slots from cold-operand GEMMs that are themselves latency-sensitive — so it is off by
        # default; stream priorities (chain above fillers) measured neutral to slightly negative and were removed.
        self.overlap_optimizer = self.world == 1 and bool(params.get('overlap_optimizer', False))
        # data parallel: AdamW bucket by bucket, each launch waiting only for ITS bucket's all-reduce — the update of the early
        # buckets (heads, top blocks) runs under the exchange of the late ones (embeddings: 94 MB that become final only at the
        # very end of the backward), so what stays exposed after the backward is the last bucket's exchange + its update
        # instead of every exchange still in flight + the whole 1.3 ms optimizer pass.
        self.pipeline_optimizer = self.world > 1 and bool(params.get('pipeline_optimizer', True))
        # data parallel, default: SHARDED optimizer.  Every bucket is reduce-scattered instead of all-reduced, each rank runs AdamW on
        # its 1/N shard of the bucket only (moments m, v live on the owner alone), the updated fp32 masters are all-gathered in place
        # and every rank re-casts the bucket's bf16 operand copy.  Same bytes on the wire as the all-reduce (its two halves), but the
        # optimizer pass per rank drops from 30 B/parameter to 28/N + 6 B/parameter, and all of it — update, gather, cast — runs on
        # a side stream under the rest of the backward; what stays exposed after the backward is the LAST bucket's chain only.
        # The fp32 masters stay replicated (bit-identical on every rank); `consolidate_optimizer_state()` gathers m / v for checkpoints.
        self.shard_optimizer = (self.world > 1 and 64 % self.world == 0 and bool(params.get('shard_optimizer', True))
                                and dist.get_backend(self.ddp.pg) == 'nccl')
        if self.shard_optimizer:
            self.pipeline_optimizer = False
        self.rank = dist.get_rank(self.ddp.pg) if self.world > 1 else 0
        self._scratch = torch.zeros(8, dtype=torch.float32, device=dev)
        self._scratch16 = torch.zeros(8, dtype=torch.bfloat16, device=dev)
        self._opt_stream = torch.cuda.Stream(device=dev) if (self.overlap_optimizer or self.shard_optimizer) else None
        self._cast_stream = torch.cuda.Stream(device=dev) if self.shard_optimizer else None
        # optional: exchange a bucket in pieces of <= shard_chunk_mb.  Off by default — measured at 8 GPUs (profiles/r02_ab_8gpu_exchange.txt):
        # every extra collective costs ~0.1-0.2 ms of latency on NCCL's in-order stream, and behind the backward those latencies are
        # exposed (26 pieces: 1.47 ms after the backward; 19 whole buckets: 1.07 ms)
        self.shard_chunk = int(params.get('shard_chunk_mb', 0.0) * (1 << 20) / 4) or (1 << 62)
        self._opt_pending = None
        self.opt_chunk = int(params.get('optimizer_chunk', 24 << 20))      # elements per optimizer launch (~0.13 ms of HBM time)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup_steps)):       # eager warm-up: allocator pools, lazy kernel attributes, NCCL
                optimizer.push_device_scalars()
                self._eager_step()
            torch.cuda.synchronize(dev)
            optimizer.push_device_scalars()
            self._capture()
            optimizer.step_count -= 1                    # capture records the step, it does not run it
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        if scheduler is not None:
            for _ in range(max(1, warmup_steps)):
                scheduler.step()

    # -- one step, eagerly (warm-up): same calls as the captured version, exchange through the wrapper's hook logic
    def _opt_hook(self, lo, hi, streams):
        """`VisualDialogEncoder.async_range_hook`: gradients of [lo, hi) are final once `streams` reach this point.  Ranges arrive
        from the arena's tail; they are merged to >= `opt_chunk` and updated on the optimizer stream, which waits for exactly
        those streams — nothing waits for the optimizer until the end of the step."""
        if self._opt_pending is None:
            self._opt_pending = [lo, hi, list(streams)]
        else:
            self._opt_pending[0] = lo
            self._opt_pending[2] += [s for s in streams if s not in self._opt_pending[2]]
        if self._opt_pending[1] - self._opt_pending[0] >= self.opt_chunk or lo == 0:
            plo, phi, pstreams = self._opt_pending
            self._opt_pending = None
            for s in pstreams:
                self._opt_stream.wait_stream(s)
            with torch.cuda.stream(self._opt_stream):
                self.opt.step_range_captured(plo, phi)

    def _stages(self):
        """One forward + backward; with `overlap_optimizer` the optimizer is launched from inside it (and joined here)."""
        if not self.overlap_optimizer:
            yield from self.enc.train_step_stages(self.static, self.nsp_coeff, self.reg_coeff)
            return
        self.enc.async_range_hook = self._opt_hook
        try:
            yield from self.enc.train_step_stages(self.static, self.nsp_coeff, self.reg_coeff)
        finally:
            self.enc.async_range_hook = None
        torch.cuda.current_stream().wait_stream(self._opt_stream)
        self.enc.arena.mark_bf16_fresh()

    def _optimizer_tail(self):
        if self.pipeline_optimizer or self.shard_optimizer:
            from . import _lib as L
            L.cast_f32_to_bf16(self._scratch, self._scratch16)      # the last segment only anchors the per-bucket updates `step()` enqueues: keep it non-empty
        elif not self.overlap_optimizer:
            self.opt.step_captured()

    # -- sharded optimizer (data parallel): reduce-scatter -> AdamW on the own shard -> all-gather of the masters -> bf16 re-cast
    def shard_bounds(self, lo, hi):
        """This rank's shard [slo, shi) of the bucket [lo, hi) (bucket bounds are multiples of 64 elements, 64 % world == 0)."""
        n = (hi - lo) // self.world
        return lo + self.rank * n, lo + (self.rank + 1) * n

    def _scatter_bucket(self, lo, hi):
        """Start the reduce-scatter (average) of the gradient bucket; in place: the shard's slot inside the bucket receives it."""
        g = self.enc.arena.g32
        slo, shi = self.shard_bounds(lo, hi)
        return dist.reduce_scatter_tensor(g[slo:shi], g[lo:hi], op=dist.ReduceOp.AVG, group=self.ddp.pg, async_op=True), lo, hi

    def _chunks(self, lo, hi):
        """A bucket as exchange units of <= `shard_chunk` elements (multiples of 64 * world); one unit unless params['shard_chunk_mb']
        is set (see __init__ for the measurement)."""
        unit = 64 * self.world
        step = max(unit, self.shard_chunk // unit * unit)
        n = max(1, -(-(hi - lo) // step))
        per = -(-(hi - lo) // n)
        per = -(-per // unit) * unit
        out, a = [], lo
        while a < hi:
            b = min(hi, a + per)
            if hi - b < unit:
                b = hi
            out.append((a, b))
            a = b
        return out

    def _update_shard(self, work, lo, hi, ev=None):
        """Optimizer stream: wait for the piece's reduce-scatter, AdamW on the own shard, start the all-gather of the masters.
        Cast stream: wait for that all-gather, re-cast the piece's bf16 operand copy — so the next piece's AdamW does not queue behind
        this piece's gather."""
        from . import _lib as L
        arena = self.enc.arena
        slo, shi = self.shard_bounds(lo, hi)
        with torch.cuda.stream(self._opt_stream):
            work.wait()
            self.opt.step_range_captured(slo, shi, refresh_bf16=False)
            gather = dist.all_gather_into_tensor(arena.w32[lo:hi], arena.w32[slo:shi], group=self.ddp.pg, async_op=True)
        with torch.cuda.stream(self._cast_stream):
            gather.wait()
            L.cast_f32_to_bf16(arena.w32[lo:hi], arena.w16[lo:hi])
            if ev is not None:
                ev.record()

    def _run_sharded(self, units, on_segment=None, on_bucket=None):
        """`units` yields a bucket (lo, hi) right after the work that finishes it has been enqueued on the compute stream.  A piece's
        update chain is enqueued one piece late, so that the NEXT piece's reduce-scatter sits ahead of this one's all-gather on
        NCCL's (in-order) stream: the gradient exchange never waits behind a parameter gather."""
        cur = torch.cuda.current_stream()
        self._opt_stream.wait_stream(cur)
        self._cast_stream.wait_stream(cur)
        pend = []
        for bucket in units:
            if bucket is None:
                continue
            if on_segment:
                on_segment()
            for lo, hi in reversed(self._chunks(*bucket)):       # tail first, like the backward
                pend.append(self._scatter_bucket(lo, hi))
                while len(pend) > 1:
                    item = pend.pop(0)
                    self._update_shard(*item, ev=on_bucket(*item[1:]) if on_bucket else None)
        while pend:
            item = pend.pop(0)
            self._update_shard(*item, ev=on_bucket(*item[1:]) if on_bucket else None)
        cur.wait_stream(self._opt_stream)
        cur.wait_stream(self._cast_stream)
        self.enc.arena.mark_bf16_fresh()
        self.opt.moments_partial = True          # until consolidate_optimizer_state(): optimizer.state_dict() refuses to run

    def _replay_units(self):
        for g, bucket in self.segments:
            g.replay()
            yield bucket

    def consolidate_optimizer_state(self):
        """Sharded optimizer: all-gather the Adam moments from their owners so that every rank holds the complete m / v (what
        `FusedAdamW.state_dict()` saves — CRCT/train.py:283-288 writes the optimizer state on rank 0).  Collective: call on every rank."""
        if not self.shard_optimizer:
            return
        for bucket in [b for _, b in self.segments if b is not None]:
            for lo, hi in self._chunks(*bucket):
                slo, shi = self.shard_bounds(lo, hi)
                for t in (self.opt.m, self.opt.v):
                    dist.all_gather_into_tensor(t[lo:hi], t[slo:shi], group=self.ddp.pg)
        self.opt.moments_partial = False

    def _eager_step(self):
        self.opt.zero_grad()
        if self.shard_optimizer:
            self._run_sharded(self._buckets(self._stages()))
            self._optimizer_tail()
            return
        works = []
        for lo, hi in self._buckets(self._stages()):
            if self.world > 1:
                works.append((dist.all_reduce(self.enc.arena.g32[lo:hi], op=dist.ReduceOp.AVG, group=self.ddp.pg, async_op=True), lo, hi))
        for w, lo, hi in works:
            w.wait()
            if self.pipeline_optimizer:
                self.opt.step_range_captured(lo, hi)
        if self.pipeline_optimizer:
            self.enc.arena.mark_bf16_fresh()
        self._optimizer_tail()

    def _buckets(self, stages):
        """Merge the finished ranges (descending, contiguous) into buckets of >= the wrapper's bucket size."""
        cap = self.ddp.bucket_elems if self.ddp is not None else 1 << 62
        pend = None
        for lo, hi in stages:
            pend = (lo, hi) if pend is None else (lo, pend[1])
            if pend[1] - pend[0] >= cap:
                yield pend
                pend = None
        if pend is not None:
            yield pend

    def _capture(self):
        from . import _lib as L
        l0 = L.LAUNCHES
        pool = torch.cuda.graph_pool_handle()      # all segments share one memory pool (they replay in capture order)

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool)
            return g

        g = begin()
        self.opt.zero_grad()
        for lo, hi in self._buckets(self._stages()):
            if self.world > 1:                   # cut here: the bucket [lo, hi) is final once this segment has run
                g.capture_end()
                self.segments.append((g, (lo, hi)))
                g = begin()
        self._optimizer_tail()
        g.capture_end()
        self.segments.append((g, None))
        self.loss = self.enc.last_scalars[0:1]
        self.logits, self.reg = self.enc.last_logits, self.enc.last_reg
        self.launches_per_step = L.LAUNCHES - l0

    def prefetch(self, batch: Dict[str, torch.Tensor]):
        """Start copying the next batch (pinned host or device tensors) into device staging buffers on the copy stream;
        the next `step()` / `step_async()` without a batch consumes it."""
        if self._staged is None:
            self._staged = {k: torch.empty_like(v) for k, v in self.static.items()}
        cs = self._copy_stream
        if self._staged_free is not None:
            cs.wait_event(self._staged_free)
        with torch.cuda.stream(cs):
            for k, dst in self._staged.items():
                dst.copy_(batch[k], non_blocking=True)
            self._staged_ready = torch.cuda.Event()
            self._staged_ready.record(cs)

    def step_async(self, batch: Optional[Dict[str, torch.Tensor]] = None):
        """`step()` + an asynchronous copy of the loss to pinned host memory; returns a handle with `.item()`."""
        self.step(batch)
        slot = self._loss_slot
        self._loss_slot = (slot + 1) % self._loss_host.numel()
        self._loss_host[slot:slot + 1].copy_(self.loss, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return _PendingLoss(self._loss_host, slot, ev)

    def trace_step(self, batch: Optional[Dict[str, torch.Tensor]] = None) -> dict:
        """One `step()` with CUDA events on the compute stream: when each graph segment has finished (= the bucket it completes is
        handed to NCCL) and when the compute stream gets past each bucket's all-reduce (then runs that bucket's AdamW).  Times in ms
        from the start of the step; `exposed_ms` = end of step - end of the last backward segment = exchange + update time that
        nothing hides.  Synchronises; for timelines (profiles/), not for throughput numbers."""
        ev = lambda: torch.cuda.Event(enable_timing=True)
        t0 = ev(); t0.record()
        if batch is not None:
            for k, dst in self.static.items():
                dst.copy_(batch[k], non_blocking=True)
        self.opt.push_device_scalars()
        works, seg_end, ready = [], [], []
        if self.shard_optimizer:
            def seg():
                e = ev(); e.record(); seg_end.append(e)

            sizes = []

            def bucket_ev(lo, hi):
                e = ev(); ready.append(e); sizes.append((hi - lo) * 4 / 2 ** 20)
                return e
            self._run_sharded(self._replay_units(), on_segment=seg, on_bucket=bucket_ev)
            t1 = ev(); t1.record()
            torch.cuda.synchronize()
            if self.sched is not None:
                self.sched.step()
            bk = [b for _, b in self.segments if b is not None]
            out = {'step_ms': t0.elapsed_time(t1), 'segment_end_ms': [t0.elapsed_time(e) for e in seg_end],
                   'bucket_done_ms': [t0.elapsed_time(e) for e in ready], 'bucket_mb': sizes,
                   'optimizer': 'sharded: reduce-scatter -> AdamW on 1/N -> all-gather -> bf16 cast, per bucket on a side stream'}
            out['backward_end_ms'] = out['segment_end_ms'][-1]
            out['exposed_ms'] = out['step_ms'] - out['backward_end_ms']
            return out
        for g, bucket in self.segments:
            if bucket is None:
                for w, lo, hi in works:
                    w.wait()
                    if self.pipeline_optimizer:
                        self.opt.step_range_captured(lo, hi)
                    e = ev(); e.record(); ready.append((e, lo, hi))
                if self.pipeline_optimizer:
                    self.enc.arena.mark_bf16_fresh()
            g.replay()
            if bucket is not None:
                e = ev(); e.record(); seg_end.append(e)
                works.append((dist.all_reduce(self.enc.arena.g32[bucket[0]:bucket[1]], op=dist.ReduceOp.AVG, group=self.ddp.pg, async_op=True),
                              bucket[0], bucket[1]))
        t1 = ev(); t1.record()
        torch.cuda.synchronize()
        if self.sched is not None:
            self.sched.step()
        out = {'step_ms': t0.elapsed_time(t1), 'segment_end_ms': [t0.elapsed_time(e) for e in seg_end],
               'bucket_done_ms': [t0.elapsed_time(e) for e, _, _ in ready], 'bucket_mb': [(hi - lo) * 4 / 2 ** 20 for _, lo, hi in ready],
               'optimizer': 'per bucket, behind its all-reduce' if self.pipeline_optimizer else 'whole arena after the last all-reduce'}
        if seg_end:
            out['backward_end_ms'] = out['segment_end_ms'][-1]
            out['exposed_ms'] = out['step_ms'] - out['backward_end_ms']
        return out

    def step(self, batch: Optional[Dict[str, torch.Tensor]] = None):
        """Copy `batch` (or the prefetched batch) into the static inputs, replay the captured step, return the (static)
        loss tensor."""
        if batch is not None:
            for k, dst in self.static.items():
                dst.copy_(batch[k], non_blocking=True)
        elif self._staged_ready is not None:
            cur = torch.cuda.current_stream()
            cur.wait_event(self._staged_ready)
            for k, dst in self.static.items():
                dst.copy_(self._staged[k], non_blocking=True)
            self._staged_free = torch.cuda.Event()
            self._staged_free.record(cur)
            self._staged_ready = None
        self.opt.push_device_scalars()
        if self.shard_optimizer:
            self._run_sharded(self._replay_units())
            if self.sched is not None:
                self.sched.step()
            return self.loss
        works = []
        for g, bucket in self.segments:
            if bucket is None:
                for w, lo, hi in works:
                    w.wait()                     # compute stream waits for this bucket's exchange ...
                    if self.pipeline_optimizer:
                        self.opt.step_range_captured(lo, hi)      # ... and updates it while the later buckets are still on the wire
                if self.pipeline_optimizer:
                    self.enc.arena.mark_bf16_fresh()
            g.replay()
            if bucket is not None:
                works.append((dist.all_reduce(self.enc.arena.g32[bucket[0]:bucket[1]], op=dist.ReduceOp.AVG,
                                              group=self.ddp.pg, async_op=True), bucket[0], bucket[1]))
        if self.sched is not None:
            self.sched.step()
        return self.loss


class _PendingLoss:
    """Loss of one enqueued step: `.item()` blocks until that step (not the queue behind it) has finished."""

    def __init__(self, buf, slot, event):
        self.buf, self.slot, self.event = buf, slot, event

    def item(self) -> float:
        self.event.synchronize()
        return float(self.buf[self.slot])
