"""Benchmark of the CRCT question-answering hot path (BASELINE.json metric: train samples/s at B=80 per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|eval|stress]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A step = one pass of the hot path over one synthetic PlotQA-shaped batch: forward + backward (+ bucketed gradient
all-reduce overlapped with backward for N > 1) + fused AdamW, dropout ON, B = 80 sequences per GPU (weak scaling).
  value : whole-job samples/s with the batches already resident in HBM (CUDA events, max over ranks)
  e2e   : same metric through the reference-facing API (`glue_forward(model, batch, params)`) with HOST batches:
          pinned host -> device copies of every input and a device -> host read of the loss inside the timed region
  roofline : the tcgen05 GEMM kernel (dominant: 97.7 % of the FLOPs): algorithmic FLOPs (valid rows only — the packed
          layout does not run the padding) of all GEMM launches of one step / their summed durations, against the measured
          bf16 peak in MEASURED_PEAKS.json.  Durations: CUDA events recorded INSIDE a captured single-stream replay of the
          step (event-record graph nodes around every GEMM node: no host enqueue gaps), empty-pair overhead calibrated in
          the same graph; the ncu launch list of the same step is committed under profiles/ and must agree within 5 %.
  cpu_baseline : the reference's own PyTorch module (oracle/_ref, staged there by __graft_entry__.build(); kind "reference")
          or, where that copy is absent, its CPU restatement (oracle/, kind "port"), timed on the host cores on a bounded
          sample of the same workload (rank 0, N = 1 only)
  workloads : (N = 1, --workload train) the other two BASELINE configs — eval B = 512 and the stress shape T = 248 / R = 88 —
          measured in the same process with their own value / e2e / roofline
`--impl reference` times the CPU path alone (all host threads) and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line, the JSON result: everything else any library writes to file descriptor 1 (NCCL prints its
# version banner there) is sent to stderr; the result line goes to the saved descriptor.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_RESULT_FD, (json.dumps(line) + '\n').encode())


import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402

CFG = os.path.join(ROOT, 'cqa_crct_b200', 'config', 'vilbert.json')
WORKLOADS = {   # name -> (B per GPU, T, R, train?)  — BASELINE.json configs[1] / [3] / [4]
    'train': (80, 124, 44, True),
    'eval': (512, 124, 44, False),
    'stress': (80, 248, 88, True),
}
FWD_GFLOP_PER_SAMPLE = {(124, 44): 40.396, (248, 88): 82.547}


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get('bf16_tflops_sustained', 1365.6), d.get('bf16_tflops', 1624.7), d.get('hbm_gbs', 6548.8), 'measured'
    return 1400.0, 1590.0, 6650.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        pw = [float(r[2]) for r in self.rows if len(r) >= 7 and r[2].replace('.', '').isdigit()]
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'power_w_max': max(pw) if pw else None, 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
REF_DIR = os.path.join(ROOT, 'oracle', '_ref')           # __graft_entry__.build() stages the reference's own files here (git-ignored)


def _reference_arm_available():
    return os.path.isfile(os.path.join(REF_DIR, 'CRCT', 'backbone', 'vilbert.py'))


def cpu_reference_run(B, T, R, train, steps, warmup, budget_s=270.0):
    """The reference on the host cores, fp32, all threads.  With oracle/_ref present: the UNMODIFIED reference module through
    its own `encoder_decorator.forward`, dropout on, `loss.backward()`, the reference's `get_optimizer` AdamW + schedule
    (CRCT/train.py:167-215 without autocast: CPU) on full B-sequence batches whenever `steps + warmup` of them fit the time
    budget ("same_config"); otherwise (or without the copy) a bounded sample of `bs` sequences per step.
    Returns (samples/s, bs, seconds per step, threads, kind, note)."""
    from cqa_crct_b200.spec import ModelConfig, synth_state_dict
    from cqa_crct_b200.synthetic import default_params, make_batch
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    cfg = ModelConfig(CFG)
    sd = synth_state_dict(cfg, 228, 0, 'mild')
    if _reference_arm_available():
        os.environ['CRCT_REFERENCE_ROOT'] = REF_DIR
        from oracle import ref_shim
        ref_shim.REFERENCE_ROOT = REF_DIR
        params = default_params(CFG, max_seq_len=T, max_vis_features=R, L1=True)
        params.update(lr=2e-5, image_lr=2e-5, wd=0.01)
        enc = ref_shim.RefEncoder(CFG, params, seed=0)
        m = enc.module
        m.bert_pretrained.load_state_dict(sd, strict=True)
        opt = sched = None
        if train:
            m.train()
            cwd = os.getcwd()
            os.chdir(os.path.join(REF_DIR, 'CRCT'))          # utils.get_optimizer opens 'config/language_weights.json' relatively
            try:
                import utils as ref_utils
                opt = ref_utils.get_optimizer(params, m)
                sched = ref_utils.WarmupLinearScheduleNonZero(opt, warmup_steps=3000, t_total=200000, min_lr=1.3e-5)
            finally:
                os.chdir(cwd)
        else:
            m.eval()

        def one(bs, seed):
            batch = make_batch(bs, T, R, cfg.v_feature_size, seed=seed)
            t0 = time.perf_counter()
            if train:
                loss = enc.glue_forward(m, batch, params)[0]
                loss.backward()
                opt.step()
                opt.zero_grad()
                sched.step()
            else:
                with torch.no_grad():
                    enc.glue_forward(m, batch, params, evaluation=True)
            return time.perf_counter() - t0
        kind = 'reference'
        note = 'UNMODIFIED reference module (oracle/_ref), encoder_decorator.forward' + (', dropout on, backward, get_optimizer AdamW + schedule' if train else ', eval')
    else:
        from oracle import crct_oracle as O
        ocfg = O.Config(cfg.__dict__)

        def one(bs, seed):
            batch = make_batch(bs, T, R, cfg.v_feature_size, seed=seed)
            t0 = time.perf_counter()
            if train:
                out, cache = O.forward(sd, ocfg, batch, train=True, l1=True)
                O.backward(cache)
            else:
                with torch.no_grad():
                    O.forward(sd, ocfg, batch, train=False, keep_cache=False)
            return time.perf_counter() - t0
        kind = 'port'
        note = 'oracle port (no dropout, no optimizer step)'

    nprobe = min(8, B)
    t_probe = one(nprobe, 1)                              # calibrate: seconds for 8 sequences (includes first-touch)
    t_probe = min(t_probe, one(nprobe, 2))
    per_seq = t_probe / nprobe
    total_steps = steps + warmup
    bs = int(max(2, min(B, budget_s / max(1, total_steps) / per_seq)))
    if bs >= 0.8 * B:
        bs = B                                            # close enough: run the full batch (same_config)
    for i in range(warmup):
        one(bs, 10 + i)
    times = [one(bs, 100 + i) for i in range(steps)]
    sec = sum(times) / len(times)
    return bs / sec, bs, sec, threads, kind, note


# ------------------------------------------------------------------------------------------------ GPU arm
def describe(name, B, T, R, train, gpus):
    return {'workload': f'CRCT {"train step (fwd+bwd+allreduce+AdamW, dropout on, L1 regression loss)" if train else "eval forward (hybrid head argmax + regression)"}'
                        f', B={B}/GPU, T={T}, R={R}, full vilbert.json model (252.7M params), random-init weights',
            'per_gpu_batch': B, 'global_batch': B * max(1, gpus), 'text_len': T, 'regions': R, 'parallelism': f'dp{max(1, gpus)}',
            'rows': 'var-len packed: only valid tokens / regions run (lengths U[48..T] / U[4..R] per sequence, device-side counts)',
            'l2': 'working set per step (0.5 GB bf16 weights + >3 GB activations) exceeds the 126 MB L2; inputs rotate over 4 batches'}


def run_workload(name, args, rank, world, local_rank, steps, warmup, with_clocks=True):
    """One BASELINE workload on this process's GPU: returns the fields of the JSON line (rank 0) or None."""
    from cqa_crct_b200 import _lib as L
    from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward
    from cqa_crct_b200.optim import FusedAdamW, WarmupLinearScheduleNonZero
    from cqa_crct_b200.parallel import DistributedDataParallel
    from cqa_crct_b200.synthetic import default_params, make_batch
    B, T, R, train = WORKLOADS[name]
    config = describe(name, B, T, R, train, args.gpus)
    dev = torch.device('cuda', local_rank)
    params = default_params(CFG, device=str(dev), max_seq_len=T, max_vis_features=R, L1=True, overlap_optimizer=args.opt_overlap,
                            varlen=not args.padded, pipeline_optimizer=not args.no_opt_pipeline, shard_optimizer=not args.no_shard_optimizer)
    torch.manual_seed(0)
    enc = VisualDialogEncoder(params).to(dev)
    model = DistributedDataParallel(enc, bucket_cap_mb=args.bucket_mb) if world > 1 else enc
    opt = sched = None
    if train:
        enc.train()
        opt = FusedAdamW(model, lr=2e-5, image_lr=2e-5, weight_decay=0.01)
        sched = WarmupLinearScheduleNonZero(opt, warmup_steps=3000, t_total=200000, min_lr=1.3e-5)
    else:
        enc.eval()
    if train:
        host = [make_batch(B, T, R, 1024, seed=1234 + 17 * rank + i) for i in range(4)]
        pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
        h2d_bytes = sum(v.numel() * v.element_size() for k, v in host[0].items() if k != 'needs_reg')
        d2h_bytes = 4
    else:
        # evaluation batches are QUESTIONS with their candidate answers (CRCT/evaluation.py:231-317): B candidate sequences
        # = 16 questions x 32 candidates; visual tensors once per question (cqa_crct_b200.evaluate, f3)
        from cqa_crct_b200.evaluate import evaluate_batch, to_device
        from cqa_crct_b200.synthetic import make_question_batch
        NQ = 16
        host = [make_question_batch(NQ, T, R, 1024, seed=1234 + 17 * rank + i, total=B) for i in range(4)]
        pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
        resident = [to_device(b, dev) for b in host]
        h2d_bytes = sum(v.numel() * v.element_size() for k, v in host[0].items()) + (B + NQ + 1) * 8
        d2h_bytes = NQ * 8 + NQ * 4
        config['questions_per_batch'] = NQ

    gstep = None
    captured_launches = 0
    if train and not args.no_graph:
        from cqa_crct_b200.graph import GraphedTrainStep
        gstep = GraphedTrainStep(model, opt, params, resident[0], scheduler=sched, warmup_steps=2)
        captured_launches = gstep.launches_per_step
        if world > 1:
            config['exchange'] = ('per bucket: reduce-scatter (fp32 avg) -> AdamW on the 1/N shard -> all-gather of the fp32 masters -> bf16 re-cast, on a side stream'
                                  if gstep.shard_optimizer else 'per bucket: all-reduce (fp32 avg) -> replicated AdamW' if gstep.pipeline_optimizer
                                  else 'per bucket: all-reduce (fp32 avg); whole-arena AdamW after the last one')

    def step(batch, read_loss=False):
        if gstep is not None:                                # captured step: copy inputs into the static buffers, replay
            loss = gstep.step(batch)
            return float(loss) if read_loss else None
        if train:
            opt.zero_grad()
            loss = glue_forward(model, batch, params)[0]
            loss.backward()
            opt.step()
            sched.step()
            return float(loss) if read_loss else None
        out = evaluate_batch(model, batch, params, eval_batch_size=B)       # host batch: copied to the device in here
        if read_loss:                                                        # per-question answers + regression values to the host
            return out['answers'].cpu(), out['reg_output'].cpu()
        return None

    eval_pipe = None
    if not train and not args.no_eval_pipeline:
        from cqa_crct_b200.evaluate import EvalPipeline
        eval_pipe = EvalPipeline(model, params, eval_batch_size=B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def pipelined(batches, n):
        """e2e loop of the captured step: the next batch's host->device copies are staged while the current step runs,
        every step's loss is read on the host one step late (so the queue never drains); the last loss is read before
        the region closes.  Evaluation: the same overlap through `evaluate.EvalPipeline` (H2D of batch i+1 under the kernels
        of batch i, answers + regression values of batch i read after batch i+1 has been enqueued)."""
        prev = None
        if not train:
            for i in range(n):
                h = eval_pipe.submit(batches[i % 4])
                if prev is not None:
                    prev.result()
                prev = h
            return prev.result()
        gstep.prefetch(batches[0])
        for i in range(n):
            h = gstep.step_async()
            gstep.prefetch(batches[(i + 1) % 4])
            if prev is not None:
                prev.item()
            prev = h
        return prev.item()

    def timed(batches, read_loss):
        pipe = read_loss and (gstep is not None or eval_pipe is not None)
        if pipe:
            pipelined(batches, warmup)
        else:
            for i in range(warmup):
                step(batches[i % 4], read_loss)
        barrier()
        sampler = ClockSampler(local_rank) if (rank == 0 and with_clocks) else None
        if sampler:
            sampler.start()
        l0 = L.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if pipe:
            pipelined(batches, steps)
        else:
            for i in range(steps):
                # e2e (read_loss): the batch is HOST memory; glue_forward / the encoder copy every input to the device
                # inside this region and the loss is read back every step
                step(batches[i % 4], read_loss)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = (L.LAUNCHES - l0) + (0 if gstep is None else captured_launches * steps)      # replayed graph launches + what ran outside the graphs
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, launches, clocks

    ms, launches, clocks = timed(resident, False)
    ms_step = ms / steps
    if args.trace and gstep is not None and world > 1:            # per-bucket timeline of one step (rank 0 writes it)
        barrier()
        tr = gstep.trace_step(resident[0])
        barrier()
        if rank == 0:
            os.makedirs(os.path.dirname(os.path.abspath(args.trace)), exist_ok=True)
            json.dump(dict(tr, n_gpus=world, ms_per_step_timed=ms_step), open(args.trace, 'w'), indent=1)
    value = B * world / (ms_step / 1e3)
    e2e = None
    if not args.no_e2e:
        ms2, _, _ = timed(pinned, True)
        e2e = {'value': B * world / (ms2 / steps / 1e3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes,
               'ms_per_step': ms2 / steps}

    # ---- roofline of the dominant kernel: every GEMM launch of ONE more step
    roof = None
    if rank == 0:
        sustained, burst, hbm, src = load_peaks()
        pairs, meta = [], []
        orig = L.gemm
        in_graph = [False]

        def timed_gemm(A, Bm, D, **kw):
            a = torch.cuda.Event(enable_timing=True, external=in_graph[0])
            b = torch.cuda.Event(enable_timing=True, external=in_graph[0])
            a.record()
            orig(A, Bm, D, **kw)
            b.record()
            pairs.append((a, b))
            meta.append((kw['M'], kw['N'], kw['K'], kw.get('a_major', 0), kw.get('rows_dev'), kw.get('epilogue', L.EPI_BIAS),
                         kw.get('aux') is not None, kw.get('D2') is not None))

        orig_grouped = L.gemm_wgrad_grouped

        def timed_grouped(problems):
            a = torch.cuda.Event(enable_timing=True, external=in_graph[0])
            b = torch.cuda.Event(enable_timing=True, external=in_graph[0])
            a.record()
            orig_grouped(problems)
            b.record()
            pairs.append((a, b))
            meta.append([(q.M, q.N, q.K, 1, wg_rows.get(q.a_rows_dev), L.EPI_F32, False, False) for q in problems])

        wg_rows = {}                              # device pointer of a row count -> its tensor (grouped problems carry raw pointers)
        orig_args = L.gemm_args

        def tracking_args(A, Bm, D, **kw):
            if kw.get('rows_dev') is not None:
                wg_rows[kw['rows_dev'].data_ptr()] = kw['rows_dev']
            return orig_args(A, Bm, D, **kw)

        def null_pairs(n):
            out = []
            for _ in range(n):
                a = torch.cuda.Event(enable_timing=True, external=in_graph[0])
                b = torch.cuda.Event(enable_timing=True, external=in_graph[0])
                a.record(); b.record()
                out.append((a, b))
            return out

        import cqa_crct_b200.encoder as E
        require = getattr(model, 'require_sync', None)
        if require is not None:
            model.require_sync = False            # rank 0 alone runs this extra step: no collective
        lanes_were = enc.overlap_streams
        enc.overlap_streams = False               # one stream: a GEMM's event pair must not span kernels of the other lanes
        hook_were, enc.grad_ready_hook = enc.grad_ready_hook, None
        seg_were, enc.segment_ranges = enc.segment_ranges, False
        timing = None
        nulls, graph_keep = [], None
        try:
            L.gemm = E.L.gemm = timed_gemm
            L.gemm_wgrad_grouped, L.gemm_args = timed_grouped, tracking_args
            if train and not args.no_graph_timing:
                try:                              # events as graph nodes: no host in the loop
                    in_graph[0] = True
                    side = torch.cuda.Stream(device=dev)
                    side.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(side):
                        g = torch.cuda.CUDAGraph()
                        g.capture_begin()
                        enc.zero_grad()
                        for _ in enc.train_step_stages(resident[0], 1.0, 1.0):
                            pass
                        nulls = null_pairs(64)
                        g.capture_end()
                        for _ in range(3):
                            g.replay()
                    torch.cuda.synchronize()
                    graph_keep = g
                    _ = pairs[0][0].elapsed_time(pairs[0][1])
                    timing = ('CUDA events recorded as graph nodes around every GEMM node of a captured single-stream replay of the step '
                              '(no host enqueue gaps; the timed steps overlap the text lane, the visual lane and the weight gradients on three streams)')
                except Exception as exc:          # external events unavailable: fall back to host-enqueued events
                    sys.stderr.write(f'[bench] in-graph GEMM timing unavailable ({type(exc).__name__}: {exc}); using host-enqueued events\n')
                    pairs.clear(); meta.clear(); nulls = []
                    in_graph[0] = False
                    torch.cuda.synchronize()
            if timing is None:
                in_graph[0] = False
                if train:
                    enc.zero_grad()
                    glue_forward(enc, resident[0], params)[0].backward()
                else:
                    evaluate_batch(enc, resident[0], params, eval_batch_size=B)
                nulls = null_pairs(200)
                torch.cuda.synchronize()
                timing = 'CUDA events around every GEMM launch of one extra step enqueued from Python on ONE stream'
        finally:
            L.gemm = E.L.gemm = orig
            L.gemm_wgrad_grouped, L.gemm_args = orig_grouped, orig_args
            enc.overlap_streams = lanes_were
            enc.grad_ready_hook, enc.segment_ranges = hook_were, seg_were
            if require is not None:
                model.require_sync = True
        overhead_ms = statistics.median(a.elapsed_time(b) for a, b in nulls)
        raw_ms = sum(a.elapsed_time(b) for a, b in pairs)
        gemm_ms = max(raw_ms - overhead_ms * len(pairs), 0.5 * raw_ms)
        flops = gbytes = 0.0
        n_problems = sum(len(m_) if isinstance(m_, list) else 1 for m_ in meta)
        flat = [q for m_ in meta for q in (m_ if isinstance(m_, list) else [m_])]
        for M_, N_, K_, a_major, rows_dev, epi, has_aux, has_d2 in flat:
            if rows_dev is not None:              # packed rows: the kernel runs the device-side count, not the allocation
                r = int(rows_dev)
                if a_major:
                    K_ = min(K_, r)
                else:
                    M_ = min(M_, r)
            flops += 2.0 * M_ * N_ * K_
            mn = M_ * N_                          # operands once + output (+ residual / multiplier read, + GELU' write)
            res32 = epi == L.EPI_BIAS_RES_F32
            gbytes += 2.0 * (M_ * K_ + N_ * K_) + mn * (4.0 if epi in (L.EPI_F32, L.EPI_BIAS_RES_F32) else 2.0) \
                + ((4.0 if res32 else 2.0) * mn if has_aux else 0.0) + (2.0 * mn if has_d2 else 0.0)
        dump = os.environ.get('CRCT_BENCH_DUMP_GEMMS')       # per-launch table for tools/gemm_table.py
        if dump:
            rows_ = []
            for (a, b), m_ in zip(pairs, meta):
                grp = m_ if isinstance(m_, list) else [m_]
                us = (a.elapsed_time(b) - overhead_ms) * 1e3
                tot_f = sum(2.0 * q[0] * q[1] * q[2] for q in grp)
                for (M_, N_, K_, a_major, rows_dev, epi, has_aux, has_d2) in grp:      # a grouped launch's time is split by nominal FLOPs
                    r = int(rows_dev) if rows_dev is not None else None
                    rows_.append({'M': M_, 'N': N_, 'K': K_, 'a_major': a_major, 'rows': r, 'epi': epi, 'us': us * (2.0 * M_ * N_ * K_) / tot_f,
                                  'grouped': len(grp)})
            json.dump({'workload': name, 'overhead_us': overhead_ms * 1e3, 'launches': rows_}, open(f'{dump}.{name}.json', 'w'))
        del graph_keep
        achieved = flops / (gemm_ms / 1e3) / 1e12
        traffic, traffic_src = None, None         # DRAM bytes per GEMM launch from the committed ncu capture of this workload
        for tp in ('r02_gemm_dram_traffic.json', 'r01_gemm_dram_traffic.json'):
            tp = os.path.join(ROOT, 'profiles', tp)
            if os.path.exists(tp):
                td = json.load(open(tp))
                if td.get('workload') == name and td.get('gemm_launches_per_step') == len(pairs):
                    traffic, traffic_src = td['dram_bytes_per_launch'], os.path.basename(tp)
                    break
        padded_tflop = FWD_GFLOP_PER_SAMPLE.get((T, R), 40.396) * (3 if train else 1) * B / 1e3
        roof = {'bound': 'tensor', 'kernel': 'gemm_tcgen05_kernel', 'achieved': achieved, 'peak': sustained, 'unit': 'TFLOP/s',
                'frac': achieved / sustained, 'traffic': traffic, 'traffic_unit': f'DRAM bytes per launch (ncu, profiles/{traffic_src})',
                'algorithmic_bytes_per_launch': gbytes / max(1, len(pairs)), 'peak_source': f'{src} (bf16_tflops_sustained; burst {burst})',
                'launches_per_step': len(pairs), 'gemm_problems_per_step': n_problems, 'timing': timing, 'gemm_ms_per_step': gemm_ms, 'gemm_ms_per_step_raw': raw_ms,
                'event_pair_overhead_us': overhead_ms * 1e3, 'gemm_share_of_step': gemm_ms / ms_step,
                'algorithmic_tflop_per_step': flops / 1e12, 'padded_layout_tflop_per_step': padded_tflop,
                'executed_tflops_whole_step': flops / 1e12 / (ms_step / 1e3),
                'padded_equivalent_tflops_whole_step': padded_tflop / (ms_step / 1e3)}
    if world > 1:
        dist.barrier()
    out = None
    if rank == 0:
        out = {'metric': 'crct_train_samples_per_sec' if train else 'crct_eval_sequences_per_sec', 'value': value, 'unit': 'samples/s',
               'ms_per_step': ms_step, 'config': dict(config, launch='one CUDA graph per step (captured glue_forward + backward + AdamW)' if gstep is not None else 'per-kernel launches from Python'),
               'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'gpu_launches_per_step': launches / steps, 'roofline': roof}
        if not train:                         # BASELINE's "eval Q/s": questions (with all their candidate answers) per second
            out['questions_per_sec'] = value * config['questions_per_batch'] / B
            if e2e is not None:
                e2e['questions_per_sec'] = e2e['value'] * config['questions_per_batch'] / B
    del gstep, model, enc, opt, resident, pinned
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='train', choices=list(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-sub', action='store_true', help='skip the eval / stress sub-results of the default train line')
    ap.add_argument('--padded', action='store_true', help="run the reference's padded row layout instead of the packed one (A/B)")
    ap.add_argument('--no-graph-timing', action='store_true', help='time the GEMMs with host-enqueued events (round-1 method)')
    ap.add_argument('--opt-overlap', action='store_true', help='run AdamW under the backward instead of after it (A/B; measured slower)')
    ap.add_argument('--bucket-mb', type=float, default=25.0, help='gradient all-reduce bucket size (fp32 MB)')
    ap.add_argument('--trace', default=None, help='N > 1: write the per-bucket timeline of one step (JSON) to this path')
    ap.add_argument('--no-eval-pipeline', action='store_true', help='eval e2e: synchronous evaluate_batch per batch instead of evaluate.EvalPipeline (A/B)')
    ap.add_argument('--no-shard-optimizer', action='store_true', help='N > 1: all-reduce + replicated AdamW instead of reduce-scatter / sharded AdamW / all-gather (A/B)')
    ap.add_argument('--no-opt-pipeline', action='store_true', help='N > 1: whole-arena AdamW after the last all-reduce (A/B)')
    ap.add_argument('--no-graph', action='store_true', help='enqueue every launch from Python instead of replaying the captured step')
    args = ap.parse_args()
    B, T, R, train = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    metric = 'crct_train_samples_per_sec' if train else 'crct_eval_sequences_per_sec'

    if args.impl == 'reference':
        if rank != 0:
            return
        config = describe(args.workload, B, T, R, train, args.gpus)
        v, bs, sec, threads, kind, note = cpu_reference_run(B, T, R, train, max(1, args.steps), max(0, args.warmup))
        line = {'impl': 'reference', 'metric': metric, 'value': v, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic', 'config': config, 'same_config': bool(bs == B and kind == 'reference'),
                'cpu_baseline': {'value': v, 'unit': 'samples/s', 'cores': threads, 'kind': kind,
                                 'sample': f'{bs} sequences per step of the same shape (T={T}, R={R}), fp32 on the host cores: {note}'},
                'e2e': {'value': v, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return

    if not torch.cuda.is_available():
        raise SystemExit('bench.py --impl ours needs a B200; there is no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from cqa_crct_b200 import _lib as L
    L.device_check()
    head = run_workload(args.workload, args, rank, world, local_rank, args.steps, args.warmup)
    subs = None
    if world == 1 and args.workload == 'train' and not args.no_sub:
        # the two BASELINE configs that have no line of their own: same process, fewer steps, same measurement
        subs = {name: run_workload(name, args, rank, world, local_rank, max(3, args.steps // 2), max(3, args.warmup // 2 + 1), with_clocks=False)
                for name in ('eval', 'stress')}

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        v, bs, sec, threads, kind, note = cpu_reference_run(B, T, R, train, 1, 0, budget_s=25.0)
        cpu = {'value': v, 'unit': 'samples/s', 'cores': threads, 'kind': kind,
               'sample': f'1 step of {bs} sequences of the same shape, fp32 on the host cores ({sec:.1f} s): {note}'}

    if rank == 0:
        line = {'metric': head['metric'], 'value': head['value'], 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
                'data': 'synthetic', 'config': head['config'], 'clocks': head['clocks'], 'e2e': head['e2e'], 'gpu_launches': head['gpu_launches'],
                'gpu_launches_per_step': head['gpu_launches_per_step'], 'roofline': head['roofline'], 'cpu_baseline': cpu}
        if subs is not None:
            line['workloads'] = subs
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
