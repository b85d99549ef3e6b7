"""Host side of the B200-native CRCT question-answering stage.

`VisualDialogEncoder` mirrors the reference wrapper (CRCT/backbone/encoder_decorator.py:9-54) and the
model under it (CRCT/backbone/vilbert.py:1499-1661): same constructor argument (`params` dict), same
`forward(...)` signature and return tuple, same `state_dict()` keys / `named_parameters()` order, same
branch selection by kwargs.  Every arithmetic step is a kernel of libcrct_b200 reached through the C ABI
(include/crct_b200.h); PyTorch only owns memory, streams and the autograd hook that lets
`loss.backward()` (CRCT/train.py:208) drive the hand-written backward.  There is no CPU / eager fallback.

Data layout (all in HBM, caller = PyTorch allocator):
  * parameters: ONE flat fp32 arena (master weights) + ONE flat fp32 gradient arena + ONE flat bf16 operand
    copy, tensors ordered by `spec.arena_order` (forward execution order, fused q|k|v adjacent, dead
    tensors last); every nn.Parameter / .grad is a view into them;
  * activations: bf16 row-major [B*T, H] / [B*R, Hv]; packed projections [rows, 3H]; LayerNorm statistics,
    attention log-sum-exp, head activations and losses in fp32.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _lib as L
from .spec import ModelConfig, P, TIED, arena_offsets, arena_order, param_spec

_SITE = {  # dropout stream ids (one counter-based stream per dropout call site of the reference)
    'emb_t': 1, 'emb_v': 2, 'attn': 3, 'attn_out': 4, 'ffn_out': 5, 'co_attn1': 6, 'co_attn2': 7, 'co_out_v': 8,
    'co_out_t': 9, 'co_ffn_v': 10, 'co_ffn_t': 11, 'cls': 12}


def _seed(step_seed: int, site: str, layer: int = 0) -> int:
    x = (step_seed * 0x9E3779B97F4A7C15 + _SITE[site] * 0xBF58476D1CE4E5B9 + layer * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    return (x * 0xD6E8FEB86659FD93) & 0xFFFFFFFFFFFFFFFF


class ParamArena:
    """Flat fp32 master / fp32 gradient / bf16 operand storage with per-tensor views."""

    def __init__(self, cfg: ModelConfig, categories: int):
        self.spec: List[P] = param_spec(cfg, categories)
        self.order: List[P] = arena_order(cfg, self.spec)
        self.offsets, self.live_end, self.total = arena_offsets(self.order)
        self.by_name: Dict[str, P] = {p.name: p for p in self.spec}
        self.w32 = torch.zeros(self.total, dtype=torch.float32)
        self.g32: Optional[torch.Tensor] = None
        self.w16: Optional[torch.Tensor] = None
        self._w16_version = -1
        self.param_views: List[torch.Tensor] = []      # the nn.Parameters viewing w32 (set by the encoder)

    def view(self, flat: torch.Tensor, name: str) -> torch.Tensor:
        p = self.by_name[name]
        o = self.offsets[name]
        return flat[o:o + p.numel].view(p.shape)

    def fused(self, flat: torch.Tensor, names: List[str], suffix: str) -> torch.Tensor:
        """Zero-copy view of adjacent tensors (q|k|v weights -> [3*out, in], biases -> [3*out])."""
        ps = [self.by_name[n + suffix] for n in names]
        o0 = self.offsets[ps[0].name]
        cur = o0
        for p in ps:
            if self.offsets[p.name] != cur:
                raise RuntimeError(f'{p.name} is not adjacent to its fusion group in the arena')
            cur += p.numel
        shape = (sum(p.shape[0] for p in ps),) + tuple(ps[0].shape[1:])
        return flat[o0:cur].view(shape)

    def ensure_device_buffers(self):
        if self.g32 is None or self.g32.device != self.w32.device:
            self.g32 = torch.zeros_like(self.w32)
        if self.w32.is_cuda and (self.w16 is None or self.w16.device != self.w32.device):
            self.w16 = torch.empty(self.total, dtype=torch.bfloat16, device=self.w32.device)
            self._w16_version = -1

    def version(self):
        """Changes whenever a master weight may have changed: in-place writes to the arena itself AND to any of the
        nn.Parameter views (after `.to()` / `.cuda()` rebinds `prm.data`, a view no longer shares the arena's version counter:
        load_state_dict, torch.optim optimizers and manual `copy_` bump only the parameter's own counter)."""
        return self.w32._version + sum(p._version for p in self.param_views)

    def refresh_bf16(self):
        """Re-cast the operand copy if any master weight changed since the last cast."""
        v = self.version()
        if self._w16_version != v:
            L.cast_f32_to_bf16(self.w32, self.w16)
            self._w16_version = v

    def mark_bf16_fresh(self):
        self._w16_version = self.version()

    def invalidate(self):
        """Force the next forward to re-cast the bf16 operand copy (for writers that bypass torch's version counters)."""
        self._w16_version = -1


class _Node(nn.Module):
    """Plain container used to rebuild the reference's module tree (names only; no arithmetic)."""


def _attach(root: nn.Module, dotted: str, param: nn.Parameter):
    parts = dotted.split('.')
    mod = root
    for part in parts[:-1]:
        if part not in mod._modules:
            mod.add_module(part, _Node())
        mod = mod._modules[part]
    mod.register_parameter(parts[-1], param)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Lanes:
    """The two dataflow lanes of the two-stream encoder: text kernels run on the caller's stream, visual kernels on a
    side stream; the lanes meet only where data crosses (co-attention, heads).  With one stream both lanes are the
    caller's stream and every wait is a no-op."""

    def __init__(self, text, vis):
        self.t, self.v = text, vis
        self.split = vis is not text

    def vis(self):
        return torch.cuda.stream(self.v) if self.split else _NullCtx()

    def v_wait_t(self):
        if self.split:
            self.v.wait_stream(self.t)

    def t_wait_v(self):
        if self.split:
            self.t.wait_stream(self.v)

    def meet(self):
        self.v_wait_t()
        self.t_wait_v()


class _Saved:
    """Activations a forward keeps for its backward."""
    pass


class _Rows:
    """Row layout of one stream (text tokens / image regions) for one pass.

    Padded (the reference's layout, `cu is None`): sample b owns rows [b*L, (b+1)*L) and an additive key mask hides the padding.
    Packed (`cu` = int32[B+1] on the device, csrc/varlen.cu): only the rows whose mask is 1 exist, sample b owns rows
    [cu[b], cu[b+1]); `n` = the device word cu[B] every kernel reads its row count from (buffers are still allocated for B*L
    rows — the host never learns the count, so one captured CUDA graph serves every batch), `src[r]` = padded position of row r."""

    def __init__(self, B, L, cu=None, src=None, mask_add=None, fill=None):
        self.B, self.L, self.cu, self.src, self.mask_add = B, L, cu, src, mask_add
        self.n = cu[B:B + 1] if cu is not None else None
        if self.n is not None:
            # the caller's estimate of the valid fraction -> expected row count, read by the GEMM's tile-shape choice only
            self.n.hint = int(fill * B * L) if fill else 0


class _CrctFunction(torch.autograd.Function):
    """Ties the hand-written forward/backward into autograd: inputs = one anchor parameter, outputs =
    (nsp_loss[1], reg_loss[B]); backward runs the whole CUDA backward and accumulates into the gradient arena."""

    @staticmethod
    def forward(ctx, anchor, enc, saved):
        ctx.enc, ctx.saved = enc, saved
        return saved.nsp_loss, saved.reg_loss

    @staticmethod
    def backward(ctx, d_nsp, d_reg):
        enc, saved = ctx.enc, ctx.saved
        ctx.saved = None
        enc._backward(saved, d_nsp, d_reg)
        return None, None, None


class VisualDialogEncoder(nn.Module):
    """Drop-in for CRCT/backbone/encoder_decorator.py:9 `VisualDialogEncoder(params)`."""

    def __init__(self, params: dict, precision: Optional[str] = None):
        """`precision`: 'bf16' (default; production kernels) or 'fp32' — the CHECK MODE of csrc/check_f32.cu: same schedule,
        same dropout streams, fp32 activations and plain fp32 arithmetic (slow; for parity work, not for training runs).
        Also read from params['precision']."""
        super().__init__()
        precision = precision or params.get('precision', 'bf16')
        if precision not in ('bf16', 'fp32'):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.fp32 = precision == 'fp32'
        self.act = torch.float32 if self.fp32 else torch.bfloat16
        # var-len ("packed") rows: only the tokens / regions whose mask is 1 go through the encoder (csrc/varlen.cu); exact —
        # masked keys have probability 0 in the reference.  The fp32 check mode runs the reference's padded layout.
        self.varlen = bool(params.get('varlen', True)) and not self.fp32
        # (text, visual) expected fraction of valid rows — an ESTIMATE (e.g. of the previous / an example batch) that only steers
        # the GEMM tile shapes; set by graph.GraphedTrainStep / evaluate.evaluate_batch from host-side batch data, or by the user
        self.row_fill_hint = params.get('row_fill_hint', None)
        config_path = params['model_config']
        assert os.path.exists(config_path), "model_config file not found"      # encoder_decorator.py:13
        self.params = params
        self.cfg = ModelConfig(config_path)
        if params.get('dataset', 'plotqa') in ('figure_qa', 'dvqa') or params.get('CE_REG') or params.get('binary_answers'):
            raise NotImplementedError('only the PlotQA regression path (PlotQA_Regressor_v20) is implemented')
        if params.get('mask_prob_img', 0) > 0:
            raise NotImplementedError('mask_prob_img > 0 is not implemented (default 0, CRCT/options.py:33)')
        self.arena = ParamArena(self.cfg, params['categories'])
        self._init_reference_distributions()
        root = _Node()
        self._params_by_name: Dict[str, nn.Parameter] = {}
        for p in self.arena.spec:
            prm = nn.Parameter(self.arena.view(self.arena.w32, p.name))     # all require grad, as in the reference
            self._params_by_name[p.name] = prm
            _attach(root, p.name, prm)
        for alias, src in TIED.items():
            _attach(root, alias, self._params_by_name[src])
        self.bert_pretrained = root
        self.arena.param_views = list(self._params_by_name.values())
        self._step = 0                       # site seeds are launch constants; per-step variation comes from the device salt
        self._salt = None                    # int64[1] on the device, advanced by crct_bump_salt once per training forward
        self.grad_ready_hook = None          # set by cqa_crct_b200.parallel.DistributedDataParallel
        self.overlap_streams = True          # visual lane / text lane on two streams, weight gradients on a third
        self.segment_ranges = False          # set by graph.GraphedTrainStep when it cuts the step at bucket boundaries
        self.report_min_elems = 0            # fine mode: report (= join the lanes) only once this many gradient elements are final
        self.async_range_hook = None         # f(lo, hi, streams): gradients of [lo, hi) are final once `streams` reach this point
                                             # (no stream is joined; used to run the optimizer under the rest of the backward)
        self._x_hold = None                  # tensors read across lanes, kept until the lanes have met again
        self._side_stream = None
        self._wg_stream = None
        self._wg_hold = []
        self._wg_pending = []                # deferred weight-gradient problems [(crct_gemm_t, operands, producer stream)]
        self.group_wgrads = bool(params.get('group_wgrads', True))
        self._tail = {}                      # id(row layout of the stream) -> (last pre-LayerNorm sum, its LayerNorm): read by the heads
        self.train()                         # encoder_decorator.py:17

    # ------------------------------------------------------------------ parameters / devices
    def _init_reference_distributions(self):
        """vilbert.py:1099-1110: N(0, initializer_range) weights, zero biases, unit LayerNorm; regressor keeps
        nn.Linear's default init (it is constructed after `apply(init)`, vilbert.py:1510 vs 1523)."""
        g = torch.Generator().manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        for p in self.arena.spec:
            v = self.arena.view(self.arena.w32, p.name)
            if p.kind in ('w', 'emb'):
                v.normal_(0.0, self.cfg.initializer_range, generator=g)
            elif p.kind == 'lnw':
                v.fill_(1.0)
            elif p.kind in ('rw', 'rb'):
                fan_in = p.shape[1] if p.kind == 'rw' else self.arena.by_name[p.name[:-4] + 'weight'].shape[1]
                v.uniform_(-1.0 / fan_in ** 0.5, 1.0 / fan_in ** 0.5, generator=g)

    def _apply(self, fn, recurse=True):
        new = fn(self.arena.w32)
        if new.dtype != torch.float32:
            raise TypeError('master parameters stay fp32; bf16 operands are derived inside the library')
        old_g = self.arena.g32
        self.arena.w32 = new
        self.arena.g32 = None if old_g is None else fn(old_g)
        self.arena.w16 = None
        for name, prm in self._params_by_name.items():
            prm.data = self.arena.view(self.arena.w32, name)
            if prm.grad is not None:
                prm.grad = self.arena.view(self.arena.g32, name) if self.arena.g32 is not None else None
        return self

    def _bind_grads(self):
        self.arena.ensure_device_buffers()
        for name, prm in self._params_by_name.items():
            if self.arena.by_name[name].live and (prm.grad is None or prm.grad.data_ptr() != self.arena.view(self.arena.g32, name).data_ptr()):
                prm.grad = self.arena.view(self.arena.g32, name)

    def zero_grad(self, set_to_none: bool = False):
        """Gradients live in the arena: zeroing is one memset; `set_to_none` is ignored on purpose."""
        if self.arena.g32 is not None:
            if self.arena.g32.is_cuda:
                L.fill_zero(self.arena.g32[:self.arena.live_end])       # one memset node on the stream (no ATen fill kernels in the step)
            else:
                self.arena.g32[:self.arena.live_end].zero_()

    def flat_parameters(self):
        return self.arena.w32, self.arena.g32, self.arena.w16

    # ------------------------------------------------------------------ small helpers
    def _pass_salt(self, dev):
        """Dropout salt of one forward pass.  The live counter `self._salt` (one device word, seeded from torch's seed and the
        data-parallel RANK — the reference never seeds, so its ranks draw independent masks, CRCT/train.py:55) is advanced
        once per training forward; the pass and its backward read a per-pass SNAPSHOT of it, so two forwards before the
        first backward each recompute their own masks."""
        if self._salt is None or self._salt.device != dev:
            rank = 0
            if torch.distributed.is_available() and torch.distributed.is_initialized():
                rank = torch.distributed.get_rank()
            z = (rank + 1) * 0x9E3779B97F4A7C15 & 0xFFFFFFFFFFFFFFFF          # splitmix64(rank)
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
            z ^= z >> 31
            seed = ((torch.initial_seed() * 0x9E3779B97F4A7C15) ^ (z if rank else 0)) & 0x7FFFFFFFFFFFFFFF
            self._salt = torch.full((1,), seed, dtype=torch.int64, device=dev)
        snap = self._salt
        if self.training:
            snap = torch.empty(1, dtype=torch.int64, device=dev)
            L.bump_salt(self._salt, snap)
        L.SALT = snap
        return snap

    def _wflat(self):        # GEMM operand arena: the bf16 copy, or the fp32 masters themselves in check mode
        return self.arena.w32 if self.fp32 else self.arena.w16

    def _w(self, name):      # GEMM operand view
        return self.arena.view(self._wflat(), name)

    def _p(self, name):      # fp32 master view
        return self.arena.view(self.arena.w32, name)

    def _g(self, name):      # fp32 gradient view
        return self.arena.view(self.arena.g32, name)

    def _drop(self, p):
        return float(p) if self.training else 0.0

    def _lanes(self, dev):
        cur = torch.cuda.current_stream(dev)
        if not self.overlap_streams:
            return _Lanes(cur, cur)
        if self._side_stream is None or self._side_stream.device != dev:
            self._side_stream = torch.cuda.Stream(device=dev)
        return _Lanes(cur, self._side_stream)

    # ------------------------------------------------------------------ building blocks (forward)
    def _linear(self, x, W, bias, M, epilogue=L.EPI_BIAS, aux=None, D2=None, p=0.0, seed=0, n=None):
        N, K = W.shape
        D = torch.empty(M, N, dtype=self.act, device=x.device)
        L.gemm(x, W, D, M=M, N=N, K=K, bias=bias, epilogue=epilogue, aux=aux, D2=D2, dropout_p=p, seed=seed, rows_dev=n)
        return D

    def _linear_res(self, x, W, bias, M, aux, p, seed, n=None, src=None):
        """z = dropout(x W^T + b) + aux, the pre-LayerNorm sum: kept in fp32 (CRCT_EPI_BIAS_RES_F32) — rounding z to bf16
        before the LayerNorm was the largest single contributor to the end-to-end error (tools/parity_sensitivity.py)."""
        N, K = W.shape
        z = torch.empty(M, N, dtype=torch.float32, device=x.device)
        L.gemm(x, W, z, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES if self.fp32 else L.EPI_BIAS_RES_F32,
               aux=aux if self.fp32 else aux.res, dropout_p=p, seed=seed, rows_dev=n, drop_rows=src)
        aux.res = None                     # consumed (same stream): the fp32 copy is not kept for the backward
        return z

    def _ln(self, z, pre, keep, n=None):
        rows, H = z.shape
        y = torch.empty(rows, H, dtype=self.act, device=z.device)
        # the residual stream stays fp32: `y` (bf16) is the next GEMM's operand, `y.res` the copy the next residual add reads
        y.res = None if self.fp32 else torch.empty(rows, H, dtype=torch.float32, device=z.device)
        mean = rstd = None
        if keep:
            mean = torch.empty(rows, dtype=torch.float32, device=z.device)
            rstd = torch.empty(rows, dtype=torch.float32, device=z.device)
        L.layernorm_fwd(z, self._p(pre + '.weight'), self._p(pre + '.bias'), y, mean, rstd, rows_dev=n, y32=y.res)
        return y, mean, rstd

    def _ffn_fwd(self, a, pre_i, pre_o, p_drop, seed, keep, rw=None):
        """intermediate + output blocks (vilbert.py:454-457,467-471 / 585-588,598-602)."""
        M = a.shape[0]
        W1, W2 = self._w(pre_i + '.dense.weight'), self._w(pre_o + '.dense.weight')
        dg = torch.empty(M, W1.shape[0], dtype=self.act, device=a.device) if keep else None      # gelu'(u), for the backward
        n, src = (rw.n, rw.src) if rw is not None else (None, None)
        h = self._linear(a, W1, self._p(pre_i + '.dense.bias'), M, L.EPI_BIAS_GELU, D2=dg, n=n)
        z = self._linear_res(h, W2, self._p(pre_o + '.dense.bias'), M, a, p_drop, seed, n, src)
        y, mean, rstd = self._ln(z, pre_o + '.LayerNorm', keep, n)
        self._tail[id(rw)] = (z, pre_o + '.LayerNorm')           # the stream's latest pre-LayerNorm sum (fp32): what the heads read
        # (keyed by the stream's row layout, not by width: hidden_size may equal v_hidden_size)
        s = None
        if keep:
            s = _Saved()
            s.a, s.dg, s.h, s.z, s.mean, s.rstd, s.p, s.seed, s.n, s.src = a, dg, h, z, mean, rstd, p_drop, seed, n, src
        return y, s

    def _ffn_bwd(self, dy, s, pre_i, pre_o):
        M, H = dy.shape
        n = s.n
        dz = torch.empty_like(dy)
        dzm = torch.empty_like(dy) if s.p > 0 else None
        self._ln_bwd(dy, s.z, s.mean, s.rstd, pre_o + '.LayerNorm', dz, self._g(pre_o + '.dense.bias'), dzm, p_out=s.p, seed_out=s.seed, n=n,
                     src=s.src)
        gz = dzm if dzm is not None else dz
        W1, W2 = self._w(pre_i + '.dense.weight'), self._w(pre_o + '.dense.weight')
        I = W1.shape[0]
        self._wgrad(gz, s.h, self._g(pre_o + '.dense.weight'), n=n)
        du = torch.empty(M, I, dtype=self.act, device=dy.device)
        L.gemm(gz, W2, du, M=M, N=I, K=H, b_major=1, epilogue=L.EPI_MUL, aux=s.dg, rows_dev=n)
        self._wgrad(du, s.a, self._g(pre_i + '.dense.weight'), self._g(pre_i + '.dense.bias'), n=n)
        da = torch.empty_like(dy)
        L.gemm(du, W1, da, M=M, N=H, K=I, b_major=1, epilogue=L.EPI_BIAS_RES, aux=dz, rows_dev=n)
        return da

    def _wg(self, *tensors):
        """Context: what runs inside is a weight gradient — nothing in the backward chain waits for it — so it goes to a
        second stream, ordered after everything enqueued so far on the current one: its CTAs take the SMs that the
        chain's kernels (partial last waves, small grids) leave idle.  `tensors` are its operands: they stay referenced
        until `_wgrad_join` (called before a gradient range is reported finished) has made the chain wait for it."""
        if not self.overlap_streams:
            return _NullCtx()
        dev = tensors[0].device
        cur = torch.cuda.current_stream(dev)
        if self._wg_stream is None or self._wg_stream.device != dev:
            self._wg_stream = torch.cuda.Stream(device=dev)
        self._wg_stream.wait_stream(cur)
        self._wg_hold.append((tensors, cur))
        return torch.cuda.stream(self._wg_stream)

    def _wgrad(self, dy, x, gW, gb=None, n=None):
        """gW[out,in] += dy^T x  (fp32 accumulate, split-K) and, when `gb` is given, gb += colsum(dy); `n`: device row count."""
        rows, No = dy.shape
        Ki = x.shape[1]
        kw = dict(M=No, N=Ki, K=rows, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, lda=dy.stride(0), ldb=x.stride(0), ldd=Ki,
                  rows_dev=n)
        if self.group_wgrads and not self.fp32:
            # deferred: a layer's weight gradients go out together as ONE grouped launch (crct_gemm_wgrad_grouped)
            if gb is not None:
                with self._wg(dy):
                    L.colsum_bf16(dy, gb, rows_dev=n)
            self._wg_pending.append((L.gemm_args(dy, x, gW, **kw), (dy, x), torch.cuda.current_stream(dy.device)))
            if len(self._wg_pending) >= 8:
                self._wgrad_flush()
            return
        with self._wg(dy, x):
            if gb is not None:
                L.colsum_bf16(dy, gb, rows_dev=n)
            L.gemm(dy, x, gW, **kw)

    def _wgrad_flush(self):
        """Issue the deferred weight gradients as one grouped launch on the weight-gradient stream, ordered after everything their
        producer streams have enqueued so far."""
        if not self._wg_pending:
            return
        pend, self._wg_pending = self._wg_pending, []
        args = [a for a, _, _ in pend]
        if not self.overlap_streams:
            L.gemm_wgrad_grouped(args)
            return
        dev = pend[0][1][0].device
        if self._wg_stream is None or self._wg_stream.device != dev:
            self._wg_stream = torch.cuda.Stream(device=dev)
        for cur in {id(c): c for _, _, c in pend}.values():
            self._wg_stream.wait_stream(cur)
        with torch.cuda.stream(self._wg_stream):
            L.gemm_wgrad_grouped(args)
        for _, tensors, cur in pend:
            self._wg_hold.append((tensors, cur))       # operands stay referenced until the chain has joined the stream

    def _wgrad_join(self):
        self._wgrad_flush()
        if self._wg_hold:
            for cur in {id(c): c for _, c in self._wg_hold}.values():
                cur.wait_stream(self._wg_stream)
            self._wg_hold.clear()

    def _ln_bwd(self, dy, z, mean, rstd, pre_ln, dz, g_dbias=None, dzm=None, p_in=0.0, seed_in=0, p_out=0.0, seed_out=0, n=None, src=None):
        """LayerNorm backward in its split form: dz (and the dropout-masked dzm) on the chain's stream, the three column
        sums (dgamma, dbeta, dense-bias gradient) with the weight gradients."""
        L.layernorm_bwd(dy, z, mean, rstd, self._p(pre_ln + '.weight'), dz, dzm=dzm, p_in=p_in, seed_in=seed_in, p_out=p_out,
                        seed_out=seed_out, rows_dev=n, drop_rows=src)
        with self._wg(dy, z, dz, dzm):
            L.layernorm_bwd_params(dy, z, mean, rstd, dz, self._g(pre_ln + '.weight'), self._g(pre_ln + '.bias'), dbias=g_dbias,
                                   dzm=dzm, p_in=p_in, seed_in=seed_in, p_out=p_out, rows_dev=n, drop_rows=src)

    def _attn_out_fwd(self, ctx, x, pre_dense, pre_ln, p_drop, seed, keep, rw=None):
        """dense + dropout + residual + LayerNorm (vilbert.py:424-428 / 555-559 / 749-756)."""
        M = x.shape[0]
        n, src = (rw.n, rw.src) if rw is not None else (None, None)
        z = self._linear_res(ctx, self._w(pre_dense + '.weight'), self._p(pre_dense + '.bias'), M, x, p_drop, seed, n, src)
        a, mean, rstd = self._ln(z, pre_ln, keep, n)
        s = None
        if keep:
            s = _Saved()
            s.ctx, s.z, s.mean, s.rstd, s.p, s.seed, s.n, s.src = ctx, z, mean, rstd, p_drop, seed, n, src
        return a, s

    def _attn_out_bwd(self, da, s, pre_dense, pre_ln):
        """returns (dz for the residual path, dctx)."""
        M, H = da.shape
        n = s.n
        dz = torch.empty_like(da)
        dzm = torch.empty_like(da) if s.p > 0 else None
        self._ln_bwd(da, s.z, s.mean, s.rstd, pre_ln, dz, self._g(pre_dense + '.bias'), dzm, p_out=s.p, seed_out=s.seed, n=n, src=s.src)
        gz = dzm if dzm is not None else dz
        W = self._w(pre_dense + '.weight')
        self._wgrad(gz, s.ctx, self._g(pre_dense + '.weight'), n=n)
        dctx = torch.empty(M, W.shape[1], dtype=self.act, device=da.device)
        L.gemm(gz, W, dctx, M=M, N=W.shape[1], K=H, b_major=1, rows_dev=n)
        return dz, dctx

    def _self_layer_fwd(self, x, rw, nh, pre, names, drops, layer, keep):
        """BertLayer / BertImageLayer (vilbert.py:474-485, 605-616).  `rw`: the stream's row layout (_Rows)."""
        M, H = x.shape
        B, Lseq, n = rw.B, rw.L, rw.n
        dh = H // nh
        step = self._step
        Wqkv = self.arena.fused(self._wflat(), [pre + '.attention.self.' + nm for nm in names], '.weight')
        bqkv = self.arena.fused(self.arena.w32, [pre + '.attention.self.' + nm for nm in names], '.bias')
        qkv = self._linear(x, Wqkv, bqkv, M, n=n)
        ctx = torch.empty(M, H, dtype=self.act, device=x.device)
        lse = torch.empty(B, nh, Lseq, dtype=torch.float32, device=x.device) if keep else None
        p_att, s_att = self._drop(drops[0]), _seed(step, 'attn', layer)
        L.attn_fwd(qkv, qkv[:, H:], qkv[:, 2 * H:], rw.mask_add, ctx, lse, B=B, nh=nh, dh=dh, Lq=Lseq, Lk=Lseq, ldq=3 * H, ldk=3 * H,
                   ldv=3 * H, ldo=H, dropout_p=p_att, seed=s_att, cu_q=rw.cu, cu_k=rw.cu)
        a, s_out = self._attn_out_fwd(ctx, x, pre + '.attention.output.dense', pre + '.attention.output.LayerNorm',
                                      self._drop(drops[1]), _seed(step, 'attn_out', layer), keep, rw)
        y, s_ffn = self._ffn_fwd(a, pre + '.intermediate', pre + '.output', self._drop(drops[1]), _seed(step, 'ffn_out', layer), keep, rw)
        s = None
        if keep:
            s = _Saved()
            s.x, s.qkv, s.lse, s.out, s.ffn, s.p_att, s.s_att = x, qkv, lse, s_out, s_ffn, p_att, s_att
            s.rw, s.nh = rw, nh
        return y, s

    def _self_layer_bwd(self, dy, s, pre, names):
        M, H = dy.shape
        dh = H // s.nh
        da = self._ffn_bwd(dy, s.ffn, pre + '.intermediate', pre + '.output')
        dz1, dctx = self._attn_out_bwd(da, s.out, pre + '.attention.output.dense', pre + '.attention.output.LayerNorm')
        dqkv = torch.empty_like(s.qkv)
        rw = s.rw
        L.attn_bwd(s.qkv, s.qkv[:, H:], s.qkv[:, 2 * H:], rw.mask_add, s.out.ctx, dctx, s.lse, dqkv, dqkv[:, H:], dqkv[:, 2 * H:],
                   B=rw.B, nh=s.nh, dh=dh, Lq=rw.L, Lk=rw.L, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H, lddo=H, lddq=3 * H, lddk=3 * H,
                   lddv=3 * H, dropout_p=s.p_att, seed=s.s_att, cu_q=rw.cu, cu_k=rw.cu)
        mods = [pre + '.attention.self.' + nm for nm in names]
        self._wgrad(dqkv, s.x, self.arena.fused(self.arena.g32, mods, '.weight'), self.arena.fused(self.arena.g32, mods, '.bias'), n=rw.n)
        Wqkv = self.arena.fused(self._wflat(), mods, '.weight')
        dx = torch.empty_like(dy)
        L.gemm(dqkv, Wqkv, dx, M=M, N=H, K=3 * H, b_major=1, epilogue=L.EPI_BIAS_RES, aux=dz1, rows_dev=rw.n)
        self._wgrad_flush()                # this layer's four weight gradients: one launch
        return dx

    def _co_layer_fwd(self, v, t, rv, rt, pre, layer, keep, lanes):
        """BertConnectionLayer (vilbert.py:774-788); stream 1 = visual, stream 2 = text.  Each lane projects its own
        q/k/v, the lanes meet once (each attention reads the other lane's keys/values), then run apart again."""
        cfg, step = self.cfg, self._step
        Mv, Hv = v.shape
        Mt, H = t.shape
        B, T, R = rt.B, rt.L, rv.L
        nh, Hb = cfg.bi_num_attention_heads, cfg.bi_hidden_size
        dh = Hb // nh
        ld = 3 * Hb
        m1 = [pre + '.biattention.' + nm for nm in ('query1', 'key1', 'value1')]
        m2 = [pre + '.biattention.' + nm for nm in ('query2', 'key2', 'value2')]
        with lanes.vis():
            qkv1 = self._linear(v, self.arena.fused(self._wflat(), m1, '.weight'), self.arena.fused(self.arena.w32, m1, '.bias'), Mv, n=rv.n)
        qkv2 = self._linear(t, self.arena.fused(self._wflat(), m2, '.weight'), self.arena.fused(self.arena.w32, m2, '.bias'), Mt, n=rt.n)
        lanes.meet()
        self._x_hold = (qkv1, qkv2)        # replaces the previous layer's pair: both lanes are past its readers now
        p1, s1 = self._drop(cfg.v_attention_probs_dropout_prob), _seed(step, 'co_attn1', layer)      # dropout1, vilbert.py:642,696
        p2, s2 = self._drop(cfg.attention_probs_dropout_prob), _seed(step, 'co_attn2', layer)        # dropout2, vilbert.py:649,718
        # biOutput is called with crossed arguments (vilbert.py:780): visual <- ctx2 via dense1/LayerNorm1, text <- ctx1 via dense2/LayerNorm2
        with lanes.vis():                  # visual queries over text keys/values
            ctx2 = torch.empty(Mv, Hb, dtype=self.act, device=t.device)
            lse2 = torch.empty(B, nh, R, dtype=torch.float32, device=t.device) if keep else None
            L.attn_fwd(qkv1, qkv2[:, Hb:], qkv2[:, 2 * Hb:], rt.mask_add, ctx2, lse2, B=B, nh=nh, dh=dh, Lq=R, Lk=T, ldq=ld, ldk=ld, ldv=ld,
                       ldo=Hb, dropout_p=p2, seed=s2, cu_q=rv.cu, cu_k=rt.cu)
            av, so_v = self._attn_out_fwd(ctx2, v, pre + '.biOutput.dense1', pre + '.biOutput.LayerNorm1',
                                          self._drop(cfg.v_hidden_dropout_prob), _seed(step, 'co_out_v', layer), keep, rv)
            yv, sf_v = self._ffn_fwd(av, pre + '.v_intermediate', pre + '.v_output', self._drop(cfg.v_hidden_dropout_prob),
                                     _seed(step, 'co_ffn_v', layer), keep, rv)
        ctx1 = torch.empty(Mt, Hb, dtype=self.act, device=t.device)          # text queries over visual keys/values
        lse1 = torch.empty(B, nh, T, dtype=torch.float32, device=t.device) if keep else None
        L.attn_fwd(qkv2, qkv1[:, Hb:], qkv1[:, 2 * Hb:], rv.mask_add, ctx1, lse1, B=B, nh=nh, dh=dh, Lq=T, Lk=R, ldq=ld, ldk=ld, ldv=ld,
                   ldo=Hb, dropout_p=p1, seed=s1, cu_q=rt.cu, cu_k=rv.cu)
        at, so_t = self._attn_out_fwd(ctx1, t, pre + '.biOutput.dense2', pre + '.biOutput.LayerNorm2',
                                      self._drop(cfg.hidden_dropout_prob), _seed(step, 'co_out_t', layer), keep, rt)
        yt, sf_t = self._ffn_fwd(at, pre + '.t_intermediate', pre + '.t_output', self._drop(cfg.hidden_dropout_prob),
                                 _seed(step, 'co_ffn_t', layer), keep, rt)
        s = None
        if keep:
            s = _Saved()
            s.v, s.t, s.qkv1, s.qkv2, s.lse1, s.lse2 = v, t, qkv1, qkv2, lse1, lse2
            s.so_v, s.so_t, s.sf_v, s.sf_t = so_v, so_t, sf_v, sf_t
            s.p1, s.s1, s.p2, s.s2, s.rv, s.rt, s.B, s.T, s.R = p1, s1, p2, s2, rv, rt, B, T, R
        return yv, yt, s

    def _co_layer_bwd(self, dyv, dyt, s, pre, lanes):
        """Mirror of the forward's dataflow: each lane runs its FFN / output-block backward, the lanes meet, each runs
        one attention direction (disjoint slices of dqkv1/dqkv2), they meet again, each lane finishes its projection."""
        cfg = self.cfg
        Mv, Hv = dyv.shape
        Mt, H = dyt.shape
        nh, Hb = cfg.bi_num_attention_heads, cfg.bi_hidden_size
        dh, ld = Hb // nh, 3 * Hb
        m1 = [pre + '.biattention.' + nm for nm in ('query1', 'key1', 'value1')]
        m2 = [pre + '.biattention.' + nm for nm in ('query2', 'key2', 'value2')]
        rv, rt = s.rv, s.rt
        with lanes.vis():
            dav = self._ffn_bwd(dyv, s.sf_v, pre + '.v_intermediate', pre + '.v_output')
            dzv, dctx2 = self._attn_out_bwd(dav, s.so_v, pre + '.biOutput.dense1', pre + '.biOutput.LayerNorm1')
            dqkv1 = torch.empty_like(s.qkv1)
        dat = self._ffn_bwd(dyt, s.sf_t, pre + '.t_intermediate', pre + '.t_output')
        dzt, dctx1 = self._attn_out_bwd(dat, s.so_t, pre + '.biOutput.dense2', pre + '.biOutput.LayerNorm2')
        dqkv2 = torch.empty_like(s.qkv2)
        lanes.meet()
        # direction 1: q = text (qkv2[:, :Hb]), k/v = visual  -> dq2, dk1, dv1
        L.attn_bwd(s.qkv2, s.qkv1[:, Hb:], s.qkv1[:, 2 * Hb:], rv.mask_add, s.so_t.ctx, dctx1, s.lse1, dqkv2, dqkv1[:, Hb:], dqkv1[:, 2 * Hb:],
                   B=s.B, nh=nh, dh=dh, Lq=s.T, Lk=s.R, ldq=ld, ldk=ld, ldv=ld, ldo=Hb, lddo=Hb, lddq=ld, lddk=ld, lddv=ld,
                   dropout_p=s.p1, seed=s.s1, cu_q=rt.cu, cu_k=rv.cu)
        with lanes.vis():
            # direction 2: q = visual (qkv1[:, :Hb]), k/v = text -> dq1, dk2, dv2
            L.attn_bwd(s.qkv1, s.qkv2[:, Hb:], s.qkv2[:, 2 * Hb:], rt.mask_add, s.so_v.ctx, dctx2, s.lse2, dqkv1, dqkv2[:, Hb:], dqkv2[:, 2 * Hb:],
                       B=s.B, nh=nh, dh=dh, Lq=s.R, Lk=s.T, ldq=ld, ldk=ld, ldv=ld, ldo=Hb, lddo=Hb, lddq=ld, lddk=ld, lddv=ld,
                       dropout_p=s.p2, seed=s.s2, cu_q=rv.cu, cu_k=rt.cu)
        lanes.meet()
        self._x_hold = (dqkv1, dqkv2, dctx1, dctx2)
        with lanes.vis():
            self._wgrad(dqkv1, s.v, self.arena.fused(self.arena.g32, m1, '.weight'), self.arena.fused(self.arena.g32, m1, '.bias'), n=rv.n)
            dv = torch.empty_like(dyv)
            L.gemm(dqkv1, self.arena.fused(self._wflat(), m1, '.weight'), dv, M=Mv, N=Hv, K=ld, b_major=1, epilogue=L.EPI_BIAS_RES, aux=dzv,
                   rows_dev=rv.n)
        self._wgrad(dqkv2, s.t, self.arena.fused(self.arena.g32, m2, '.weight'), self.arena.fused(self.arena.g32, m2, '.bias'), n=rt.n)
        dt = torch.empty_like(dyt)
        L.gemm(dqkv2, self.arena.fused(self._wflat(), m2, '.weight'), dt, M=Mt, N=H, K=ld, b_major=1, epilogue=L.EPI_BIAS_RES, aux=dzt,
               rows_dev=rt.n)
        self._wgrad_flush()                # the block's eight weight gradients (both streams): one launch
        return dv, dt

    # ------------------------------------------------------------------ heads (fp32, CUDA cores)
    # Every Linear of the heads is a tiny latency-bound problem (M = batch size); independent ones are issued as ONE
    # batched launch (crct_linear_f32_batched): the two regressor pipes and the poolers side by side in the forward,
    # {wgrad, bias-grad, dgrad} of a layer (of both pipes) in the backward.
    def _fwd_problem(self, x, name, act, out=None, ldc=None):
        W, b = self._p(name + '.weight'), self._p(name + '.bias')
        M, (N, K) = x.shape[0], W.shape
        if out is None:
            out = torch.empty(M, N, dtype=torch.float32, device=x.device)
            ldc = N
        return L.lin_problem(x, x.stride(0), 1, W, 1, K, out, ldc, M, N, K, bias=b, act=act), out

    def _bwd_problems(self, dy, ldy, x, name, dx=None, dmask=None, slope=0.0, accumulate_dx=0):
        """dy [M,N] (row stride ldy) = gradient of the PRE-activation output of Linear `name` applied to x [M,K].
        Returns ([wgrad, bias-grad, dgrad] problems, dx); dx = (dy W) * (dmask > 0 ? 1 : slope)."""
        W = self._p(name + '.weight')
        N, K = W.shape
        M = x.shape[0]
        probs = [L.lin_problem(dy, 1, ldy, x, x.stride(0), 1, self._g(name + '.weight'), K, N, K, M, accumulate=1),
                 L.lin_problem(None, 0, 0, dy, ldy, 1, self._g(name + '.bias'), N, 1, N, M, accumulate=1)]
        if dx is None:
            dx = torch.empty(M, K, dtype=torch.float32, device=x.device)
        probs.append(L.lin_problem(dy, ldy, 1, W, K, 1, dx, K, M, K, N, dmask=dmask, ldm=(dmask.stride(0) if dmask is not None else 0),
                                   slope=slope, accumulate=accumulate_dx))
        return probs, dx

    def _heads_fwd(self, t, v, rt, rv, labels, Rt, kind, keep):
        cfg, dev = self.cfg, t.device
        B, T, R = rt.B, rt.L, rv.L
        H, Hv, Hb = cfg.hidden_size, cfg.v_hidden_size, cfg.bi_hidden_size
        hw0 = torch.empty(B, H, dtype=torch.float32, device=dev)
        hv0 = torch.empty(B, Hv, dtype=torch.float32, device=dev)
        # first token / region of every sample (vilbert.py:958,973 / 1599-1600), row cu[b] when packed: LayerNorm of the last
        # pre-LayerNorm sum in fp32 — the heads never see the bf16 rounding of the sequence outputs t / v
        (z_t, ln_t), (z_v, ln_v) = self._tail[id(rt)], self._tail[id(rv)]
        L.layernorm_rows_f32(z_t, self._p(ln_t + '.weight'), self._p(ln_t + '.bias'), hw0, rt.cu, T)
        L.layernorm_rows_f32(z_v, self._p(ln_v + '.weight'), self._p(ln_v + '.bias'), hv0, rv.cu, R)
        prefusion = torch.empty(B, 512, dtype=torch.float32, device=dev)          # cat((hv, hw), -1), regressor.py:40
        p_t, pt = self._fwd_problem(hw0, 'bert.t_pooler.dense', L.ACT_RELU)
        p_v, pv = self._fwd_problem(hv0, 'bert.v_pooler.dense', L.ACT_RELU)
        acts_v, acts_t = [hv0], [hw0]
        probs = [p_t, p_v]
        for i, idx in enumerate((0, 2, 4, 6)):
            last = i == 3
            pr_v, ov = self._fwd_problem(acts_v[-1], f'regressor.vis_pipe.{idx}', L.ACT_NONE if last else L.ACT_LEAKY,
                                         out=prefusion if last else None, ldc=512 if last else None)
            pr_t, ot = self._fwd_problem(acts_t[-1], f'regressor.txt_pipe.{idx}', L.ACT_NONE if last else L.ACT_LEAKY,
                                         out=prefusion[:, 256:] if last else None, ldc=512 if last else None)
            L.linear_f32_batched(probs + [pr_v, pr_t])
            probs = []
            acts_v.append(ov)
            acts_t.append(ot)
        pooled = torch.empty_like(pt)
        p_cls, s_cls = self._drop(0.1), _seed(self._step, 'cls')      # nn.Dropout(0.1), vilbert.py:1045
        L.pool_mul_fwd(pt, pv, pooled, p_cls, s_cls)
        pr_c, logits = self._fwd_problem(pooled, 'cls.bi_seq_relationship', L.ACT_NONE)
        acts_f = [prefusion]
        probs = [pr_c]
        for i, idx in enumerate((0, 2, 4, 6)):
            pr_f, of = self._fwd_problem(acts_f[-1], f'regressor.fusion.{idx}', L.ACT_TANH if i == 3 else L.ACT_LEAKY)
            L.linear_f32_batched(probs + [pr_f])
            probs = []
            acts_f.append(of)
        reg = acts_f[-1]
        outs = [torch.empty(B, dtype=torch.float32, device=dev) for _ in range(4)]
        scalars = torch.empty(5, dtype=torch.float32, device=dev)
        dlogits = torch.empty(B, 2, dtype=torch.float32, device=dev) if keep else None
        dpre = torch.empty(B, 1, dtype=torch.float32, device=dev) if keep else None
        L.hybrid_loss(logits, reg, labels, Rt, *outs, scalars, dlogits, dpre, l1=bool(self.params['L1']),
                      zero_impossible=(kind != 'L1'), tol_margin=float(self.params['tol_margin']),
                      nsp_coeff=float(self.params.get('nsp_loss_coeff', 1.0)), reg_coeff=float(self.params.get('reg_loss_coeff', 1.0)),
                      unit_grads=1)
        s = None
        if keep:
            s = _Saved()
            s.hw0, s.hv0, s.pt, s.pv, s.pooled, s.p_cls, s.s_cls = hw0, hv0, pt, pv, pooled, p_cls, s_cls
            s.acts_v, s.acts_t, s.acts_f, s.prefusion, s.dlogits, s.dpre = acts_v, acts_t, acts_f, prefusion, dlogits, dpre
        return logits, outs, scalars, s

    def _heads_bwd(self, s, d_nsp, d_reg, rt, rv):
        cfg, dev = self.cfg, s.hw0.device
        B, T, R = rt.B, rt.L, rv.L
        H, Hv = cfg.hidden_size, cfg.v_hidden_size
        dlogits, dpre = torch.empty_like(s.dlogits), torch.empty_like(s.dpre)
        L.scale_rows(s.dlogits, d_nsp.reshape(-1)[:1].contiguous().float(), dlogits)
        L.scale_rows(s.dpre, d_reg.reshape(-1).contiguous().float(), dpre)
        # Per layer: the input gradient (dgrad) is the serial chain everything after it waits for; the weight and bias gradients
        # are not — they go to the weight-gradient stream like the transformer's (`_wg`), so the chain's launches stay small
        # (16.95 -> 16.84 ms per step, profiles/r01_ab_heads_wgrad_s19.txt).
        def run(*layers):          # layers: (problems [wgrad, bias-grad, dgrad], dy, x)
            L.linear_f32_batched([pr[2] for pr, _, _ in layers])
            with self._wg(*[t for _, dy_, x_ in layers for t in (dy_, x_)]):
                L.linear_f32_batched([q for pr, _, _ in layers for q in pr[:2]])

        # classifier + last fusion layer
        pc, dpooled = self._bwd_problems(dlogits, 2, s.pooled, 'cls.bi_seq_relationship')
        pf, d = self._bwd_problems(dpre, 1, s.acts_f[3], 'regressor.fusion.6', dmask=s.acts_f[3], slope=0.01)
        run((pc, dlogits, s.pooled), (pf, dpre, s.acts_f[3]))
        dut, duv = torch.empty_like(s.pt), torch.empty_like(s.pv)
        L.pool_mul_bwd(dpooled, s.pt, s.pv, dut, duv, s.p_cls, s.s_cls)
        # poolers + fusion.4
        pt_, dhw0 = self._bwd_problems(dut, dut.stride(0), s.hw0, 'bert.t_pooler.dense')
        pv_, dhv0 = self._bwd_problems(duv, duv.stride(0), s.hv0, 'bert.v_pooler.dense')
        d_in = d
        pf, d = self._bwd_problems(d_in, d_in.stride(0), s.acts_f[2], 'regressor.fusion.4', dmask=s.acts_f[2], slope=0.01)
        run((pt_, dut, s.hw0), (pv_, duv, s.hv0), (pf, d_in, s.acts_f[2]))
        d_in = d
        pf, d = self._bwd_problems(d_in, d_in.stride(0), s.acts_f[1], 'regressor.fusion.2', dmask=s.acts_f[1], slope=0.01)
        run((pf, d_in, s.acts_f[1]))
        d_in = d
        pf, dpref = self._bwd_problems(d_in, d_in.stride(0), s.acts_f[0], 'regressor.fusion.0')      # input = cat(pipe outputs): no activation
        run((pf, d_in, s.acts_f[0]))
        # the two pipes, layer by layer, side by side
        dv_, dt_ = dpref, dpref[:, 256:]
        ldv = ldt = 512
        for i, idx in reversed(list(enumerate((0, 2, 4, 6)))):
            xv, xt = s.acts_v[i], s.acts_t[i]
            if i > 0:      # x is the LeakyReLU output of the previous Linear: fold its derivative into dx
                pv_, nv = self._bwd_problems(dv_, ldv, xv, f'regressor.vis_pipe.{idx}', dmask=xv, slope=0.01)
                pt_, nt = self._bwd_problems(dt_, ldt, xt, f'regressor.txt_pipe.{idx}', dmask=xt, slope=0.01)
            else:          # first layer: its input gradient joins the pooler's (first-token hidden state)
                pv_, nv = self._bwd_problems(dv_, ldv, xv, f'regressor.vis_pipe.{idx}', dx=dhv0, accumulate_dx=1)
                pt_, nt = self._bwd_problems(dt_, ldt, xt, f'regressor.txt_pipe.{idx}', dx=dhw0, accumulate_dx=1)
            run((pv_, dv_, xv), (pt_, dt_, xt))
            dv_, dt_, ldv, ldt = nv, nt, nv.stride(0), nt.stride(0)
        dt = torch.empty(B * T, H, dtype=self.act, device=dev)
        dv = torch.empty(B * R, Hv, dtype=self.act, device=dev)
        L.fill_zero(dt)
        L.fill_zero(dv)
        if rt.cu is not None:
            L.scatter_rows_f32(dhw0, dt, rt.cu)
            L.scatter_rows_f32(dhv0, dv, rv.cu)
        else:
            L.scatter_first(dhw0, dt, T * H)
            L.scatter_first(dhv0, dv, R * Hv)
        return dt, dv

    # ------------------------------------------------------------------ whole model
    def _run_forward(self, ids, types, loc, feat, box, cls, amask, imask, labels, Rt, kind, keep, group=None):
        cfg, arena = self.cfg, self.arena
        L.device_check()
        self._bind_grads()
        arena.refresh_bf16()
        dev = ids.device
        B, T = ids.shape
        Bv, R, F = feat.shape          # Bv = B, or the number of questions when `group` maps candidates to questions (f3)
        H, Hv = cfg.hidden_size, cfg.v_hidden_size
        if group is None and Bv != B:
            raise ValueError(f'{B} text rows but {Bv} visual rows (pass image_group to share visual inputs between candidates)')
        if group is not None and (keep or group.shape != (B,)):
            raise ValueError('image_group [B] is an inference-only input (one question index per candidate sequence)')
        if F != cfg.v_feature_size:
            raise ValueError(f'image_feat has {F} features, config says {cfg.v_feature_size}')
        if box.shape[-1] != 4:
            raise ValueError('image_loc must be [B,R,4] (CRCT/fig_dataloader.py:346 strips the 5th column)')
        step = self._step
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device=dev)
        if self.varlen:
            # packed rows (csrc/varlen.cu): compact the valid tokens / regions once; no additive masks from here on
            ft, fv = self.row_fill_hint if self.row_fill_hint else (None, None)
            rt = _Rows(B, T, i32(B + 1), i32(B * T), fill=ft)
            L.row_map(amask, rt.cu, rt.src)
            rvq = _Rows(Bv, R, i32(Bv + 1), i32(Bv * R), fill=fv)  # per question (f3) == per sequence without `group`
            L.row_map(imask, rvq.cu, rvq.src)
            rv = rvq
            if group is not None:          # candidate n shares the packed region rows of question group[n]
                rv = _Rows(B, R, i32(B + 1), i32(B * R), fill=fv)
                L.group_map(rvq.cu, group, rv.cu, rv.src)
        else:
            t_mask = torch.empty(B, T, dtype=torch.float32, device=dev)
            v_mask = torch.empty(Bv, R, dtype=torch.float32, device=dev)
            L.additive_mask(amask, t_mask)
            L.additive_mask(imask, v_mask)
            if group is not None:          # per-candidate copy of the per-question mask (read by both lanes: allocated here)
                m_q, v_mask = v_mask, torch.empty(B, R, dtype=torch.float32, device=dev)
                L.expand_blocks(m_q, group, v_mask)
            rt, rvq = _Rows(B, T, mask_add=t_mask), _Rows(Bv, R)
            rv = _Rows(B, R, mask_add=v_mask)
        sv = _Saved() if keep else None
        # --- embeddings (vilbert.py:1412-1413); from here to the heads the visual lane runs on its own stream
        lanes = self._lanes(dev)
        lanes.v_wait_t()
        e = 'bert.embeddings'
        t = torch.empty(B * T, H, dtype=self.act, device=dev)
        t.res = None if self.fp32 else torch.empty(B * T, H, dtype=torch.float32, device=dev)
        zt = torch.empty(B * T, H, dtype=torch.float32, device=dev) if keep else None      # pre-LayerNorm sums stay fp32
        mt = torch.empty(B * T, dtype=torch.float32, device=dev) if keep else None
        rt_ = torch.empty(B * T, dtype=torch.float32, device=dev) if keep else None
        p_et, s_et = self._drop(cfg.hidden_dropout_prob), _seed(step, 'emb_t')
        L.embed_text_fwd(ids, types, loc, self._p(e + '.word_embeddings.weight'), self._p(e + '.position_embeddings.weight'),
                         self._p(e + '.plotqa_type_embeddings.weight'), self._p(e + '.txt_location_embeddings.weight'),
                         self._p(e + '.txt_location_embeddings.bias'), self._p(e + '.LayerNorm.weight'), self._p(e + '.LayerNorm.bias'),
                         t, zt, mt, rt_, dropout_p=p_et, seed=s_et, src_row=rt.src, rows_dev=rt.n, y32=t.res)
        e = 'bert.v_embeddings'
        feat2 = feat.reshape(Bv * R, F)
        box2, cls2 = box.reshape(Bv * R, 4), cls.reshape(Bv * R)
        p_ev, s_ev = self._drop(cfg.hidden_dropout_prob), _seed(step, 'emb_v')      # nn.Dropout(config.hidden_dropout_prob), vilbert.py:1470
        with lanes.vis():
            probs = torch.empty(Bv * R, F, dtype=self.act, device=dev)
            L.softmax_rows(feat2, probs, src_row=rvq.src, rows_dev=rvq.n)
            gimg = self._linear(probs, self._w(e + '.new_image_embeddings.weight'), self._p(e + '.new_image_embeddings.bias'), Bv * R,
                                n=rvq.n)
            v = torch.empty(Bv * R, Hv, dtype=self.act, device=dev)
            v.res = None if self.fp32 else torch.empty(Bv * R, Hv, dtype=torch.float32, device=dev)
            zv = torch.empty(Bv * R, Hv, dtype=torch.float32, device=dev) if keep else None
            mv = torch.empty(Bv * R, dtype=torch.float32, device=dev) if keep else None
            rv_ = torch.empty(Bv * R, dtype=torch.float32, device=dev) if keep else None
            L.embed_vis_fwd(gimg, box2, cls2, self._p(e + '.new_loc_emb.weight'), self._p(e + '.new_loc_emb.bias'),
                            self._p(e + '.color_emb.weight'), self._p(e + '.LayerNorm.weight'), self._p(e + '.LayerNorm.bias'),
                            v, zv, mv, rv_, dropout_p=p_ev, seed=s_ev, src_row=rvq.src, rows_dev=rvq.n, y32=v.res)
            if group is not None:
                # f3: the visual embedding depends on the image only — computed once per question above, fanned out to the
                # question's candidate sequences here (the reference replicates the fp32 inputs on the host instead,
                # fig_dataloader.py:690-693); from the first co-attention on the visual stream is per candidate
                v_q = v
                v = torch.empty(B * R, Hv, dtype=self.act, device=dev)
                v.res = None if self.fp32 else torch.empty(B * R, Hv, dtype=torch.float32, device=dev)
                for src_, dst_ in ((v_q, v), (v_q.res, v.res)):
                    if src_ is None:
                        continue
                    if self.varlen:
                        L.gather_rows(src_, rv.src, dst_, rows_dev=rv.n)
                    else:
                        L.expand_blocks(src_.view(Bv, R * Hv), group, dst_.view(B, R * Hv))
        # --- encoder (vilbert.py:852-939): text layers on the text lane, visual layers on the visual lane (v_layer[k-1]
        # and layer[5+k] are independent, vilbert.py:868-886; the 3520-row visual kernels fill the SMs the text kernels'
        # partial waves leave idle); the lanes meet inside every connection layer.
        layers = []
        for kind_, i in cfg.schedule():
            if kind_ == 't':
                t, s = self._self_layer_fwd(t, rt, cfg.num_attention_heads, f'bert.encoder.layer.{i}', ('query', 'key', 'value'),
                                            (cfg.attention_probs_dropout_prob, cfg.hidden_dropout_prob), i, keep)
            elif kind_ == 'v':
                with lanes.vis():
                    v, s = self._self_layer_fwd(v, rv, cfg.v_num_attention_heads, f'bert.encoder.v_layer.{i}',
                                                ('query', 'key', 'value'),
                                                (cfg.v_attention_probs_dropout_prob, cfg.v_hidden_dropout_prob), 100 + i, keep)
            else:
                v, t, s = self._co_layer_fwd(v, t, rv, rt, f'bert.encoder.c_layer.{i}', 200 + i, keep, lanes)
            layers.append(s)
        lanes.t_wait_v()
        self._x_hold = None
        logits, outs, scalars, s_heads = self._heads_fwd(t, v, rt, rv, labels, Rt, kind, keep)
        self._tail = {}
        if keep:
            sv.B, sv.T, sv.R, sv.rows_t, sv.rows_v = B, T, R, rt, rv
            sv.ids, sv.types, sv.loc, sv.box2, sv.cls2, sv.probs = ids, types, loc, box2, cls2, probs
            sv.zt, sv.mt, sv.rt, sv.p_et, sv.s_et = zt, mt, rt_, p_et, s_et
            sv.zv, sv.mv, sv.rv, sv.p_ev, sv.s_ev = zv, mv, rv_, p_ev, s_ev
            sv.layers, sv.heads = layers, s_heads
        return logits, outs, scalars, t, sv

    def _backward(self, sv, d_nsp, d_reg):
        """Hand-written backward; reports every finished gradient range to `grad_ready_hook` (data-parallel exchange)."""
        hook = self.grad_ready_hook
        for lo, hi in self._backward_stages(sv, d_nsp, d_reg):
            if hook:
                hook(lo, hi)
        if hook:
            hook(None, None)         # backward finished

    def _backward_stages(self, sv, d_nsp, d_reg):
        """Generator form of the backward: yields the [lo, hi) arena range whose gradients have just become final,
        from the tail of the arena (heads) to its head (embeddings).  `graph.GraphedTrainStep` drives it directly to
        cut the captured step into segments between which the bucketed all-reduces are launched."""
        cfg, arena = self.cfg, self.arena
        L.SALT = sv.salt                   # the word THIS pass's forward drew its masks with (not the live counter)
        B, T, R = sv.B, sv.T, sv.R
        rows_t, rows_v = sv.rows_t, sv.rows_v
        dt, dv = self._heads_bwd(sv.heads, d_nsp, d_reg, rows_t, rows_v)
        dev = dt.device
        # Finished ranges are only reported one by one when somebody consumes them (the data-parallel hook, or a graph
        # cut into per-bucket segments): reporting a range makes the caller's stream wait for both other streams.
        fine = self.grad_ready_hook is not None or self.segment_ranges
        ahook = None if fine else self.async_range_hook
        lanes = self._lanes(dev)
        if fine:
            self._wgrad_join()             # the heads' weight gradients run on the weight-gradient stream
            yield arena.offsets['bert.t_pooler.dense.weight'], arena.live_end
        elif ahook:
            ahook(arena.offsets['bert.t_pooler.dense.weight'], arena.live_end, [torch.cuda.current_stream(dev)] +
                  ([self._wg_stream] if (self.overlap_streams and self._wg_stream is not None) else []))
        lanes.v_wait_t()
        dv_heads = dv                      # allocated on the text lane, read on the visual lane: keep it until the end
        pending = []                       # finished layers whose range has not been reported yet

        def report():
            self._wgrad_join()
            lanes.t_wait_v()
            out = list(pending)
            pending.clear()
            return out

        def vis_embeddings_bwd(dv):
            e = 'bert.v_embeddings'
            with lanes.vis():
                dzv = torch.empty_like(dv)
                self._ln_bwd(dv, sv.zv, sv.mv, sv.rv, e + '.LayerNorm', dzv, self._g(e + '.new_image_embeddings.bias'), p_in=sv.p_ev,
                             seed_in=sv.s_ev, n=rows_v.n, src=rows_v.src)
                self._wgrad(dzv, sv.probs, self._g(e + '.new_image_embeddings.weight'), self._g(e + '.new_loc_emb.bias'), n=rows_v.n)
                L.embed_vis_bwd(dzv, sv.box2, sv.cls2, self._g(e + '.color_emb.weight'), self._g(e + '.new_loc_emb.weight'),
                                src_row=rows_v.src, rows_dev=rows_v.n)

        items = list(reversed(list(zip(cfg.schedule(), sv.layers))))
        last_c = max((k for k, ((kind_, _), _) in enumerate(items) if kind_ == 'c'), default=-1)
        for k, ((kind_, i), s) in enumerate(items):
            if kind_ == 't':
                pre = f'bert.encoder.layer.{i}'
                dt = self._self_layer_bwd(dt, s, pre, ('query', 'key', 'value'))
            elif kind_ == 'v':
                pre = f'bert.encoder.v_layer.{i}'
                with lanes.vis():
                    dv = self._self_layer_bwd(dv, s, pre, ('query', 'key', 'value'))
            else:
                pre = f'bert.encoder.c_layer.{i}'
                dv, dt = self._co_layer_bwd(dv, dt, s, pre, lanes)
            pending.append(self._block_range(pre))
            if k == last_c:
                vis_embeddings_bwd(dv)     # nothing visual is left but the embeddings: next to the remaining text layers
            if fine and (kind_ == 'c' or k > last_c) and sum(hi - lo for lo, hi in pending) >= self.report_min_elems:
                for r in report():         # every report joins the three streams: one per exchange bucket, not per block
                    yield r
                lanes.v_wait_t()           # a graph cut may follow a report: the visual lane re-forks from the text lane
            elif ahook and (kind_ == 'c' or k > last_c):
                streams = [torch.cuda.current_stream(dev)] + ([self._wg_stream] if (self.overlap_streams and self._wg_stream is not None) else []) + \
                    ([lanes.v] if lanes.split else [])
                ahook(min(r[0] for r in pending), max(r[1] for r in pending), streams)
                pending.clear()
        if last_c < 0:
            vis_embeddings_bwd(dv)
        e = 'bert.embeddings'
        dzt = torch.empty_like(dt)
        self._ln_bwd(dt, sv.zt, sv.mt, sv.rt, e + '.LayerNorm', dzt, p_in=sv.p_et, seed_in=sv.s_et, n=rows_t.n, src=rows_t.src)
        L.embed_text_bwd(sv.ids, sv.types, sv.loc, dzt, self._g(e + '.word_embeddings.weight'), self._g(e + '.position_embeddings.weight'),
                         self._g(e + '.plotqa_type_embeddings.weight'), self._g(e + '.txt_location_embeddings.weight'),
                         self._g(e + '.txt_location_embeddings.bias'), src_row=rows_t.src, rows_dev=rows_t.n)
        rest = report()
        self._x_hold = None
        del dv_heads
        if fine:
            for r in rest:
                yield r
            yield 0, self._block_range('bert.v_embeddings')[1]
        else:
            if ahook:                      # what is left: both embedding blocks (everything is joined into this stream by now)
                ahook(0, self._block_range('bert.v_embeddings')[1], [torch.cuda.current_stream(dev)])
            yield 0, arena.live_end

    def train_step_stages(self, batch, nsp_coeff: float = 1.0, reg_coeff: float = 1.0):
        """Forward + backward of one training batch WITHOUT autograd (same kernels, same order as
        `glue_forward(...)[0].backward()`), as a generator over finished gradient ranges.  After the first `next()`
        the forward and the heads' backward have been enqueued; `self.last_scalars` = {loss, nsp, mean reg, counts},
        `self.last_logits`, `self.last_reg` hold the (device) outputs."""
        dev = self.arena.w32.device
        ids = batch['tokens']
        B, T = ids.shape
        seq_len = torch.gather(batch['sep_indices'], 1, batch['hist_len'].view(-1, 1)).squeeze(1) + 1       # encoder_decorator.py:118-119
        amask = torch.arange(T, device=dev).unsqueeze(0) < seq_len.unsqueeze(1)
        salt = self._pass_salt(dev)
        logits, outs, scalars, _, sv = self._run_forward(
            ids, batch['segments'], batch['loc'], batch['image_feat'], batch['image_loc'], batch['image_target'], amask,
            batch['image_mask'], batch['next_sentence_labels'].view(-1), batch['R'], 'L1_smooth', True)
        sv.salt = salt
        self.last_scalars, self.last_logits, self.last_reg = scalars, logits, outs
        d_nsp = torch.full((1,), float(nsp_coeff), dtype=torch.float32, device=dev)
        d_reg = torch.full((B,), float(reg_coeff) / B, dtype=torch.float32, device=dev)      # d mean_B(reg_loss) / d reg_loss[b]
        yield from self._backward_stages(sv, d_nsp, d_reg)

    def _block_range(self, prefix):
        """[lo, hi) element range of the live arena tensors under `prefix` (contiguous by construction)."""
        names = [p.name for p in self.arena.order if p.live and p.name.startswith(prefix + '.')]
        lo = self.arena.offsets[names[0]]
        last = self.arena.by_name[names[-1]]
        hi = self.arena.offsets[last.name] + (last.numel + 63) // 64 * 64
        return lo, hi

    def forward(self, input_ids, txt_loc, image_feat, image_loc, sep_indices=None, sep_len=None, token_type_ids=None,
                attention_mask=None, masked_lm_labels=None, next_sentence_label=None, head_mask=None, random_round_indices=None,
                output_nsp_scores=False, output_lm_scores=False, image_attention_mask=None, image_label=None, image_target=None,
                gt_reg=None, areas=None, legend_pred=None, image_group=None):
        """Same signature and return tuple as encoder_decorator.py:19-54, plus one optional inference-only argument:
        `image_group [B] int64` — when given, `image_feat / image_loc / image_attention_mask / image_target` hold one row per
        QUESTION and `image_group[b]` names the row candidate sequence b belongs to (f3, `cqa_crct_b200.evaluate`).  `sep_indices`, `sep_len`, `masked_lm_labels`
        (presence only), `image_label`, `head_mask`, `random_round_indices`, `legend_pred` take no part in the arithmetic
        (SURVEY.md §8b); `areas` is FigureQA/DVQA-only."""
        dev = self.arena.w32.device
        if dev.type != 'cuda':
            raise L.CrctError('cqa_crct_b200 has no CPU path: move the model to a B200 with .to("cuda")')
        if areas is not None:
            raise NotImplementedError('`areas` is only used by the FigureQA / DVQA configurations')
        if gt_reg is None:
            raise ValueError('gt_reg=[R, kind] is required (vilbert.py:1586)')
        train_branch = next_sentence_label is not None and masked_lm_labels is not None and image_target is not None   # :31-32
        cvt = lambda x, dt=None: None if x is None else x.to(device=dev, dtype=dt, non_blocking=True).contiguous()
        ids = cvt(input_ids, torch.int64)
        B, T = ids.shape
        types = cvt(token_type_ids, torch.int64) if token_type_ids is not None else torch.zeros_like(ids)           # vilbert.py:1368-1369
        loc = cvt(txt_loc, torch.float32)
        feat = cvt(image_feat, torch.float32)
        box = cvt(image_loc, torch.float32)
        R = feat.shape[1]
        if image_target is None:
            raise ValueError('image_target (RoI class ids for color_emb, vilbert.py:1479) is required')
        cls = cvt(image_target, torch.int64)
        amask = cvt(attention_mask) if attention_mask is not None else torch.ones(B, T, dtype=torch.int64, device=dev)   # :1366-1367
        imask = cvt(image_attention_mask) if image_attention_mask is not None else torch.ones(feat.shape[0], R, dtype=torch.int64, device=dev)
        if amask.dtype not in (torch.bool, torch.uint8, torch.int64, torch.float32):
            amask = amask.to(torch.int64)
        if imask.dtype not in (torch.bool, torch.uint8, torch.int64, torch.float32):
            imask = imask.to(torch.int64)
        Rt, kind = gt_reg[0], gt_reg[1]
        Rt = cvt(Rt, torch.float32)
        labels = cvt(next_sentence_label, torch.int64).view(-1) if train_branch else None
        keep = train_branch and torch.is_grad_enabled()
        salt = self._pass_salt(dev)         # fresh dropout masks for this pass (also on CUDA-graph replay)
        group = cvt(image_group, torch.int64)
        if group is not None and train_branch:
            raise ValueError('image_group is an inference-only input')
        logits, outs, scalars, seq_t, sv = self._run_forward(ids, types, loc, feat, box, cls, amask, imask, labels, Rt, kind, keep, group)
        reg_pred, reg_loss, reg_l1, reg_dist = outs
        nsp_loss = scalars[1:2]
        if keep:
            sv.nsp_loss, sv.reg_loss, sv.salt = nsp_loss, reg_loss, salt
            anchor = self._params_by_name['bert.embeddings.word_embeddings.weight']
            nsp_loss, reg_loss = _CrctFunction.apply(anchor, self, sv)
        reg = [reg_pred, reg_loss, reg_l1, (scalars[3], scalars[4]), reg_dist]          # vilbert.py:1590-1648 (counts stay on device)
        legend_loss = torch.zeros(1, dtype=torch.float32, device=dev)                   # vilbert.py:1583
        if train_branch:
            zero = torch.zeros(1, 1, dtype=torch.float32, device=dev)                   # vilbert.py:1652-1653
            out = (zero, zero.clone(), nsp_loss)
        else:
            out = (None, None, None)
        out = out + (logits,)
        if output_lm_scores:
            out = out + (None,)                                                          # prediction_scores_t is None, vilbert.py:1059
        return out + (reg, legend_loss)


def glue_forward(dialog_encoder, batch, params, output_nsp_scores=False, output_lm_scores=False, evaluation=False, sample_ids=None):
    """Same contract as the module-level `forward` of CRCT/backbone/encoder_decorator.py:73-158 (the function
    train.py:173 / evaluation.py:247 call): slices the batch, builds the text mask from sep_indices/hist_len, calls the
    model, combines the losses.  Index / mask bookkeeping on [B,T] integers stays in torch (host glue, not arithmetic)."""
    dev = params['device']
    idx = slice(None) if sample_ids is None else sample_ids
    take = lambda k: batch[k][idx]
    tokens, txt_loc, segments = take('tokens'), take('loc'), take('segments')
    sep_indices, mask, hist_len = take('sep_indices'), take('mask'), take('hist_len')
    regression_target = take('R')
    if not evaluation:
        next_sentence_labels = take('next_sentence_labels')
        image_label = take('image_label')
        regression_target = [regression_target, 'L1_smooth']                 # encoder_decorator.py:104
    else:
        next_sentence_labels, image_label = None, None
        regression_target = [regression_target, 'L1']                        # encoder_decorator.py:106
    seq_len = torch.gather(sep_indices, 1, hist_len.view(-1, 1)).squeeze(1) + 1                    # :118-119
    attention_mask = torch.arange(tokens.shape[1], device=seq_len.device).unsqueeze(0) < seq_len.unsqueeze(1)   # sequence_mask :57-70
    lm_loss, img_loss, nsp_loss, nsp_scores, regression, legend_loss = dialog_encoder(
        tokens, txt_loc, take('image_feat'), take('image_loc'), sep_indices=sep_indices, sep_len=hist_len + 1,
        token_type_ids=segments, masked_lm_labels=mask, attention_mask=attention_mask, next_sentence_label=next_sentence_labels,
        output_nsp_scores=output_nsp_scores, output_lm_scores=output_lm_scores, image_attention_mask=take('image_mask'),
        image_label=image_label, image_target=take('image_target'), gt_reg=regression_target,
        areas=batch['areas'][idx] if 'areas' in batch else None)
    reg_loss = regression[1].mean()
    loss = None
    if not evaluation:
        loss = ((params['nsp_loss_coeff'] * nsp_loss) + (params['reg_loss_coeff'] * reg_loss)).sum()   # :144-153
        return loss, lm_loss, nsp_loss, img_loss, nsp_scores, regression, legend_loss
    return loss, lm_loss, nsp_loss, img_loss, nsp_scores, regression
