"""Python binding of the library's whole-model entry points (include/crct_b200.h: crct_create / crct_bind_params /
crct_workspace_bytes / crct_forward) — the INFERENCE forward scheduled by the library itself (csrc/model.cu).

This is what a host without the Python schedule of `cqa_crct_b200.encoder` binds: one call per batch, raw device pointers.
`CModel` exists for the parity tests (its outputs equal `VisualDialogEncoder.forward` in evaluation mode bit for bit) and as a
low-latency path for small evaluation batches (one question with its candidates: ~300 launches enqueued from C++, no Python in
between).  Reference boundary: CRCT/backbone/encoder_decorator.py:73-158 with `evaluation=True`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L
from .spec import ModelConfig

_MASK_KIND = {torch.bool: 0, torch.uint8: 0, torch.int64: 1, torch.float32: 2}


def config_args(cfg: ModelConfig, params: dict) -> L.ConfigArgs:
    c = L.ConfigArgs()
    for k in ('hidden_size', 'num_hidden_layers', 'num_attention_heads', 'intermediate_size', 'v_hidden_size', 'v_num_hidden_layers',
              'v_num_attention_heads', 'v_intermediate_size', 'v_feature_size', 'bi_hidden_size', 'bi_num_attention_heads',
              'max_position_embeddings'):
        setattr(c, k, int(getattr(cfg, k)))
    n = len(cfg.v_biattention_id)
    if n > L.MAX_CONNECTIONS:
        raise L.CrctError(f'{n} connection layers (the C ABI holds up to {L.MAX_CONNECTIONS})')
    c.num_connections = n
    for i in range(n):
        c.v_biattention_id[i], c.t_biattention_id[i] = int(cfg.v_biattention_id[i]), int(cfg.t_biattention_id[i])
    c.l1, c.tol_margin = int(bool(params.get('L1', False))), float(params.get('tol_margin', 0.01))
    return c


class CModel:
    """`crct_handle_t` bound to the parameter arenas of a `VisualDialogEncoder` (fp32 masters + bf16 operand copy)."""

    def __init__(self, encoder):
        self.enc = encoder
        arena = encoder.arena
        if not arena.w32.is_cuda:
            raise L.CrctError('cqa_crct_b200 has no CPU path: move the model to a B200 with .to("cuda")')
        self._h = C.c_void_p()
        cfg = config_args(encoder.cfg, encoder.params)
        L.check(L.lib().crct_create(C.byref(cfg), C.byref(self._h)))
        self._ws: Optional[torch.Tensor] = None
        self.bind()

    def bind(self):
        """(Re-)bind every live parameter by name: fp32 master and bf16 copy, as the arena lays them out."""
        arena = self.enc.arena
        arena.ensure_device_buffers()
        live = [p for p in arena.order if p.live]
        n = len(live)
        names = (C.c_char_p * n)(*[p.name.encode() for p in live])
        w32 = (C.c_void_p * n)(*[arena.w32.data_ptr() + 4 * arena.offsets[p.name] for p in live])
        w16 = (C.c_void_p * n)(*[arena.w16.data_ptr() + 2 * arena.offsets[p.name] for p in live])
        numel = (C.c_size_t * n)(*[p.numel for p in live])
        L.check(L.lib().crct_bind_params(self._h, names, w32, w16, numel, n))
        self._bound = (arena.w32.data_ptr(), arena.w16.data_ptr())

    def forward(self, batch: Dict[str, torch.Tensor], group: Optional[torch.Tensor] = None, fill=(0.0, 0.0)) -> Dict[str, torch.Tensor]:
        """`batch`: device tensors tokens / segments / loc / attention_mask / image_feat / image_loc / image_target / image_mask / R
        (the arguments `glue_forward(..., evaluation=True)` hands to the model); `group` [B] int64 = question of every candidate
        when the visual tensors are per question (f3).  Returns logits [B,2], reg_pred / reg_loss / reg_l1 / reg_dist [B], scalars [5]."""
        arena = self.enc.arena
        arena.refresh_bf16()
        if self._bound != (arena.w32.data_ptr(), arena.w16.data_ptr()):
            self.bind()
        dev = arena.w32.device
        want = {'tokens': torch.int64, 'segments': torch.int64, 'image_target': torch.int64, 'loc': torch.float32, 'image_feat': torch.float32,
                'image_loc': torch.float32, 'R': torch.float32}
        t = {}
        for k, v in batch.items():
            if k in want:
                v = v.to(device=dev, dtype=want[k])
            elif v.dtype not in _MASK_KIND:                      # masks: bool / uint8 / int64 / fp32 are read as they are
                v = v.to(device=dev, dtype=torch.int64)
            t[k] = v.to(dev).contiguous()
        if t['image_loc'].shape[-1] != 4 or t['image_feat'].shape[-1] != self.enc.cfg.v_feature_size:
            raise ValueError('image_loc must be [Bq,R,4] and image_feat [Bq,R,v_feature_size]')
        B, T = t['tokens'].shape
        Bq, R = t['image_feat'].shape[:2]
        a = L.BatchArgs()
        a.tokens, a.segments, a.loc, a.attention_mask = L.ptr(t['tokens']), L.ptr(t['segments']), L.ptr(t['loc']), L.ptr(t['attention_mask'])
        a.image_feat, a.image_loc, a.image_target, a.image_mask = (L.ptr(t['image_feat']), L.ptr(t['image_loc']), L.ptr(t['image_target']),
                                                                   L.ptr(t['image_mask']))
        a.R4 = L.ptr(t['R'])
        if group is not None:
            group = group.to(torch.int64).contiguous()
        a.group = L.ptr(group)
        a.B, a.Bq, a.T, a.R = B, Bq, T, R
        a.attention_mask_kind, a.image_mask_kind = _MASK_KIND[t['attention_mask'].dtype], _MASK_KIND[t['image_mask'].dtype]
        a.text_fill, a.region_fill = float(fill[0]), float(fill[1])
        need = int(L.lib().crct_workspace_bytes(self._h, B, Bq, T, R))
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        out = {'logits': torch.empty(B, 2, dtype=torch.float32, device=dev), 'scalars': torch.empty(5, dtype=torch.float32, device=dev)}
        for k in ('reg_pred', 'reg_loss', 'reg_l1', 'reg_dist'):
            out[k] = torch.empty(B, dtype=torch.float32, device=dev)
        o = L.OutArgs()
        o.logits, o.reg_pred, o.reg_loss, o.reg_l1, o.reg_dist, o.scalars = (L.ptr(out['logits']), L.ptr(out['reg_pred']), L.ptr(out['reg_loss']),
                                                                             L.ptr(out['reg_l1']), L.ptr(out['reg_dist']), L.ptr(out['scalars']))
        L.check(L.lib().crct_forward(self._h, C.byref(a), C.byref(o), L.ptr(self._ws), self._ws.numel(), L.stream_ptr()))
        self._keep = (t, group)           # inputs stay referenced until the next call (the work is only enqueued)
        return out

    def forward_graphed(self, batch: Dict[str, torch.Tensor], group: Optional[torch.Tensor] = None, fill=(0.0, 0.0), max_graphs: int = 8):
        """`forward` replayed from a CUDA graph: crct_forward only enqueues kernels on the caller's stream and reads its row counts on
        the device, so one capture per input SHAPE serves every batch of that shape (fixed-size interactive requests: one question
        with its candidates).  Inputs are copied into the capture's static buffers; the returned tensors are the capture's static
        outputs (overwritten by the next call with the same shapes)."""
        key = tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.items())) + (None if group is None else tuple(group.shape),)
        graphs = self.__dict__.setdefault('_graphs', {})
        ent = graphs.get(key)
        if ent is None:
            if len(graphs) >= max_graphs:
                graphs.pop(next(iter(graphs)))
            static = {k: v.clone() for k, v in batch.items()}
            sgroup = None if group is None else group.clone()
            self.forward(static, sgroup, fill)                      # warm-up outside the capture (lazy kernel attributes, workspace)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self.forward(static, sgroup, fill)
            ent = graphs[key] = (g, static, sgroup, out, self._bound)
        g, static, sgroup, out, bound = ent
        arena = self.enc.arena
        arena.refresh_bf16()
        if bound != (arena.w32.data_ptr(), arena.w16.data_ptr()):   # the parameters moved: captured pointers are stale
            graphs.clear()
            return self.forward_graphed(batch, group, fill, max_graphs)
        for k, v in batch.items():
            static[k].copy_(v, non_blocking=True)
        if group is not None:
            sgroup.copy_(group, non_blocking=True)
        g.replay()
        return out

    def __del__(self):
        try:
            if self._h:
                L.lib().crct_destroy(self._h)
        except Exception:
            pass
