"""Parameter inventory of the CRCT question-answering model.

`param_spec` lists every checkpoint tensor in the reference's `named_parameters()` order
(reference: CRCT/backbone/vilbert.py:297-317, 1444-1472, 361-377, 417-423, 443-466, 619-649,
728-744, 761-772, 949-968, 1017-1046, 1065-1072; CRCT/backbone/regressor.py:5-34), so that
`state_dict()` keys, optimizer param-group order (CRCT/utils.py:228-249) and checkpoints
(CRCT/train.py:284-291) stay interchangeable with the reference.

`arena_order` is the *memory* order: forward-execution order with the tensors of one fused
GEMM adjacent (q,k,v weights then q,k,v biases), live tensors first and the 36 never-used
tensors last, so that (a) a fused QKV weight is a zero-copy view, (b) the
backward pass finishes gradient regions from the tail towards the head — which is what the
bucketed all-reduce overlaps with — and (c) dead tensors are excluded from the all-reduce
without `find_unused_parameters` (reference: CRCT/train.py:139-142).
"""
from __future__ import annotations

import json
import math
import zlib
from dataclasses import dataclass
from typing import Dict, List, Tuple

import torch


class ModelConfig:
    """Same key handling as `BertConfig.from_json_file` (vilbert.py:245-258): JSON over defaults."""

    _DEFAULTS = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                     hidden_act='gelu', hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                     max_position_embeddings=512, type_vocab_size=2, initializer_range=0.02,
                     v_feature_size=1024, v_target_size=1601, v_hidden_size=768, v_num_hidden_layers=3,
                     v_num_attention_heads=12, v_intermediate_size=3072, bi_hidden_size=1024,
                     bi_num_attention_heads=16, v_attention_probs_dropout_prob=0.1, v_hidden_act='gelu',
                     v_hidden_dropout_prob=0.1, v_initializer_range=0.2, v_biattention_id=[0, 1],
                     t_biattention_id=[10, 11], fusion_method='mul', plotqa_vocab_types=12, vocab_size=30522)

    def __init__(self, src):
        if isinstance(src, str):
            with open(src, 'r', encoding='utf-8') as f:
                src = json.load(f)
        d = dict(self._DEFAULTS)
        d.update(src)
        self.__dict__.update(d)
        if len(self.v_biattention_id) != len(self.t_biattention_id):
            raise ValueError('v_biattention_id / t_biattention_id length mismatch')     # vilbert.py:192
        if max(self.v_biattention_id) >= self.v_num_hidden_layers or max(self.t_biattention_id) >= self.num_hidden_layers:
            raise ValueError('biattention id beyond layer count')                         # vilbert.py:193-194
        if self.hidden_act != 'gelu' or self.v_hidden_act != 'gelu' or self.fusion_method != 'mul':
            raise ValueError('only gelu activations and "mul" fusion are implemented (vilbert.json)')
        for h, n in ((self.hidden_size, self.num_attention_heads), (self.v_hidden_size, self.v_num_attention_heads),
                     (self.bi_hidden_size, self.bi_num_attention_heads)):
            if h % n:
                raise ValueError(f'hidden size {h} not a multiple of heads {n}')          # vilbert.py:364-368

    def schedule(self) -> List[Tuple[str, int]]:
        """Flat execution order of `BertEncoder.forward` (vilbert.py:852-939)."""
        sched, v0, t0 = [], 0, 0
        for count, (v1, t1) in enumerate(zip(self.v_biattention_id, self.t_biattention_id)):
            sched += [('v', i) for i in range(v0, v1)] + [('t', i) for i in range(t0, t1)] + [('c', count)]
            v0, t0 = v1, t1
        sched += [('v', i) for i in range(v0, self.v_num_hidden_layers)]
        sched += [('t', i) for i in range(t0, self.num_hidden_layers)]
        return sched


@dataclass(frozen=True)
class P:
    name: str
    shape: Tuple[int, ...]
    kind: str          # 'w' Linear weight | 'b' bias | 'emb' Embedding | 'lnw' | 'lnb' | 'rw' / 'rb' regressor Linear
    live: bool = True  # receives a gradient (False = the reference's never-used tensors)

    @property
    def numel(self) -> int:
        return int(math.prod(self.shape))


def _lin(pre, o, i, live=True, reg=False):
    return [P(pre + '.weight', (o, i), 'rw' if reg else 'w', live), P(pre + '.bias', (o,), 'rb' if reg else 'b', live)]


def _ln(pre, h, live=True):
    return [P(pre + '.weight', (h,), 'lnw', live), P(pre + '.bias', (h,), 'lnb', live)]


def _self_layer(pre, h, inter):
    out = []
    for n in ('query', 'key', 'value'):
        out += _lin(f'{pre}.attention.self.{n}', h, h)
    out += _lin(f'{pre}.attention.output.dense', h, h) + _ln(f'{pre}.attention.output.LayerNorm', h)
    out += _lin(f'{pre}.intermediate.dense', inter, h)
    out += _lin(f'{pre}.output.dense', h, inter) + _ln(f'{pre}.output.LayerNorm', h)
    return out


def param_spec(cfg: ModelConfig, categories: int = 228) -> List[P]:
    H, Hv, Hb, F = cfg.hidden_size, cfg.v_hidden_size, cfg.bi_hidden_size, cfg.v_feature_size
    s: List[P] = []
    e = 'bert.embeddings'
    s += [P(f'{e}.word_embeddings.weight', (cfg.vocab_size, H), 'emb'),
          P(f'{e}.position_embeddings.weight', (cfg.max_position_embeddings, H), 'emb')]
    s += _lin(f'{e}.txt_location_embeddings', H, 4)
    s += [P(f'{e}.plotqa_type_embeddings.weight', (cfg.plotqa_vocab_types, H), 'emb')] + _ln(f'{e}.LayerNorm', H)
    e = 'bert.v_embeddings'
    s += _lin(f'{e}.new_image_embeddings', Hv, F)
    s += [P(f'{e}.type_embeddings.weight', (13, Hv), 'emb', live=False),
          P(f'{e}.color_emb.weight', (categories + 1, Hv), 'emb')]
    s += _lin(f'{e}.new_loc_emb', Hv, 4) + _ln(f'{e}.LayerNorm', Hv)
    for i in range(cfg.num_hidden_layers):
        s += _self_layer(f'bert.encoder.layer.{i}', H, cfg.intermediate_size)
    for i in range(cfg.v_num_hidden_layers):
        s += _self_layer(f'bert.encoder.v_layer.{i}', Hv, cfg.v_intermediate_size)
    for i in range(len(cfg.v_biattention_id)):
        c = f'bert.encoder.c_layer.{i}'
        for n in ('query1', 'key1', 'value1'):
            s += _lin(f'{c}.biattention.{n}', Hb, Hv)
        for n in ('query2', 'key2', 'value2'):
            s += _lin(f'{c}.biattention.{n}', Hb, H)
        s += _lin(f'{c}.biOutput.dense1', Hv, Hb) + _ln(f'{c}.biOutput.LayerNorm1', Hv)
        s += _lin(f'{c}.biOutput.q_dense1', Hv, Hb, live=False)
        s += _lin(f'{c}.biOutput.dense2', H, Hb) + _ln(f'{c}.biOutput.LayerNorm2', H)
        s += _lin(f'{c}.biOutput.q_dense2', H, Hb, live=False)
        s += _lin(f'{c}.v_intermediate.dense', cfg.v_intermediate_size, Hv)
        s += _lin(f'{c}.v_output.dense', Hv, cfg.v_intermediate_size) + _ln(f'{c}.v_output.LayerNorm', Hv)
        s += _lin(f'{c}.t_intermediate.dense', cfg.intermediate_size, H)
        s += _lin(f'{c}.t_output.dense', H, cfg.intermediate_size) + _ln(f'{c}.t_output.LayerNorm', H)
    s += _lin('bert.t_pooler.dense', Hb, H) + _lin('bert.v_pooler.dense', Hb, Hv)
    s += [P('cls.predictions.bias', (cfg.vocab_size,), 'b', live=False)]
    s += _lin('cls.predictions.transform.dense', H, H, live=False) + _ln('cls.predictions.transform.LayerNorm', H, live=False)
    s += _lin('cls.bi_seq_relationship', 2, Hb)
    s += _lin('cls.imagePredictions.transform.dense', Hv, Hv, live=False)
    s += _ln('cls.imagePredictions.transform.LayerNorm', Hv, live=False)
    s += _lin('cls.imagePredictions.decoder', cfg.v_target_size, Hv, live=False)
    for pipe, h in (('txt_pipe', H), ('vis_pipe', Hv)):
        dims = [(h, h), (512, h), (256, 512), (256, 256)]
        for idx, (o, i) in zip((0, 2, 4, 6), dims):
            s += _lin(f'regressor.{pipe}.{idx}', o, i, reg=True)
    for idx, (o, i) in zip((0, 2, 4, 6), [(512, 512), (256, 512), (256, 256), (1, 256)]):
        s += _lin(f'regressor.fusion.{idx}', o, i, reg=True)
    return s


TIED = {'cls.predictions.decoder.weight': 'bert.embeddings.word_embeddings.weight'}   # vilbert.py:1024-1029


def fused_groups(cfg: ModelConfig) -> List[List[str]]:
    """Module prefixes whose weights (and biases) must sit back to back in the arena."""
    g = []
    for i in range(cfg.num_hidden_layers):
        g.append([f'bert.encoder.layer.{i}.attention.self.{n}' for n in ('query', 'key', 'value')])
    for i in range(cfg.v_num_hidden_layers):
        g.append([f'bert.encoder.v_layer.{i}.attention.self.{n}' for n in ('query', 'key', 'value')])
    for i in range(len(cfg.v_biattention_id)):
        g.append([f'bert.encoder.c_layer.{i}.biattention.{n}' for n in ('query1', 'key1', 'value1')])
        g.append([f'bert.encoder.c_layer.{i}.biattention.{n}' for n in ('query2', 'key2', 'value2')])
    return g


def arena_order(cfg: ModelConfig, spec: List[P]) -> List[P]:
    by_name = {p.name: p for p in spec}
    fused_first = {grp[0]: grp for grp in fused_groups(cfg)}
    fused_rest = {m for grp in fused_groups(cfg) for m in grp[1:]}

    def module_of(name):
        return name.rsplit('.', 1)[0]

    def block(prefix):
        """All live tensors under `prefix`, spec order, fused groups re-packed."""
        out, seen = [], set()
        for p in spec:
            if not p.live or not p.name.startswith(prefix + '.') or p.name in seen:
                continue
            m = module_of(p.name)
            if m in fused_rest:
                continue
            if m in fused_first:
                grp = fused_first[m]
                out += [by_name[x + '.weight'] for x in grp] + [by_name[x + '.bias'] for x in grp]
                seen.update(x + s for x in grp for s in ('.weight', '.bias'))
            else:
                out.append(p)
                seen.add(p.name)
        return out

    order = block('bert.embeddings') + block('bert.v_embeddings')
    for kind, i in cfg.schedule():
        order += block({'t': f'bert.encoder.layer.{i}', 'v': f'bert.encoder.v_layer.{i}',
                        'c': f'bert.encoder.c_layer.{i}'}[kind])
    order += block('bert.t_pooler') + block('bert.v_pooler') + block('cls') + block('regressor')
    live_names = {p.name for p in order}
    assert live_names == {p.name for p in spec if p.live}, 'arena order lost a live tensor'
    order += [p for p in spec if not p.live]
    return order


ALIGN = 64   # elements; 128 B in bf16, 256 B in fp32 — TMA needs 16 B, vector loads 16 B


def arena_offsets(order: List[P]) -> Tuple[Dict[str, int], int, int]:
    """name -> element offset; returns (offsets, live_end, total)."""
    off, cur, live_end = {}, 0, 0
    for p in order:
        off[p.name] = cur
        cur += (p.numel + ALIGN - 1) // ALIGN * ALIGN
        if p.live:
            live_end = cur
    return off, live_end, cur


def synth_tensor(p: P, seed: int, style: str, init_range: float = 0.02) -> torch.Tensor:
    """Deterministic per-tensor values, independent of tensor order.
    style 'reference': the reference constructor's distributions (vilbert.py:1099-1110; regressor keeps
    nn.Linear's default U(-1/sqrt(fan_in), 1/sqrt(fan_in)) because it is built after `apply(init)`,
    vilbert.py:1510 vs 1518-1523).
    style 'trained': same plus non-trivial biases / LayerNorm affine and 2x wider weights, so that
    bias-, gamma- and beta-paths cannot hide behind zeros in the parity tests (a deliberately ill-conditioned
    stress point: a random network this wide amplifies bf16 rounding ~45x into its gradients).
    style 'mild': non-trivial biases / LayerNorm affine on top of the reference's N(0, 0.02) weights."""
    g = torch.Generator().manual_seed((zlib.crc32(p.name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    trained = style in ('trained', 'mild')
    wide = 2.0 if style == 'trained' else 1.0
    if p.kind in ('w', 'emb'):
        return torch.randn(p.shape, generator=g) * (init_range * (wide if p.kind == 'w' else 1.0))
    if p.kind == 'b':
        return torch.randn(p.shape, generator=g) * 0.02 if trained else torch.zeros(p.shape)
    if p.kind == 'lnw':
        return 1.0 + 0.1 * torch.randn(p.shape, generator=g) if trained else torch.ones(p.shape)
    if p.kind == 'lnb':
        return 0.05 * torch.randn(p.shape, generator=g) if trained else torch.zeros(p.shape)
    if p.kind in ('rw', 'rb'):
        # weight bound 1/sqrt(fan_in); the bias only sees its own shape, a fixed small bound is used
        bound = 1.0 / math.sqrt(p.shape[1]) if p.kind == 'rw' else 0.05
        return (torch.rand(p.shape, generator=g) * 2 - 1) * bound
    raise ValueError(p.kind)


def synth_state_dict(cfg: ModelConfig, categories: int = 228, seed: int = 0, style: str = 'trained') -> Dict[str, torch.Tensor]:
    """Checkpoint-keyed (no `bert_pretrained.` prefix) fp32 weights, incl. the tied decoder key."""
    sd = {p.name: synth_tensor(p, seed, style, cfg.initializer_range) for p in param_spec(cfg, categories)}
    for k, src in TIED.items():
        sd[k] = sd[src]
    return sd
