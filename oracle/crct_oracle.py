"""TEST INFRASTRUCTURE ONLY — CPU restatement (torch fp32/fp64, no autograd) of the reference's
question-answering stage: forward AND hand-derived backward.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker.  The product path
(`cqa_crct_b200.encoder`) never imports it and fails loudly without its CUDA library.

Parity status: the reference ships no tests, golden vectors or fixtures, so
this restatement is pinned against *outputs of the reference itself run in the build container*
(`oracle/make_golden.py` -> `tests/golden/*.pt`, validated live by
`tests/test_oracle_vs_reference.py` when /root/reference is present).

The decomposition deliberately mirrors the CUDA kernels one-to-one (fused QKV weight, GEMM
epilogue -> pre-LN `z`, separate LayerNorm, dense-masked regressor, explicit backward per
kernel), so every kernel has a function here that states what it must compute.

All `file:line` citations are relative to /root/reference/CRCT/backbone/.
Weights are a dict keyed like `BertForMultiModalPreTraining.state_dict()`
(i.e. the checkpoint keys without the `bert_pretrained.` prefix).
"""
from __future__ import annotations

import json
import math
from typing import Dict, List, Optional, Tuple

import torch

LN_EPS = 1e-12           # vilbert.py:282, eps inside the sqrt (vilbert.py:293)
MASK_NEG = -10000.0      # vilbert.py:1391,1396
LEAKY = 0.01             # nn.LeakyReLU() default, regressor.py:9


# --------------------------------------------------------------------------------------
# config
# --------------------------------------------------------------------------------------
class Config:
    """vilbert.py:127-258 (`BertConfig.from_json_file`): JSON keys over constructor defaults."""

    def __init__(self, path_or_dict):
        d = dict(hidden_size=768, num_hidden_layers=12, num_attention_heads=12, intermediate_size=3072,
                 max_position_embeddings=512, v_feature_size=1024, v_hidden_size=768,
                 v_num_hidden_layers=3, v_num_attention_heads=12, v_intermediate_size=3072,
                 bi_hidden_size=1024, bi_num_attention_heads=16, v_biattention_id=[0, 1],
                 t_biattention_id=[10, 11], hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1,
                 v_hidden_dropout_prob=0.1, v_attention_probs_dropout_prob=0.1, plotqa_vocab_types=12,
                 vocab_size=30522, fusion_method='mul', initializer_range=0.02)
        if isinstance(path_or_dict, str):
            with open(path_or_dict) as f:
                path_or_dict = json.load(f)
        d.update(path_or_dict)
        self.__dict__.update(d)
        assert self.fusion_method == 'mul'               # vilbert.py:1054-1055
        assert len(self.v_biattention_id) == len(self.t_biattention_id)   # vilbert.py:192


def layer_schedule(cfg: Config) -> List[Tuple[str, int]]:
    """Execution order of `BertEncoder.forward` (vilbert.py:852-939) as a flat list of
    ('v', idx) / ('t', idx) / ('c', idx) steps.  fixed_*_layer = 0, with_coattention = True."""
    sched, v_start, t_start = [], 0, 0
    for count, (v_end, t_end) in enumerate(zip(cfg.v_biattention_id, cfg.t_biattention_id)):
        sched += [('v', i) for i in range(v_start, v_end)]
        sched += [('t', i) for i in range(t_start, t_end)]
        sched.append(('c', count))
        v_start, t_start = v_end, t_end
    sched += [('v', i) for i in range(v_start, cfg.v_num_hidden_layers)]
    sched += [('t', i) for i in range(t_start, cfg.num_hidden_layers)]
    return sched


# --------------------------------------------------------------------------------------
# optional storage-precision emulation
# --------------------------------------------------------------------------------------
# The reference arithmetic is the default (`_q` / `_qw` are identities).  Inside `bf16_emulation()` the
# restatement additionally rounds to bf16 exactly where the CUDA path STORES bf16: tensor-core GEMM operands
# (weights `_qw`, activations `_q`) and every activation / gradient tensor that goes through HBM in bf16.  The
# residual stream does not: the CUDA path keeps the pre-LayerNorm sums `z` and the LayerNorm outputs in fp32 (plus a
# bf16 copy of the latter as the next GEMM's A operand — the `_q(x)` inside `tlin`).  All
# reductions, statistics, softmax, the heads and the losses stay fp32/fp64, as in the kernels.  Comparing the
# CUDA path with this mode separates implementation error from the operand quantisation every bf16
# tensor-core path has (tests/test_model_gpu.py states both tolerances).
_EMU = [False]


class bf16_emulation:
    def __enter__(self):
        _EMU[0] = True

    def __exit__(self, *a):
        _EMU[0] = False


def _q(x):
    return x.to(torch.bfloat16).to(x.dtype) if _EMU[0] else x


_qw = _q

# --------------------------------------------------------------------------------------
# primitive ops (each = one CUDA kernel or GEMM epilogue)
# --------------------------------------------------------------------------------------
def gelu(x):                      # vilbert.py:111-117, exact erf form
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


def gelu_grad(x):
    return 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)


def ln_fwd(z, g, b):              # vilbert.py:290-294
    mean = z.mean(-1, keepdim=True)
    var = (z - mean).pow(2).mean(-1, keepdim=True)
    rstd = 1.0 / torch.sqrt(var + LN_EPS)
    return (z - mean) * rstd * g + b, mean, rstd


def ln_bwd(dy, z, mean, rstd, g):
    xhat = (z - mean) * rstd
    dg = (dy * xhat).reshape(-1, z.shape[-1]).sum(0)
    db = dy.reshape(-1, z.shape[-1]).sum(0)
    dxh = dy * g
    dz = rstd * (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True))
    return dz, dg, db


def split_heads(x, B, L, nh):     # vilbert.py:379-385
    return x.view(B, L, nh, -1).permute(0, 2, 1, 3)


def attn_fwd(q, k, v, add_mask, nh, drop=None):
    """softmax(q k^T / sqrt(d) + mask) v  (vilbert.py:397-412 / 527-543 / 684-723).
    q:[B,Lq,H] k,v:[B,Lk,H] add_mask:[B,Lk] additive.  Returns ctx [B,Lq,H], probs (post-dropout
    keep factor applied) and pre-dropout probs."""
    B, Lq, H = q.shape
    Lk = k.shape[1]
    d = H // nh
    qh, kh, vh = split_heads(q, B, Lq, nh), split_heads(k, B, Lk, nh), split_heads(v, B, Lk, nh)
    s = qh @ kh.transpose(-1, -2) / math.sqrt(d) + add_mask[:, None, None, :]
    p = torch.softmax(s, dim=-1)
    pd = p if drop is None else p * drop
    ctx = (_q(pd) @ vh).permute(0, 2, 1, 3).reshape(B, Lq, H)
    return _q(ctx), p, pd


def attn_bwd(dctx, q, k, v, p, pd, nh, drop=None):
    B, Lq, H = q.shape
    Lk = k.shape[1]
    d = H // nh
    qh, kh, vh = split_heads(q, B, Lq, nh), split_heads(k, B, Lk, nh), split_heads(v, B, Lk, nh)
    do = split_heads(dctx, B, Lq, nh)
    dv = _q(pd).transpose(-1, -2) @ do
    dpd = do @ vh.transpose(-1, -2)
    dp = dpd if drop is None else dpd * drop
    ds = _q(p * (dp - (dp * p).sum(-1, keepdim=True)) / math.sqrt(d))
    dq = ds @ kh
    dk = ds.transpose(-1, -2) @ qh
    merge = lambda t, L: _q(t.permute(0, 2, 1, 3).reshape(B, L, H))
    return merge(dq, Lq), merge(dk, Lk), merge(dv, Lk)


def lin(x, W, b):
    return x @ W.t() + b


def lin_bwd(dy, x, W):
    """dY -> (dX, dW, db) for y = x W^T + b: the dgrad GEMM, the wgrad GEMM and the column sum."""
    dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
    return dy @ W, dy2.t() @ x2, dy2.sum(0)


def tlin(x, W, b):
    """Linear on the tensor-core path (bf16 operands in emulation mode)."""
    return _q(x) @ _qw(W).t() + b          # the A operand is the bf16 copy of the (fp32) residual stream


def tlin_bwd(dy, x, W):
    dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
    return dy @ _qw(W), dy2.t() @ _q(x2), dy2.sum(0)


class _Grads(dict):
    def add(self, key, val):
        self[key] = self[key] + val if key in self else val


# --------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------
def text_position_ids(token_type_ids):
    """vilbert.py:327-335 — positions count 0,1,2.. only over question (-1) / answer (1) tokens."""
    B, T = token_type_ids.shape
    qa = (token_type_ids == -1) | (token_type_ids == 1)
    pos = torch.arange(T).unsqueeze(0).expand(B, T).clone()
    pos[~qa] = T
    pos = pos - pos.min(dim=-1)[0].unsqueeze(1)
    pos[~qa] = 0
    return pos, qa


def embed_text_fwd(w, ids, types, loc, pre='bert.embeddings.'):
    """vilbert.py:320-358."""
    pos, qa = text_position_ids(types)
    pe = w[pre + 'position_embeddings.weight'][pos] * qa.unsqueeze(-1)
    we = w[pre + 'word_embeddings.weight'][ids]
    loc_on = (loc.abs().sum(-1) != 0).unsqueeze(-1)
    le = lin(loc, w[pre + 'txt_location_embeddings.weight'], w[pre + 'txt_location_embeddings.bias']) * loc_on
    ty = types.clone()
    ty[ty == -1] = 0
    ty_on = (types != 0).unsqueeze(-1)
    te = w[pre + 'plotqa_type_embeddings.weight'][ty] * ty_on
    z = we + pe + te + le                                # pre-LayerNorm sums stay fp32 on the CUDA path
    y, mean, rstd = ln_fwd(z, w[pre + 'LayerNorm.weight'], w[pre + 'LayerNorm.bias'])
    return y, dict(z=z, mean=mean, rstd=rstd, pos=pos, qa=qa, loc_on=loc_on, ty=ty, ty_on=ty_on)


def embed_text_bwd(w, g, dy, c, ids, loc, pre='bert.embeddings.'):
    dz, dg, db = ln_bwd(dy, c['z'], c['mean'], c['rstd'], w[pre + 'LayerNorm.weight'])
    dz = _q(dz)
    g.add(pre + 'LayerNorm.weight', dg)
    g.add(pre + 'LayerNorm.bias', db)
    H = dz.shape[-1]
    dz2 = dz.reshape(-1, H)
    gw = torch.zeros_like(w[pre + 'word_embeddings.weight'])
    gw.index_add_(0, ids.reshape(-1), dz2)
    g.add(pre + 'word_embeddings.weight', gw)
    gp = torch.zeros_like(w[pre + 'position_embeddings.weight'])
    gp.index_add_(0, c['pos'].reshape(-1), (dz * c['qa'].unsqueeze(-1)).reshape(-1, H))
    g.add(pre + 'position_embeddings.weight', gp)
    gt = torch.zeros_like(w[pre + 'plotqa_type_embeddings.weight'])
    gt.index_add_(0, c['ty'].reshape(-1), (dz * c['ty_on']).reshape(-1, H))
    g.add(pre + 'plotqa_type_embeddings.weight', gt)
    dl = (dz * c['loc_on']).reshape(-1, H)
    g.add(pre + 'txt_location_embeddings.weight', dl.t() @ loc.reshape(-1, 4))
    g.add(pre + 'txt_location_embeddings.bias', dl.sum(0))


def embed_vis_fwd(w, feat, box, cls, pre='bert.v_embeddings.'):
    """vilbert.py:1474-1496 (PlotQA branch: img + loc + color, no areas, mask_prob_img = 0)."""
    p = _q(torch.softmax(feat, dim=-1))
    z = (_q(tlin(p, w[pre + 'new_image_embeddings.weight'], w[pre + 'new_image_embeddings.bias']))
         + lin(box, w[pre + 'new_loc_emb.weight'], w[pre + 'new_loc_emb.bias'])
         + w[pre + 'color_emb.weight'][cls])
    y, mean, rstd = ln_fwd(z, w[pre + 'LayerNorm.weight'], w[pre + 'LayerNorm.bias'])
    return y, dict(p=p, z=z, mean=mean, rstd=rstd)


def embed_vis_bwd(w, g, dy, c, box, cls, pre='bert.v_embeddings.'):
    dz, dg, db = ln_bwd(dy, c['z'], c['mean'], c['rstd'], w[pre + 'LayerNorm.weight'])
    dz = _q(dz)
    g.add(pre + 'LayerNorm.weight', dg)
    g.add(pre + 'LayerNorm.bias', db)
    H = dz.shape[-1]
    dz2 = dz.reshape(-1, H)
    g.add(pre + 'new_image_embeddings.weight', dz2.t() @ c['p'].reshape(-1, c['p'].shape[-1]))
    g.add(pre + 'new_image_embeddings.bias', dz2.sum(0))
    g.add(pre + 'new_loc_emb.weight', dz2.t() @ box.reshape(-1, 4))
    g.add(pre + 'new_loc_emb.bias', dz2.sum(0))
    gc = torch.zeros_like(w[pre + 'color_emb.weight'])
    gc.index_add_(0, cls.reshape(-1), dz2)
    g.add(pre + 'color_emb.weight', gc)


def _qkv(w, pre, names):
    W = torch.cat([w[pre + n + '.weight'] for n in names], 0)
    b = torch.cat([w[pre + n + '.bias'] for n in names], 0)
    return W, b


def ffn_fwd(w, a, pre_i, pre_o):
    """intermediate (vilbert.py:454-457 / 585-588) + output (vilbert.py:467-471 / 598-602)."""
    u = tlin(a, w[pre_i + 'dense.weight'], w[pre_i + 'dense.bias'])
    h = _q(gelu(u))
    z = tlin(h, w[pre_o + 'dense.weight'], w[pre_o + 'dense.bias']) + a
    y, mean, rstd = ln_fwd(z, w[pre_o + 'LayerNorm.weight'], w[pre_o + 'LayerNorm.bias'])
    return y, dict(a=a, u=u, h=h, z=z, mean=mean, rstd=rstd, y32=y)


def ffn_bwd(w, g, dy, c, pre_i, pre_o):
    dz, dg, db = ln_bwd(dy, c['z'], c['mean'], c['rstd'], w[pre_o + 'LayerNorm.weight'])
    g.add(pre_o + 'LayerNorm.weight', dg)
    g.add(pre_o + 'LayerNorm.bias', db)
    g.add(pre_o + 'dense.bias', dz.reshape(-1, dz.shape[-1]).sum(0))      # summed before the bf16 store
    dz = _q(dz)
    dh, dW, _ = tlin_bwd(dz, c['h'], w[pre_o + 'dense.weight'])
    g.add(pre_o + 'dense.weight', dW)
    du = _q(dh * _q(gelu_grad(c['u'])))      # the CUDA path stores gelu'(u) in bf16 from the forward epilogue
    da, dW, dbias = tlin_bwd(du, c['a'], w[pre_i + 'dense.weight'])
    g.add(pre_i + 'dense.weight', dW)
    g.add(pre_i + 'dense.bias', dbias)
    return _q(da + dz)


def self_layer_fwd(w, x, add_mask, nh, pre):
    """BertLayer / BertImageLayer (vilbert.py:474-485, 605-616)."""
    H = x.shape[-1]
    Wqkv, bqkv = _qkv(w, pre + 'attention.self.', ['query', 'key', 'value'])
    qkv = _q(tlin(x, Wqkv, bqkv))
    q, k, v = qkv[..., :H], qkv[..., H:2 * H], qkv[..., 2 * H:]
    ctx, p, pd = attn_fwd(q, k, v, add_mask, nh)
    po = pre + 'attention.output.'
    z1 = tlin(ctx, w[po + 'dense.weight'], w[po + 'dense.bias']) + x
    a, mean1, rstd1 = ln_fwd(z1, w[po + 'LayerNorm.weight'], w[po + 'LayerNorm.bias'])
    y, cf = ffn_fwd(w, a, pre + 'intermediate.', pre + 'output.')
    return y, dict(x=x, q=q, k=k, v=v, p=p, pd=pd, ctx=ctx, z1=z1, mean1=mean1, rstd1=rstd1, ffn=cf, nh=nh)


def self_layer_bwd(w, g, dy, c, pre):
    da = ffn_bwd(w, g, dy, c['ffn'], pre + 'intermediate.', pre + 'output.')
    po = pre + 'attention.output.'
    dz1, dg, db = ln_bwd(da, c['z1'], c['mean1'], c['rstd1'], w[po + 'LayerNorm.weight'])
    g.add(po + 'LayerNorm.weight', dg)
    g.add(po + 'LayerNorm.bias', db)
    g.add(po + 'dense.bias', dz1.reshape(-1, dz1.shape[-1]).sum(0))
    dz1 = _q(dz1)
    dctx, dW, _ = tlin_bwd(dz1, c['ctx'], w[po + 'dense.weight'])
    dctx = _q(dctx)
    g.add(po + 'dense.weight', dW)
    dq, dk, dv = attn_bwd(dctx, c['q'], c['k'], c['v'], c['p'], c['pd'], c['nh'])
    dqkv = torch.cat([dq, dk, dv], -1)
    Wqkv, _ = _qkv(w, pre + 'attention.self.', ['query', 'key', 'value'])
    dx, dW, dbias = tlin_bwd(dqkv, c['x'], Wqkv)
    H = c['x'].shape[-1]
    for i, n in enumerate(['query', 'key', 'value']):
        g.add(pre + f'attention.self.{n}.weight', dW[i * H:(i + 1) * H])
        g.add(pre + f'attention.self.{n}.bias', dbias[i * H:(i + 1) * H])
    return _q(dx + dz1)


def co_layer_fwd(w, v, t, v_mask, t_mask, nh, pre):
    """BertConnectionLayer (vilbert.py:774-788): stream 1 = visual, stream 2 = text;
    ctx1 = text queries over visual keys/values, ctx2 = visual queries over text keys/values
    (vilbert.py:684-723); biOutput is called with crossed arguments (vilbert.py:780)."""
    pb = pre + 'biattention.'
    W1, b1 = _qkv(w, pb, ['query1', 'key1', 'value1'])
    W2, b2 = _qkv(w, pb, ['query2', 'key2', 'value2'])
    Hb = W1.shape[0] // 3
    qkv1, qkv2 = _q(tlin(v, W1, b1)), _q(tlin(t, W2, b2))
    q1, k1, v1 = qkv1[..., :Hb], qkv1[..., Hb:2 * Hb], qkv1[..., 2 * Hb:]
    q2, k2, v2 = qkv2[..., :Hb], qkv2[..., Hb:2 * Hb], qkv2[..., 2 * Hb:]
    ctx1, p1, pd1 = attn_fwd(q2, k1, v1, v_mask, nh)       # [B,T,Hb]
    ctx2, p2, pd2 = attn_fwd(q1, k2, v2, t_mask, nh)       # [B,R,Hb]
    po = pre + 'biOutput.'
    zv = tlin(ctx2, w[po + 'dense1.weight'], w[po + 'dense1.bias']) + v
    av, mv, rv = ln_fwd(zv, w[po + 'LayerNorm1.weight'], w[po + 'LayerNorm1.bias'])
    zt = tlin(ctx1, w[po + 'dense2.weight'], w[po + 'dense2.bias']) + t
    at, mt, rt = ln_fwd(zt, w[po + 'LayerNorm2.weight'], w[po + 'LayerNorm2.bias'])
    yv, cfv = ffn_fwd(w, av, pre + 'v_intermediate.', pre + 'v_output.')
    yt, cft = ffn_fwd(w, at, pre + 't_intermediate.', pre + 't_output.')
    c = dict(v=v, t=t, q1=q1, k1=k1, v1=v1, q2=q2, k2=k2, v2=v2, p1=p1, pd1=pd1, p2=p2, pd2=pd2,
             ctx1=ctx1, ctx2=ctx2, zv=zv, mv=mv, rv=rv, zt=zt, mt=mt, rt=rt, ffv=cfv, fft=cft, nh=nh)
    return yv, yt, c


def co_layer_bwd(w, g, dyv, dyt, c, pre):
    dav = ffn_bwd(w, g, dyv, c['ffv'], pre + 'v_intermediate.', pre + 'v_output.')
    dat = ffn_bwd(w, g, dyt, c['fft'], pre + 't_intermediate.', pre + 't_output.')
    po = pre + 'biOutput.'
    dzv, dg, db = ln_bwd(dav, c['zv'], c['mv'], c['rv'], w[po + 'LayerNorm1.weight'])
    g.add(po + 'LayerNorm1.weight', dg)
    g.add(po + 'LayerNorm1.bias', db)
    dzt, dg, db = ln_bwd(dat, c['zt'], c['mt'], c['rt'], w[po + 'LayerNorm2.weight'])
    g.add(po + 'LayerNorm2.weight', dg)
    g.add(po + 'LayerNorm2.bias', db)
    g.add(po + 'dense1.bias', dzv.reshape(-1, dzv.shape[-1]).sum(0))
    g.add(po + 'dense2.bias', dzt.reshape(-1, dzt.shape[-1]).sum(0))
    dzv, dzt = _q(dzv), _q(dzt)
    dctx2, dW, _ = tlin_bwd(dzv, c['ctx2'], w[po + 'dense1.weight'])
    g.add(po + 'dense1.weight', dW)
    dctx1, dW, _ = tlin_bwd(dzt, c['ctx1'], w[po + 'dense2.weight'])
    g.add(po + 'dense2.weight', dW)
    dctx1, dctx2 = _q(dctx1), _q(dctx2)
    dq2, dk1, dv1 = attn_bwd(dctx1, c['q2'], c['k1'], c['v1'], c['p1'], c['pd1'], c['nh'])
    dq1, dk2, dv2 = attn_bwd(dctx2, c['q1'], c['k2'], c['v2'], c['p2'], c['pd2'], c['nh'])
    pb = pre + 'biattention.'
    W1, _ = _qkv(w, pb, ['query1', 'key1', 'value1'])
    W2, _ = _qkv(w, pb, ['query2', 'key2', 'value2'])
    Hb = W1.shape[0] // 3
    dv_in, dW, dbias = tlin_bwd(torch.cat([dq1, dk1, dv1], -1), c['v'], W1)
    for i, n in enumerate(['query1', 'key1', 'value1']):
        g.add(pb + n + '.weight', dW[i * Hb:(i + 1) * Hb])
        g.add(pb + n + '.bias', dbias[i * Hb:(i + 1) * Hb])
    dt_in, dW, dbias = tlin_bwd(torch.cat([dq2, dk2, dv2], -1), c['t'], W2)
    for i, n in enumerate(['query2', 'key2', 'value2']):
        g.add(pb + n + '.weight', dW[i * Hb:(i + 1) * Hb])
        g.add(pb + n + '.bias', dbias[i * Hb:(i + 1) * Hb])
    return _q(dv_in + dzv), _q(dt_in + dzt)


def leaky(x):
    return torch.where(x > 0, x, LEAKY * x)


def mlp4_fwd(w, x, pre):
    """4 Linear layers with LeakyReLU between (regressor.py:8-33)."""
    acts, pres = [x], []
    h = x
    for i, idx in enumerate((0, 2, 4, 6)):
        u = lin(h, w[pre + f'{idx}.weight'], w[pre + f'{idx}.bias'])
        pres.append(u)
        h = leaky(u) if i < 3 else u
        acts.append(h)
    return h, dict(acts=acts, pres=pres)


def mlp4_bwd(w, g, dy, c, pre):
    d = dy
    for i, idx in reversed(list(enumerate((0, 2, 4, 6)))):
        if i < 3:
            d = d * torch.where(c['pres'][i] > 0, torch.ones_like(d), torch.full_like(d, LEAKY))
        d, dW, db = lin_bwd(d, c['acts'][i], w[pre + f'{idx}.weight'])
        g.add(pre + f'{idx}.weight', dW)
        g.add(pre + f'{idx}.bias', db)
    return d


def heads_fwd(w, t, v):
    """Poolers (vilbert.py:955-976), classifier (vilbert.py:1052-1060), regressor (regressor.py:36-42)
    evaluated densely on every row (rows without `needs_reg` are masked in the loss).  `t`, `v`: the UNROUNDED last
    LayerNorm outputs in emulation mode (the CUDA heads normalise the fp32 pre-LayerNorm rows themselves)."""
    hw0, hv0 = t[:, 0], v[:, 0]
    ut = lin(hw0, w['bert.t_pooler.dense.weight'], w['bert.t_pooler.dense.bias'])
    uv = lin(hv0, w['bert.v_pooler.dense.weight'], w['bert.v_pooler.dense.bias'])
    pt, pv = torch.relu(ut), torch.relu(uv)
    pooled = pt * pv
    logits = lin(pooled, w['cls.bi_seq_relationship.weight'], w['cls.bi_seq_relationship.bias'])
    hw, ct = mlp4_fwd(w, hw0, 'regressor.txt_pipe.')
    hv, cv = mlp4_fwd(w, hv0, 'regressor.vis_pipe.')
    pre = torch.cat([hv, hw], -1)                       # regressor.py:40
    f, cf = mlp4_fwd(w, pre, 'regressor.fusion.')
    reg = torch.tanh(f.squeeze(-1))
    return logits, reg, dict(hw0=hw0, hv0=hv0, ut=ut, uv=uv, pt=pt, pv=pv, pooled=pooled, ct=ct, cv=cv, cf=cf, reg=reg)


def heads_bwd(w, g, dlogits, dreg, c, T, R):
    B = dlogits.shape[0]
    dpooled, dW, db = lin_bwd(dlogits, c['pooled'], w['cls.bi_seq_relationship.weight'])
    g.add('cls.bi_seq_relationship.weight', dW)
    g.add('cls.bi_seq_relationship.bias', db)
    dut = dpooled * c['pv'] * (c['ut'] > 0)
    duv = dpooled * c['pt'] * (c['uv'] > 0)
    dhw0, dW, db = lin_bwd(dut, c['hw0'], w['bert.t_pooler.dense.weight'])
    g.add('bert.t_pooler.dense.weight', dW)
    g.add('bert.t_pooler.dense.bias', db)
    dhv0, dW, db = lin_bwd(duv, c['hv0'], w['bert.v_pooler.dense.weight'])
    g.add('bert.v_pooler.dense.weight', dW)
    g.add('bert.v_pooler.dense.bias', db)
    df = (dreg * (1.0 - c['reg'] ** 2)).unsqueeze(-1)
    dpre = mlp4_bwd(w, g, df, c['cf'], 'regressor.fusion.')
    nv = dpre.shape[-1] // 2
    dhv0 = dhv0 + mlp4_bwd(w, g, dpre[:, :nv], c['cv'], 'regressor.vis_pipe.')
    dhw0 = dhw0 + mlp4_bwd(w, g, dpre[:, nv:], c['ct'], 'regressor.txt_pipe.')
    dt = torch.zeros(B, T, dhw0.shape[-1], dtype=dhw0.dtype)
    dv = torch.zeros(B, R, dhv0.shape[-1], dtype=dhv0.dtype)
    dt[:, 0], dv[:, 0] = _q(dhw0), _q(dhv0)
    return dt, dv


def losses_fwd(logits, reg, labels, Rt, kind: str, l1: bool, tol_margin: float):
    """vilbert.py:1586-1657 restated densely.  Returns dict with the 5-element `reg` list contents,
    nsp loss, and d(loss)/d(logits), d(mean reg loss)/d(reg) for coefficient 1."""
    B = logits.shape[0]
    needs = Rt[:, 1] == 1
    scale = Rt[:, 3]
    tgt = torch.where(needs, Rt[:, 0] / scale, torch.zeros_like(scale))
    diff = reg - tgt
    l1v = diff.abs()
    if l1:                                              # vilbert.py:1525-1528
        lossv, dl = l1v, torch.sign(diff)
    else:                                               # SmoothL1, beta = 0.5
        beta = 0.5
        small = l1v < beta
        lossv = torch.where(small, 0.5 * diff * diff / beta, l1v - 0.5 * beta)
        dl = torch.where(small, diff / beta, torch.sign(diff))
    dist = l1v / tgt.abs()                              # vilbert.py:1632-1636
    dist = torch.where(tgt == 0, torch.ones_like(dist), dist)
    both0 = (reg == 0) & (tgt == 0)
    dist = torch.where(both0, torch.zeros_like(dist), dist)
    right5 = ((dist <= 0.05) | both0) & needs
    right_t = (l1v <= tol_margin) & needs
    live = needs.clone()
    if kind != 'L1':                                    # vilbert.py:1639-1641
        live = live & ~(tgt.abs() > 1)
    z = torch.zeros_like(reg)
    out = dict(
        reg_pred=torch.where(needs, reg * scale, z), reg_loss=torch.where(live, lossv, z),
        reg_l1=torch.where(needs, l1v, z), reg_right=(int(right5.sum()), int(right_t.sum())),
        reg_dist=torch.where(needs, dist, z), dreg=torch.where(live, dl, z) / B)
    if labels is not None:                              # vilbert.py:1513,1655-1657 (ignore_index = -1)
        lab = labels.view(-1)
        valid = lab != -1
        n = valid.sum().clamp(min=1)
        lsm = torch.log_softmax(logits, -1)
        pick = lsm.gather(1, lab.clamp(min=0).view(-1, 1)).squeeze(1)
        out['nsp_loss'] = -(pick * valid).sum() / n
        onehot = torch.zeros_like(logits).scatter_(1, lab.clamp(min=0).view(-1, 1), 1.0)
        out['dlogits'] = (torch.softmax(logits, -1) - onehot) * valid.unsqueeze(1) / n
    return out


def add_masks(attention_mask, image_attention_mask, dtype):
    """vilbert.py:1380-1396."""
    return ((1.0 - attention_mask.to(dtype)) * MASK_NEG, (1.0 - image_attention_mask.to(dtype)) * MASK_NEG)


def text_attention_mask(sep_indices, hist_len, T):
    """encoder_decorator.py:118-120 + sequence_mask (encoder_decorator.py:57-70)."""
    lens = torch.gather(sep_indices, 1, hist_len.view(-1, 1)).squeeze(1) + 1
    return torch.arange(T).unsqueeze(0) < lens.unsqueeze(1)


def forward(w: Dict[str, torch.Tensor], cfg: Config, batch: dict, train: bool, l1: bool = True,
            tol_margin: float = 0.01, nsp_coeff: float = 1.0, reg_coeff: float = 1.0,
            keep_cache: bool = True, dtype=torch.float32):
    """Whole question-answering stage: encoder_decorator.forward (encoder_decorator.py:73-158)
    + BertForMultiModalPreTraining.forward (vilbert.py:1540-1661), dropout off."""
    cast = lambda x: x.to(dtype)
    w = {k: cast(v) if v.is_floating_point() else v for k, v in w.items()}
    ids, types = batch['tokens'], batch['segments']
    loc, feat, box = cast(batch['loc']), cast(batch['image_feat']), cast(batch['image_loc'])
    cls = batch['image_target']
    B, T = ids.shape
    R = feat.shape[1]
    amask = text_attention_mask(batch['sep_indices'], batch['hist_len'], T)
    t_mask, v_mask = add_masks(amask, batch['image_mask'], dtype)
    t, ce_t = embed_text_fwd(w, ids, types, loc)
    v, ce_v = embed_vis_fwd(w, feat, box, cls)
    caches = []
    for kind, i in layer_schedule(cfg):
        if kind == 't':
            t, c = self_layer_fwd(w, t, t_mask, cfg.num_attention_heads, f'bert.encoder.layer.{i}.')
            t32 = c['ffn']['y32']
        elif kind == 'v':
            v, c = self_layer_fwd(w, v, v_mask, cfg.v_num_attention_heads, f'bert.encoder.v_layer.{i}.')
            v32 = c['ffn']['y32']
        else:
            v, t, c = co_layer_fwd(w, v, t, v_mask, t_mask, cfg.bi_num_attention_heads, f'bert.encoder.c_layer.{i}.')
            v32, t32 = c['ffv']['y32'], c['fft']['y32']
        caches.append(c if keep_cache else None)
    logits, reg, ch = heads_fwd(w, t32, v32)            # unrounded last LayerNorm outputs (== t, v outside emulation mode)
    kind = 'L1_smooth' if train else 'L1'               # encoder_decorator.py:104,106
    L = losses_fwd(logits, reg, batch['next_sentence_labels'] if train else None, cast(batch['R']),
                   kind, l1, tol_margin)
    out = dict(logits=logits, reg=reg, seq_t=t, seq_v=v, **{k: L[k] for k in
               ('reg_pred', 'reg_loss', 'reg_l1', 'reg_right', 'reg_dist')})
    if train:
        out['nsp_loss'] = L['nsp_loss']
        out['loss'] = nsp_coeff * L['nsp_loss'] + reg_coeff * L['reg_loss'].mean()   # encoder_decorator.py:144-153
    cache = dict(w=w, cfg=cfg, batch=batch, ce_t=ce_t, ce_v=ce_v, caches=caches, ch=ch, L=L, T=T, R=R,
                 loc=loc, box=box, nsp_coeff=nsp_coeff, reg_coeff=reg_coeff)
    return out, cache


def backward(cache) -> Dict[str, torch.Tensor]:
    """Gradients of `out['loss']` w.r.t. every weight that receives one (238.3 M of the 252.7 M
    parameters at the PlotQA config; the 36 dead tensors of SURVEY.md §2.3 get no entry)."""
    w, cfg, batch = cache['w'], cache['cfg'], cache['batch']
    g = _Grads()
    L = cache['L']
    dt, dv = heads_bwd(w, g, L['dlogits'] * cache['nsp_coeff'], L['dreg'] * cache['reg_coeff'],
                       cache['ch'], cache['T'], cache['R'])
    for (kind, i), c in reversed(list(zip(layer_schedule(cfg), cache['caches']))):
        if kind == 't':
            dt = self_layer_bwd(w, g, dt, c, f'bert.encoder.layer.{i}.')
        elif kind == 'v':
            dv = self_layer_bwd(w, g, dv, c, f'bert.encoder.v_layer.{i}.')
        else:
            dv, dt = co_layer_bwd(w, g, dv, dt, c, f'bert.encoder.c_layer.{i}.')
    embed_text_bwd(w, g, dt, cache['ce_t'], batch['tokens'], cache['loc'])
    embed_vis_bwd(w, g, dv, cache['ce_v'], cache['box'], batch['image_target'])
    return dict(g)
