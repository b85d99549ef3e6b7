"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, via oracle/ref_shim.py) in the build container.

    python oracle/make_golden.py            # writes tests/golden/{tiny,full}_*.pt

Each fixture holds the recipe (config file, seeds, weight style, shapes) and what the reference
produced: class logits, regression list, losses and, for train cases, a per-tensor summary of every
parameter gradient (L2 norm, sum and 8 sampled entries).  Weights and inputs are NOT stored: they
are regenerated from the seeds by `cqa_crct_b200.spec.synth_state_dict` / `synthetic.make_batch`
(CPU torch.Generator — same bits here and on the GPU box).
"""
from __future__ import annotations

import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim                                   # noqa: E402
from cqa_crct_b200.spec import ModelConfig, synth_state_dict   # noqa: E402
from cqa_crct_b200.synthetic import make_batch, default_params  # noqa: E402

CASES = [
    # name, config, B, T, R, l1, weight seed, batch seed, train
    ('tiny_train_l1', 'tiny.json', 6, 32, 12, True, 1, 11, True),
    ('tiny_train_smooth', 'tiny.json', 9, 32, 12, False, 2, 12, True),
    ('tiny_eval', 'tiny.json', 8, 32, 12, True, 1, 13, False),
    ('tiny_ragged', 'tiny.json', 3, 19, 5, True, 3, 14, True),     # odd T/R: padding paths
    ('full_eval_b8', 'vilbert.json', 8, 124, 44, True, 1, 21, False),   # BASELINE configs[0]
    ('full_train_b4', 'vilbert.json', 4, 124, 44, True, 1, 22, True),
    # same shapes at a better-conditioned operating point (reference-scale weights): tighter tolerances apply
    ('full_eval_b8_mild', 'vilbert.json', 8, 124, 44, True, 1, 21, False, 'mild'),
    ('full_train_b4_mild', 'vilbert.json', 4, 124, 44, True, 1, 22, True, 'mild'),
]


def sample_idx(name: str, numel: int, n: int = 8) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randint(0, numel, (n,), generator=g)


def grad_summary(named_grads) -> dict:
    out = {}
    for k, gr in named_grads.items():
        f = gr.detach().double().flatten()
        out[k] = dict(norm=float(f.norm()), sum=float(f.sum()), samples=f[sample_idx(k, f.numel())].float().clone())
    return out


def run_case(name, cfg_file, B, T, R, l1, wseed, bseed, train, style='trained'):
    cfg_path = os.path.join(ROOT, 'cqa_crct_b200', 'config', cfg_file)
    cfg = ModelConfig(cfg_path)
    params = default_params(cfg_path, max_seq_len=T, max_vis_features=R, L1=l1)
    enc = ref_shim.RefEncoder(cfg_path, params, seed=0)
    m = enc.module
    m.bert_pretrained.load_state_dict(synth_state_dict(cfg, params['categories'], wseed, style), strict=True)
    m.eval()                      # dropout off; the train/eval BRANCH is chosen by kwargs (encoder_decorator.py:31-32)
    batch = make_batch(B, T, R, cfg.v_feature_size, seed=bseed, vocab_size=cfg.vocab_size)
    rec = dict(name=name, config=cfg_file, B=B, T=T, R=R, l1=l1, weight_seed=wseed, batch_seed=bseed,
               weight_style=style, train=train, torch=torch.__version__)
    if train:
        loss, _, nsp, _, scores, reg, _ = enc.glue_forward(m, batch, params)
        loss.backward()
        rec.update(loss=float(loss), nsp_loss=float(nsp))
        rec['grads'] = grad_summary({k: p.grad for k, p in m.bert_pretrained.named_parameters() if p.grad is not None})
    else:
        with torch.no_grad():
            _, _, _, _, scores, reg = enc.glue_forward(m, batch, params, evaluation=True)
    rec.update(logits=scores.detach().clone(), reg_pred=reg[0].detach().clone(), reg_loss=reg[1].detach().clone(),
               reg_l1=reg[2].detach().clone(), reg_right=tuple(int(x) for x in reg[3]), reg_dist=reg[4].detach().clone())
    return rec


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        rec = run_case(*case)
        path = os.path.join(out_dir, case[0] + '.pt')
        torch.save(rec, path)
        print(case[0], 'logits[0]=', rec['logits'][0].tolist(), 'loss=', rec.get('loss'), os.path.getsize(path), 'bytes')


if __name__ == '__main__':
    main()
