"""TEST INFRASTRUCTURE ONLY — the fair yardstick for the bf16 production path.

    python oracle/make_yardstick.py         # writes tests/golden/yardstick_bf16.json  (needs /root/reference)

Runs the UNMODIFIED reference (oracle/ref_shim.py) twice on the golden cases' seeded weights / inputs:
once in fp32 (what tests/golden/*.pt hold) and once under `torch.autocast('cpu', dtype=torch.bfloat16)` — the
reference's own mixed-precision recipe (CRCT/train.py:172 wraps the step in `torch.cuda.amp.autocast()`; bf16 is the
same width as its fp16) — and records how far the reference's OWN reduced-precision run is from its fp32 run:
class logits (max abs error / max |logit|), regression output, and the parameter gradients (global relative L2 error
over all tensors, per-tensor median / worst).  tests/test_model_gpu.py asserts that the CUDA path's error against the
fp32 reference does not exceed this yardstick (x a stated factor) and an absolute bar.
"""
from __future__ import annotations

import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim                                    # noqa: E402
from oracle.make_golden import CASES                           # noqa: E402
from cqa_crct_b200.spec import ModelConfig, synth_state_dict   # noqa: E402
from cqa_crct_b200.synthetic import make_batch, default_params  # noqa: E402


def scale_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def run(case, autocast: bool):
    name, cfg_file, B, T, R, l1, wseed, bseed, train = case[:9]
    style = case[9] if len(case) > 9 else 'trained'
    cfg_path = os.path.join(ROOT, 'cqa_crct_b200', 'config', cfg_file)
    cfg = ModelConfig(cfg_path)
    params = default_params(cfg_path, max_seq_len=T, max_vis_features=R, L1=l1)
    enc = ref_shim.RefEncoder(cfg_path, params, seed=0)
    m = enc.module
    m.bert_pretrained.load_state_dict(synth_state_dict(cfg, params['categories'], wseed, style), strict=True)
    m.eval()
    batch = make_batch(B, T, R, cfg.v_feature_size, seed=bseed, vocab_size=cfg.vocab_size)
    ctx = torch.autocast('cpu', dtype=torch.bfloat16) if autocast else torch.autocast('cpu', enabled=False)
    grads = None
    if train:
        with ctx:
            loss, _, nsp, _, scores, reg, _ = enc.glue_forward(m, batch, params)
        loss.backward()
        grads = {k: p.grad.detach().double().clone() for k, p in m.bert_pretrained.named_parameters() if p.grad is not None}
    else:
        with torch.no_grad(), ctx:
            _, _, _, _, scores, reg = enc.glue_forward(m, batch, params, evaluation=True)
    return scores.detach().float(), reg[0].detach().float(), grads


def main():
    out = {}
    only = sys.argv[1:]
    for case in CASES:
        if only and case[0] not in only:
            continue
        s32, r32, g32 = run(case, False)
        s16, r16, g16 = run(case, True)
        rec = {'logits_err': scale_err(s16, s32), 'reg_err': scale_err(r16, r32),
               'argmax_equal': bool(torch.equal(s16.argmax(1), s32.argmax(1)))}
        if g32 is not None:
            num = den = 0.0
            per = []
            gn = max(float(v.norm()) for v in g32.values())
            for k, ref in g32.items():
                e, rn = float((g16[k] - ref).norm()), float(ref.norm())
                num, den = num + e * e, den + rn * rn
                if rn > 1e-6 * gn:
                    per.append(e / rn)
            per.sort()
            rec.update(grad_global_rel=(num / den) ** 0.5, grad_median_rel=per[len(per) // 2], grad_worst_rel=per[-1])
        out[case[0]] = rec
        print(case[0], json.dumps(rec), flush=True)
    path = os.path.join(ROOT, 'tests', 'golden', 'yardstick_bf16.json')
    if only and os.path.exists(path):
        old = json.load(open(path))
        old.update(out)
        out = old
    json.dump({'how': 'reference under torch.autocast(cpu, bfloat16) vs the same reference in fp32 (oracle/make_yardstick.py)',
               'torch': torch.__version__, **out}, open(path, 'w'), indent=1)


if __name__ == '__main__':
    main()
