"""CPU experiment (test infrastructure): which bf16 STORAGE points of the CUDA path account for its distance from the fp32
reference?  Runs the oracle's bf16 emulation (oracle/crct_oracle.py: `_q` at every tensor the kernels store in bf16) on a
golden case with groups of rounding points switched off one at a time, and prints logits / gradient error against the
un-rounded oracle.  Rounding points are identified by the oracle source line of the `_q` call.

    python tools/parity_sensitivity.py [case] [group ...]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import crct_oracle as O                        # noqa: E402
from tests.helpers import load_golden, golden_inputs      # noqa: E402

SRC = open(O.__file__).read().split('\n')


def lines_with(*needles):
    return {i + 1 for i, l in enumerate(SRC) if any(n in l for n in needles) and ('_q(' in l or '_qw(' in l) and not l.startswith('def ')}


GROUPS = {
    'weights': lines_with('_qw(W)'),
    'a_operand': lines_with('return _q(x) @', '_q(x2)'),
    'qkv': lines_with('qkv = _q(', 'qkv1, qkv2 = _q'),
    'probs_ctx': lines_with('ctx = (_q(pd)', 'return _q(ctx)', 'dv = _q(pd)'),
    'h_gelu': lines_with('h = _q(gelu'),
    'bwd_dz': lines_with('dz = _q(dz)', 'dz1 = _q(dz1)', 'dzv, dzt = _q'),
    'bwd_du_dgelu': lines_with('du = _q('),
    'bwd_res': lines_with('return _q(da + dz)', 'return _q(dx + dz1)', 'return _q(dv_in'),
    'bwd_attn': lines_with('ds = _q(', 'merge = lambda', 'dctx = _q(', 'dctx1, dctx2 = _q'),
}
FWD = ['weights', 'a_operand', 'qkv', 'probs_ctx', 'h_gelu']
BWD = ['bwd_dz', 'bwd_du_dgelu', 'bwd_res', 'bwd_attn']


def run(sd, cfg, batch, l1, off_lines, dtype=torch.float64):
    import sys as _s

    def q(x):
        if _s._getframe(1).f_lineno in off_lines:
            return x
        return x.to(torch.bfloat16).to(x.dtype)

    O._q, O._qw = q, q
    try:
        out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=l1, dtype=dtype)
        g = O.backward(cache)
    finally:
        ident = lambda x: x
        O._q, O._qw = ident, ident
    return out['logits'], g


def errs(logits, g, ref_logits, ref_g):
    le = float((logits - ref_logits).abs().max() / ref_logits.abs().max())
    num = sum(float((g[k] - v).norm() ** 2) for k, v in ref_g.items())
    den = sum(float(v.norm() ** 2) for v in ref_g.values())
    return le, (num / den) ** 0.5


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else 'full_train_b4_mild'
    rec = load_golden(case)
    cfg_path, cfg, sd, batch = golden_inputs(rec)
    torch.set_num_threads(os.cpu_count())
    all_lines = set().union(*GROUPS.values())
    ref_logits, ref_g = run(sd, cfg, batch, rec['l1'], all_lines | set(range(1, 10000)))
    rows = {}
    configs = {'all_rounded': set(), 'fwd_only_rounded': set().union(*[GROUPS[k] for k in BWD]),
               'bwd_only_rounded': set().union(*[GROUPS[k] for k in FWD]), 'weights_only': all_lines - GROUPS['weights']}
    for k in GROUPS:
        configs['all_but_' + k] = GROUPS[k]
    want = sys.argv[2:]
    for name, off in configs.items():
        if want and name not in want:
            continue
        logits, g = run(sd, cfg, batch, rec['l1'], off)
        rows[name] = errs(logits, g, ref_logits, ref_g)
        print(f'{name:28s} logits {rows[name][0]:.3e}  grads {rows[name][1]:.3e}', flush=True)
    json.dump(rows, open('/tmp/parity_sensitivity.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
