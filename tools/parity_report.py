"""Prints the end-to-end parity numbers (CUDA path vs oracle / golden) for every golden case.  GPU only."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward   # noqa: E402
from cqa_crct_b200.synthetic import default_params                     # noqa: E402
from oracle import crct_oracle as O                                    # noqa: E402
from tests.helpers import load_golden, golden_inputs, rel_err          # noqa: E402


def scale_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


def main():
    rows = []
    for name in ['tiny_eval', 'full_eval_b8', 'full_eval_b8_mild', 'tiny_train_l1', 'tiny_train_smooth', 'tiny_ragged', 'full_train_b4', 'full_train_b4_mild']:
        rec = load_golden(name)
        cfg_path, cfg, sd, batch = golden_inputs(rec)
        params = default_params(cfg_path, device='cuda', max_seq_len=rec['T'], max_vis_features=rec['R'], L1=rec['l1'])
        m = VisualDialogEncoder(params)
        m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
        m.to('cuda').eval()
        gb = {k: v.to('cuda') for k, v in batch.items()}
        r = dict(case=name)
        if rec['train']:
            m.zero_grad()
            loss, _, nsp, _, scores, reg, _ = glue_forward(m, gb, params)
            loss.backward()
            torch.cuda.synchronize()
            r['loss_abs_err'] = abs(float(loss) - rec['loss'])
            out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=rec['l1'], dtype=torch.float64)
            g = O.backward(cache)
            with O.bf16_emulation():
                oute, cachee = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=rec['l1'], dtype=torch.float64)
                ge = O.backward(cachee)
            r['emu_logits_err'] = scale_err(scores, oute['logits'])
            pe = []
            num = den = 0.0
            gn = max(float(v.norm()) for v in ge.values())
            namedp = dict(m.bert_pretrained.named_parameters())
            for k, ref in ge.items():
                if float(ref.norm()) < 1e-6 * gn:
                    continue
                got = namedp[k].grad.double().cpu()
                pe.append((rel_err(got, ref), k))
                num += float((got - ref).norm() ** 2); den += float(ref.norm() ** 2)
            pe.sort(reverse=True)
            r['emu_grad_global_rel'] = (num / den) ** 0.5
            r['emu_grad_median_rel'] = pe[len(pe) // 2][0]
            r['emu_grad_top5'] = [(round(e, 4), k) for e, k in pe[:5]]
            named = dict(m.bert_pretrained.named_parameters())
            gnorm = max(float(v.norm()) for v in g.values())
            worst_e, worst_c, worst_k = 0.0, 1.0, ''
            tot_num = tot_den = 0.0
            per = []
            for k, ref in g.items():
                if float(ref.norm()) < 1e-6 * gnorm:
                    continue
                got = named[k].grad.double().cpu()
                e = rel_err(got, ref)
                c = float(torch.nn.functional.cosine_similarity(got.flatten(), ref.double().flatten(), dim=0))
                per.append((e, c, k, float(ref.norm()) / gnorm, float((got - ref).norm()) / gnorm))
                tot_num += float((got - ref).norm() ** 2); tot_den += float(ref.norm() ** 2)
                if e > worst_e:
                    worst_e, worst_k = e, k
                worst_c = min(worst_c, c)
            per.sort(reverse=True)
            r.update(grad_worst_rel=worst_e, grad_worst_tensor=worst_k, grad_min_cos=worst_c,
                     grad_global_rel=(tot_num / tot_den) ** 0.5, grad_top5=[(round(e, 4), round(c, 5), k, '%.1e' % rn, '%.1e' % en) for e, c, k, rn, en in per[:5]],
                     grad_worst_abs_over_gnorm=max(x[4] for x in per), grad_worst_rel_big=max(x[0] for x in per if x[3] > 1e-2),
                     grad_min_cos_big=min(x[1] for x in per if x[3] > 1e-2),
                     grad_median_rel=per[len(per) // 2][0])
        else:
            with torch.no_grad():
                _, _, _, _, scores, reg = glue_forward(m, gb, params, evaluation=True)
            with O.bf16_emulation():
                oute, _ = O.forward(sd, O.Config(cfg.__dict__), batch, train=False, l1=rec['l1'], keep_cache=False)
            r['emu_logits_err'] = scale_err(scores, oute['logits'])
        r['logits_err_vs_ref'] = scale_err(scores, rec['logits'])
        r['regpred_err_vs_ref'] = scale_err(reg[0], rec['reg_pred'])
        r['regloss_abs_err'] = float((reg[1].cpu() - rec['reg_loss']).abs().max())
        r['argmax_equal'] = bool(torch.equal(scores.cpu().argmax(1), rec['logits'].argmax(1)))
        r['min_margin_over_scale'] = float((rec['logits'][:, 0] - rec['logits'][:, 1]).abs().min() / rec['logits'].abs().max())
        # the same oracle with the tensor-core GEMM weights rounded to bf16: separates operand quantisation from kernel error
        sdq = {k: (v.bfloat16().float() if (k.startswith('bert.encoder') or 'new_image_embeddings' in k) and k.endswith('weight') and v.dim() == 2 else v)
               for k, v in sd.items()}
        oq, _ = O.forward(sdq, O.Config(cfg.__dict__), batch, train=rec['train'], l1=rec['l1'], keep_cache=False)
        r['logits_err_vs_bf16w_oracle'] = scale_err(scores, oq['logits'])
        r['bf16w_oracle_vs_ref'] = scale_err(oq['logits'], rec['logits'])
        print(json.dumps(r), flush=True)
        rows.append(r)
        del m
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, 'gpurun_out', 'parity_report.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
