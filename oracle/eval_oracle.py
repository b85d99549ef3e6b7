"""TEST INFRASTRUCTURE — CPU restatement of the reference's per-question answer selection and accuracy table.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows CRCT/evaluation.py:254-258 (softmax of the two nsp scores, column 0), :287-296 (per-question argmax over the
question's `num_ans` candidates, or gt_id for '_REGS' files; selection of regression[4] / [2] / [0] of that candidate),
:306-313 (correctness flags) and :494-525 (`reduce_total_acc`).  Written as the reference writes it — a Python loop
over questions — on torch CPU tensors.  Pinned by running the reference's own lines on the same arrays in
tests/test_oracle_vs_reference.py where /root/reference is present (the reference has no stored vectors for this path).
"""
import torch
import torch.nn.functional as F


def select_and_score(nsp_scores, reg_pred, reg_dist, reg_l1, num_ans, gt_id, needs_reg, tolerance, force_gt=False):
    nsp_probs = F.softmax(nsp_scores.float(), dim=1)                       # evaluation.py:255
    output = nsp_probs[:, 0]                                                # :258
    total_options = 0
    answers, sel_dist, sel_l1, sel_pred = [], [], [], []
    for i, n in enumerate(num_ans.view(-1).tolist()):                       # :287
        if force_gt:
            ans_id = gt_id.view(-1)[i]                                      # :289
        else:
            ans_id = torch.argmax(output[total_options: total_options + n])   # :291
        answers.append(ans_id)
        sel_dist.append(reg_dist[total_options: total_options + n][ans_id.item()])   # :293
        sel_l1.append(reg_l1[total_options: total_options + n][ans_id.item()])       # :294
        sel_pred.append(reg_pred[total_options: total_options + n][ans_id.item()])   # :295
        total_options += n
    assert total_options == nsp_scores.shape[0]                             # :298
    answers = torch.stack(answers)
    sel_dist, sel_l1, sel_pred = torch.stack(sel_dist), torch.stack(sel_l1), torch.stack(sel_pred)
    nsp_right = answers == gt_id.view(-1)                                   # :305
    needs = needs_reg.view(-1).bool()                                       # :306
    reg_right = (sel_dist <= 0.05) & needs                                  # :307
    reg_t_right = (sel_l1 <= tolerance.view(-1)) & needs                    # :308
    correct = nsp_right & (needs.logical_not() | reg_right)                 # :310
    correct_t = nsp_right & (needs.logical_not() | reg_t_right)             # :311
    total = torch.zeros(6, 2, dtype=torch.float64)                          # :497-517
    total[0, 0], total[0, 1] = nsp_right.sum(), nsp_right.shape[0]
    total[1, 0], total[1, 1] = (nsp_right & needs).sum(), needs.sum()
    total[2, 0], total[2, 1] = reg_right.sum(), needs.sum()
    total[3, 0], total[3, 1] = reg_t_right.sum(), needs.sum()
    total[4, 0], total[4, 1] = correct.sum(), nsp_right.shape[0]
    total[5, 0], total[5, 1] = correct_t.sum(), nsp_right.shape[0]
    return {'answers': answers, 'prob': output, 'reg_output': sel_pred, 'reg_loss': sel_dist, 'reg_t_loss': sel_l1,
            'flags': torch.stack([nsp_right, reg_right, reg_t_right, correct, correct_t], 1).to(torch.uint8), 'total_correct': total}
