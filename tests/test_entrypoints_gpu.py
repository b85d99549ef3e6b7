"""train / evaluation entry points end to end on the B200 (tiny config): flags, loop, checkpoint file, resume, evaluation
from the checkpoint — CRCT/train.py:27-300, CRCT/evaluation.py:21-66,126-222."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import evaluation, train                       # noqa: E402
from tests.helpers import CONFIG_DIR                               # noqa: E402

TINY = os.path.join(CONFIG_DIR, 'tiny.json')
COMMON = ['-model_config', TINY, '-batch_size', '6', '-max_seq_len', '32', '-max_vis_features', '12', '-iters_per_epoch', '4',
          '-warmup', '2', '-lr', '1e-3', '-image_lr', '1e-3', '-min_lr', '1e-5', '-L1', '-eval_questions', '16', '-eval_batch_size', '64']


@pytest.mark.parametrize('graph', [False, True])
def test_train_checkpoint_resume_evaluate(tmp_path, graph):
    extra = ['-graph'] if graph else []
    r1 = train.main(COMMON + ['-save_path', str(tmp_path), '-num_epochs', '2'] + extra)
    assert r1['iter_id'] == 8 and [os.path.basename(p) for p in r1['checkpoints']] == ['plotqa_encoder_0_4.ckpt', 'plotqa_encoder_1_8.ckpt']
    assert all(torch.isfinite(torch.tensor(h['loss'])) for h in r1['loss_history'])
    payload = torch.load(r1['checkpoints'][0], weights_only=False)
    assert payload['iter_id'] == 4 and payload['scheduler_state_dict']['last_epoch'] == 4
    assert float(payload['optimizer_state_dict']['state'][0]['step']) == 4
    # resume from the first epoch's file: continues at epoch 1, iteration 4, and writes plotqa_encoder_1_8 again
    r2 = train.main(COMMON + ['-save_path', str(tmp_path / 'again'), '-num_epochs', '1', '-start_checkpoint', r1['checkpoints'][0],
                              '-continue', '-no_eval'] + extra)
    assert r2['iter_id'] == 8 and os.path.basename(r2['checkpoints'][0]) == 'plotqa_encoder_1_8.ckpt'
    a = torch.load(r1['checkpoints'][1], weights_only=False)
    b = torch.load(r2['checkpoints'][0], weights_only=False)
    assert a['scheduler_state_dict']['last_epoch'] == b['scheduler_state_dict']['last_epoch'] == 8
    # (the resumed epoch re-draws epoch 0's batches — `set_epoch(epoch_id)` restarts at 0 on resume in the reference too,
    # train.py:159-162 — so the two epoch-1 files are not expected to hold equal weights)
    # weights-only load + evaluation entry point
    total = evaluation.main(COMMON + ['-start_checkpoint', r1['checkpoints'][1]])
    t = total.cpu()
    assert float(t[0, 1]) == 16 and float(t[4, 1]) == 16 and 0 <= float(t[0, 0]) <= 16
