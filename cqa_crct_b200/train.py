"""Training entry point with the reference's flags, loop structure and checkpoint layout (CRCT/train.py:27-300).

    python -m cqa_crct_b200.train -model_config cqa_crct_b200/config/vilbert.json -save_path /tmp/run -batch_size 80 \\
        -num_epochs 1 -iters_per_epoch 50 [-start_checkpoint plotqa_encoder_0_50.ckpt [-continue]] [-graph]
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m cqa_crct_b200.train -ddp ...       # one process per GPU

What is the same as the reference: flag names and defaults (CRCT/options.py:9-81), model / optimizer / scheduler
construction order (train.py:80-87), weights-only vs `-continue` checkpoint loading (:89-127), the per-iteration body —
forward glue, loss / batch_multiply, backward, optimizer step, zero_grad, scheduler step (:167-215) — the nine-number
statistics vector summed across ranks (:178-189), the per-epoch checkpoint (:282-291).
What is different: the PlotQA dataset reader is out of scope (detection stage), so batches come from the seeded
synthetic generator with the dataset's tensor layout (`SyntheticPlotQA`); bf16 needs no autocast / GradScaler
(:157,172,208-214); statistics stay on the device and are read once every PRINT_EVERY iterations instead of six
`.item()` syncs per iteration (:174-183); `-graph` replays the whole step as CUDA graphs (`graph.GraphedTrainStep`).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from typing import Dict, Iterator, List, Optional

import torch
import torch.distributed as dist

from . import checkpoint as ckpt
from .encoder import VisualDialogEncoder, glue_forward
from .optim import FusedAdamW, WarmupLinearScheduleNonZero
from .parallel import DistributedDataParallel
from .synthetic import make_batch

_HERE = os.path.dirname(os.path.abspath(__file__))
PRINT_EVERY = 100                                      # train.py:151


def read_command_line(argv: Optional[List[str]] = None) -> dict:
    """The flags of CRCT/options.py:9-81 that reach the training path, same names and defaults; the dataset-file flags
    (-figure_feat_path, -qa_parent_dir, -dataset_config, ...) are replaced by the synthetic-data ones at the end."""
    p = argparse.ArgumentParser(description='CRCT question-answering stage on B200')
    p.add_argument('-start_checkpoint', default='')
    p.add_argument('-model_config', default=os.path.join(_HERE, 'config', 'vilbert.json'))
    p.add_argument('-num_workers', default=16, type=int)
    p.add_argument('-batch_size', default=80, type=int)
    p.add_argument('-num_epochs', default=20, type=int)
    p.add_argument('-batch_multiply', default=1, type=int)
    p.add_argument('-lr', default=2e-5, type=float)
    p.add_argument('-image_lr', default=2e-5, type=float)
    p.add_argument('-min_lr', default=1.3e-5, type=float)
    p.add_argument('-continue', action='store_true')
    p.add_argument('-max_seq_len', default=124, type=int)
    p.add_argument('-nsp_loss_coeff', default=1, type=float)
    p.add_argument('-reg_loss_coeff', default=1, type=float)
    p.add_argument('-L1', action='store_true')
    p.add_argument('-mask_prob_img', default=0, type=float)
    p.add_argument('-save_path', default='')
    p.add_argument('-save_name', default='')
    p.add_argument('-cuda_num', default=-1, type=int)
    p.add_argument('-eval_batch_size', default=512, type=int)
    p.add_argument('-ddp', action='store_true')
    p.add_argument('-rank_from', type=int, default=0)
    p.add_argument('-seed', type=int, default=0)
    p.add_argument('-qa_file', default='synthetic')
    p.add_argument('-no_eval', action='store_true')
    p.add_argument('-wd', default=0.01, type=float)
    p.add_argument('-tol_margin', default=0.01, type=float)
    p.add_argument('-warmup', default=3000, type=int)
    p.add_argument('-dataset', type=str, default='plotqa')
    p.add_argument('-categories', type=int, default=228)
    p.add_argument('-CE_REG', action='store_true')
    p.add_argument('-BOT_MODE', action='store_true')
    p.add_argument('-binary_answers', type=lambda x: str(x).lower() == 'true', default=False)
    p.add_argument('-eval_set', type=str, default='val')
    # synthetic data (replaces the PlotQA files) and launch mode
    p.add_argument('-max_vis_features', default=44, type=int)
    p.add_argument('-iters_per_epoch', default=100, type=int, help='synthetic batches per epoch and rank')
    p.add_argument('-eval_questions', default=64, type=int, help='synthetic questions in the post-epoch evaluation')
    p.add_argument('-graph', action='store_true', help='replay the step as CUDA graphs (graph.GraphedTrainStep)')
    parsed = vars(p.parse_args(args=argv))
    if parsed['save_name']:
        parsed['save_path'] = os.path.join(parsed['save_path'], parsed['save_name'])          # options.py:94-97
    parsed['rank'] = int(os.environ.get('RANK', 0))
    parsed['world_size'] = int(os.environ.get('WORLD_SIZE', 1)) if parsed['ddp'] else 1
    parsed['num_proc'] = parsed['world_size']
    parsed['dvqa_floats'] = []
    parsed['log_file'] = None
    return parsed


class SyntheticPlotQA:
    """Stands in for `DataLoader(PlotQA_Dataset(...), sampler=DistributedSampler(...), drop_last=True)` (train.py:46-75):
    `len()` batches per epoch, every (epoch, rank, index) a distinct seeded batch of the dataset's layout."""

    def __init__(self, params: dict, iters: int, feat_dim: int, vocab_size: int):
        self.p, self.iters, self.feat_dim, self.vocab = params, iters, feat_dim, vocab_size
        self.epoch = 0

    def set_epoch(self, epoch: int):                       # DistributedSampler.set_epoch, train.py:161-162
        self.epoch = epoch

    def __len__(self):
        return self.iters

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        p = self.p
        for i in range(self.iters):
            seed = ((p['seed'] * 1009 + self.epoch) * 100003 + i) * 64 + p['rank']
            yield make_batch(p['batch_size'], p['max_seq_len'], p['max_vis_features'], self.feat_dim, seed=seed & 0x7FFFFFFF,
                             vocab_size=self.vocab, categories=p['categories'])


def log_line(params, line, all_ranks=False):               # utils.py:42-47
    if params['rank'] == 0 or all_ranks:
        if params.get('log_file'):
            with open(params['log_file'], 'a') as f:
                f.write(line + '\n')
        print(line, flush=True)


def train(gpu: int, params: dict) -> dict:
    """CRCT/train.py:27-300 `train(gpu, params)`.  Returns {'iter_id', 'checkpoints', 'loss_history'} for callers/tests."""
    if not torch.cuda.is_available():
        raise SystemExit('cqa_crct_b200.train needs a B200; there is no CPU fallback')
    torch.cuda.set_device(gpu)
    params['device'] = device = torch.device('cuda', gpu if params['cuda_num'] < 0 else params['cuda_num'])
    if params['ddp'] and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=device)                                   # train.py:29-33
    world = params['world_size']
    torch.manual_seed(params['seed'])
    log_line(params, 'De facto batch_size: {}*{}*{} = {}'.format(params['batch_size'], world, params['batch_multiply'],
                                                               params['batch_size'] * world * params['batch_multiply']), all_ranks=True)
    crct_model = VisualDialogEncoder(params)                                                # train.py:80-81
    crct_model.to(device)
    cfg = crct_model.cfg
    dataloader = SyntheticPlotQA(params, params['iters_per_epoch'], cfg.v_feature_size, cfg.vocab_size)
    iters_per_epoch = len(dataloader) / params['batch_multiply']                            # train.py:77
    optimizer = FusedAdamW(crct_model, lr=params['lr'], image_lr=params['image_lr'], weight_decay=params['wd'])   # get_optimizer, :85
    scheduler = WarmupLinearScheduleNonZero(optimizer, warmup_steps=params['warmup'], min_lr=params['min_lr'],
                                            t_total=int(iters_per_epoch * 20))                # :87

    start_iter_id, cont_epoch = 0, 0
    if params['start_checkpoint']:                                                           # :91-127
        if not params['continue']:
            n = ckpt.load_weights(crct_model, params['start_checkpoint'], device)
            log_line(params, f'number of keys transferred {n}')
        else:
            cont_epoch, start_iter_id, _ = ckpt.resume(crct_model, optimizer, scheduler, params['start_checkpoint'], device)
        log_line(params, 'Current epoch: {}'.format(cont_epoch))

    model = DistributedDataParallel(crct_model) if params['ddp'] and world > 1 else crct_model     # :139-142
    enc = crct_model
    gstep = None
    if params['graph']:
        if params['batch_multiply'] != 1:
            raise ValueError('-graph captures forward + backward + optimizer step as one unit: use -batch_multiply 1')
        from .graph import GraphedTrainStep
        example = {k: v.to(device) for k, v in next(iter(dataloader)).items()}
        snapshot = (enc.arena.w32.clone(), optimizer.m.clone(), optimizer.v.clone(), optimizer.step_count, scheduler.last_epoch)
        gstep = GraphedTrainStep(model, optimizer, params, example, scheduler=scheduler, warmup_steps=2)
        # the capture warm-up ran real steps on the example batch: restore the pre-capture state
        enc.arena.w32.copy_(snapshot[0]); optimizer.m.copy_(snapshot[1]); optimizer.v.copy_(snapshot[2])
        optimizer.moments_partial = False              # restored in full on every rank
        optimizer.step_count, scheduler.last_epoch = snapshot[3], snapshot[4]
        optimizer.lr_factor = scheduler.factor(scheduler.last_epoch)
        enc.arena.refresh_bf16()

    optimizer.zero_grad()                                                                    # :148
    stats = torch.zeros(9, dtype=torch.float64, device=device)       # running sums of the 9-number vector of train.py:181-183
    seen = 0
    history, written = [], []
    num_step_iterations = 0
    start_t = time.time()
    log_line(params, 'Starting iterations...')
    for epoch_id in range(params['num_epochs']):
        dataloader.set_epoch(epoch_id)
        step_iter_id = start_iter_id + num_step_iterations
        for iter_id, batch in enumerate(dataloader):
            step_iter_id = start_iter_id + num_step_iterations
            enc.train()                                                                      # :169
            if gstep is not None:
                loss = gstep.step(batch)
                sc = enc.last_scalars                                  # {loss, nsp, mean reg, #+-5 %, #tol} on the device
                vec = torch.stack([loss.reshape(()), sc[1], sc[2], sc[3], sc[4]]).double()
                num_step_iterations += 1
            else:
                gb = {k: v.to(device, non_blocking=True) for k, v in batch.items()}
                loss, lm_loss, nsp_loss, img_loss, nsp_scores, regression, legend_loss = glue_forward(model, gb, params)   # :173
                vec = torch.stack([loss.detach().reshape(()), nsp_loss.detach().reshape(()), regression[1].detach().mean(),
                                   regression[3][0].reshape(()), regression[3][1].reshape(())]).double()
                if params['batch_multiply'] > 1:
                    loss = loss / params['batch_multiply']                                   # :205-206
                loss.backward()                                                              # :208
                if iter_id % params['batch_multiply'] == 0:                                  # :210-215
                    optimizer.step()
                    optimizer.zero_grad()
                    scheduler.step()
                    num_step_iterations += 1
            n_reg = float(batch['needs_reg'].sum())                    # host tensor: no device sync (train.py:170)
            # train.py:174-179: the logged regression loss is the mean over the rows that NEED regression (0 when there are
            # none); regression[1] is dense with zeros elsewhere, so mean_B * B / n_reg is that mean
            vec[2] *= vec.new_tensor(batch['needs_reg'].numel() / max(n_reg, 1.0))
            stats[:5] += vec
            stats[5] += n_reg
            seen += 1
            if (iter_id + 1) % PRINT_EVERY == 0 or iter_id + 1 == len(dataloader):           # :226-279, read once per window
                s = stats.clone()
                if params['ddp'] and world > 1:
                    dist.all_reduce(s, op=dist.ReduceOp.SUM)                                 # :184-189
                    s[:3] /= world
                s = s.cpu()
                n_reg = max(float(s[5]), 1.0)
                rec = {'epoch': cont_epoch + epoch_id, 'iter': step_iter_id + 1, 'loss': float(s[0]) / seen, 'nsp': float(s[1]) / seen,
                       'reg': float(s[2]) / seen, 'reg_5_acc': float(s[3]) / n_reg, 'reg_t_acc': float(s[4]) / n_reg,
                       'lr': optimizer.current_lrs()[0], 'sec': time.time() - start_t}
                history.append(rec)
                log_line(params, '[Ep: {epoch}][Iter: {iter}][loss: {loss:.4f}][nsp: {nsp:.4f}][reg: {reg:.4f}][reg_5_acc: {reg_5_acc:.3f}]'
                                 '[reg_t_acc: {reg_t_acc:.3f}][lr: {lr:.3g}][{sec:.1f}s]'.format(**rec))
                stats.zero_()
                seen = 0
                start_t = time.time()
        # ---- end of epoch: checkpoint (train.py:282-291), then evaluation on the (synthetic) validation questions
        step_iter_id = start_iter_id + num_step_iterations
        if gstep is not None and params['save_path']:
            gstep.consolidate_optimizer_state()        # sharded optimizer: the Adam moments live on their owner ranks (collective)
        if params['rank'] == 0 and params['save_path']:
            path = ckpt.save_checkpoint(params['save_path'], cont_epoch + epoch_id, step_iter_id, crct_model, optimizer, scheduler)
            written.append(path)
            log_line(params, '     --> Saving model as: {}'.format(path))
        if not params['no_eval']:
            from .evaluation import evaluate_synthetic
            table = evaluate_synthetic(model, params, n_questions=params['eval_questions'])
            enc.train()
            log_line(params, 'Eval accuracy (nsp / total+-5% / total tol): {:.3f} / {:.3f} / {:.3f}'.format(
                *[float(table[r, 0] / max(float(table[r, 1]), 1.0)) for r in (0, 4, 5)]))
    if params['ddp'] and dist.is_initialized() and params.get('_own_pg', True):
        dist.barrier()
    return {'iter_id': start_iter_id + num_step_iterations, 'checkpoints': written, 'loss_history': history}


def main(argv: Optional[List[str]] = None):
    params = read_command_line(argv)
    return train(int(os.environ.get('LOCAL_RANK', 0)), params)


if __name__ == '__main__':
    main(sys.argv[1:])
