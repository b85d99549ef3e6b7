"""Shared helpers for the parity tests (test infrastructure: may import oracle/)."""
import os
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
CONFIG_DIR = os.path.join(ROOT, 'cqa_crct_b200', 'config')


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + '.pt'), weights_only=False)


def golden_inputs(rec):
    """Regenerates the exact weights / batch a golden fixture was produced with."""
    from cqa_crct_b200.spec import ModelConfig, synth_state_dict
    from cqa_crct_b200.synthetic import make_batch
    cfg_path = os.path.join(CONFIG_DIR, rec['config'])
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, rec['weight_seed'], rec['weight_style'])
    batch = make_batch(rec['B'], rec['T'], rec['R'], cfg.v_feature_size, seed=rec['batch_seed'], vocab_size=cfg.vocab_size)
    return cfg_path, cfg, sd, batch


def sample_idx(name, numel, n=8):
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randint(0, numel, (n,), generator=g)


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))
