"""f3 (second half): the reader of the Detector's `.npy` feature records against the reference's own loader code.
The reference functions are executed where they lie when /root/reference exists (build container); a literal expected
layout is checked everywhere."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from cqa_crct_b200 import features as F


def _records(seed=0, n_fig=5, feat=32):
    g = np.random.RandomState(seed)
    recs = []
    for i in range(n_fig):
        n = int(g.randint(2, 60))
        cls = g.randint(8, 228, size=n)
        cls[0] = F.IMG_TOKEN_FEATURES_CLASS
        box = g.rand(n, 5 if i % 2 else 4) * 1.2 - 0.1
        recs.append({'image_id': 100 + i, 'vis_feat': np.maximum(g.randn(n, feat), 0).astype(np.float32), 'vis_bbox': box,
                     'class': cls, 'text_feat': {}, 'width': 640, 'height': 480})
    return recs


def test_chunk_round_trip_and_layout(tmp_path):
    recs = _records()
    path = str(tmp_path / 'features_0.npy')
    F.write_feature_chunk(path, recs)
    chunk = F.load_feature_chunk(path)
    assert sorted(chunk) == [100, 101, 102, 103, 104]
    R, C = 44, 228
    for rec in recs:
        e = F.encode_regions(chunk[rec['image_id']], R, C)
        n = min(len(rec['class']), R)
        assert e['image_feat'].shape == (R, 32) and e['image_loc'].shape == (R, 4)
        assert int(e['image_mask'].sum()) == n and bool((e['image_mask'][:n] == 1).all())
        assert int(e['image_target'][0]) == C and torch.equal(e['image_target'][1:n], torch.tensor(rec['class'][1:n]))
        assert float(e['image_loc'][0].abs().sum()) == 0.0
        assert torch.allclose(e['image_loc'][1:n], torch.tensor(rec['vis_bbox'][1:n, :4]).float())
        assert torch.equal(e['image_feat'][:n], torch.tensor(rec['vis_feat'][:n]))
        assert float(e['image_feat'][n:].abs().sum()) == 0.0 and int(e['image_target'][n:].abs().sum()) == 0
        assert int(rec['class'][0]) == F.IMG_TOKEN_FEATURES_CLASS            # the record itself is not modified
    vb = F.visual_batch([chunk[k] for k in sorted(chunk)], R, C)
    assert vb['image_feat'].shape == (5, R, 32) and vb['image_mask'].dtype == torch.int64


def test_rejects_malformed_records(tmp_path):
    recs = _records(n_fig=1)
    recs[0]['class'][0] = 7
    with pytest.raises(ValueError):
        F.encode_regions(recs[0], 44, 228)
    bad = dict(_records(n_fig=1)[0])
    del bad['vis_bbox']
    path = str(tmp_path / 'bad.npy')
    F.write_feature_chunk(path, [bad])
    with pytest.raises(ValueError):
        F.load_feature_chunk(path)


@pytest.mark.skipif(not os.path.isfile('/root/reference/CRCT/utils.py'), reason='reference checkout not present')
def test_matches_the_reference_loader_code():
    """CRCT/utils.py:174-225 `encode_image_input` imported from the reference + the body of
    CRCT/fig_dataloader.py:308-361 executed with a stand-in `self` (the class itself needs the BERT tokenizer files)."""
    sys.path.insert(0, '/root/reference/CRCT')
    try:
        import importlib
        ref_utils = importlib.import_module('utils')
        src = open('/root/reference/CRCT/fig_dataloader.py').read()
        start = src.index('    def encode_and_reshape_img(self, fig_feat):')
        end = src.index('    def encode_and_reshape(self, utterances')
        ns = {'torch': torch, 'np': np, 'encode_image_input': ref_utils.encode_image_input}
        exec('class _Stub:\n' + src[start:end], ns)
        stub = ns['_Stub']()
        stub.IMG_TOKEN_FEATURES_CLASS, stub._split, stub._max_region_num = 1000, 'val', 44
        stub.params = {'dataset': 'plotqa', 'categories': 228, 'mask_prob_img': 0}
        for rec in _records(seed=3, n_fig=6):
            mine = F.encode_regions(rec, 44, 228)
            import copy
            feats, spatials, mask, target, label, _ = stub.encode_and_reshape_img(copy.deepcopy(rec))
            assert torch.equal(mine['image_feat'], feats) and torch.equal(mine['image_loc'], spatials)
            assert torch.equal(mine['image_mask'].float(), mask) and torch.equal(mine['image_target'], target)
    finally:
        sys.path.remove('/root/reference/CRCT')
        sys.modules.pop('utils', None)
