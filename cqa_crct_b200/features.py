"""f3, second half — reader of the Detector's on-disk RoI feature records.

The detection stage (out of scope) writes `.npy` chunks: `np.save(path, feature_lst)` with one dict per figure
(reference: Detector/extract_features.py:567-575):
    {"image_id": int, "vis_feat": float [n, 1024] (box_head features, row 0 = the whole-figure `<IMG>` box),
     "vis_bbox": float [n, 4 or 5] (axis-normalised x0,y0,x1,y1 [, legend id]), "class": int [n] (row 0 = 1000),
     "text_feat": {...}, "width": int, "height": int}
and the loader turns one record into the model's visual inputs (CRCT/fig_dataloader.py:308-361 `encode_and_reshape_img`
+ CRCT/utils.py:174-225 `encode_image_input`): the `<IMG>` row loses its box and gets class id `categories`, everything is
cut / zero-padded to `max_vis_features` regions, `image_mask` marks the real ones.  This module does exactly that for the
PlotQA configuration at evaluation time (`mask_prob_img = 0`: no feature masking, `image_label` carries no information) and
assembles the question-level visual half of an evaluation batch for `cqa_crct_b200.evaluate.evaluate_batch`.  The text half
(tokenisation, CRCT/fig_dataloader.py:524-690) needs the BERT vocabulary and stays with the caller.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import numpy as np
import torch

IMG_TOKEN_FEATURES_CLASS = 1000          # fig_dataloader.py:55; written by Detector/extract_features.py:74


def load_feature_chunk(path: str) -> Dict[int, dict]:
    """One `.npy` chunk -> {image_id: record} (fig_dataloader.py:143-150 loads the same file with allow_pickle=True)."""
    arr = np.load(path, allow_pickle=True)
    out = {}
    for rec in list(arr):
        for key in ('image_id', 'vis_feat', 'vis_bbox', 'class'):
            if key not in rec:
                raise ValueError(f'{path}: record without "{key}" (Detector/extract_features.py:567-575 layout expected)')
        out[int(rec['image_id'])] = rec
    return out


def encode_regions(rec: dict, max_regions: int, categories: int) -> Dict[str, torch.Tensor]:
    """fig_dataloader.py:308-361 + utils.py:174-225 for one figure (PlotQA, evaluation).  Does not modify `rec`."""
    cls = np.asarray(rec['class']).astype(np.int64).copy()
    if cls.shape[0] == 0 or int(cls[0]) != IMG_TOKEN_FEATURES_CLASS:
        raise ValueError(f'record {rec.get("image_id")}: row 0 must be the <IMG> box (class {IMG_TOKEN_FEATURES_CLASS})')   # :323
    boxes = np.asarray(rec['vis_bbox'], dtype=np.float64).copy()
    feats = np.asarray(rec['vis_feat'], dtype=np.float64)
    if feats.shape[0] != boxes.shape[0] or feats.shape[0] != cls.shape[0]:
        raise ValueError(f'record {rec.get("image_id")}: {feats.shape[0]} features, {boxes.shape[0]} boxes, {cls.shape[0]} classes')
    boxes[0, :4] = 0                                        # <IMG> token doesn't need location, :314
    cls[0] = categories                                     # :342
    boxes = boxes[:, :4]                                    # :346 (a 5th column is the legend id)
    n = min(int(boxes.shape[0]), max_regions)               # utils.py:176
    feat_pad = np.zeros((max_regions, feats.shape[-1]))
    box_pad = np.zeros((max_regions, 4))
    cls_pad = np.zeros((max_regions,), dtype=np.int64)
    feat_pad[:n], box_pad[:n], cls_pad[:n] = feats[:n], boxes[:n], cls[:n]
    mask = torch.zeros(max_regions, dtype=torch.int64)
    mask[:n] = 1                                            # utils.py:209-212
    return {'image_feat': torch.tensor(feat_pad).float(), 'image_loc': torch.tensor(box_pad).float(), 'image_mask': mask,
            'image_target': torch.tensor(cls_pad), 'image_label': torch.full((max_regions,), -1, dtype=torch.int64)}


def visual_batch(records: Sequence[dict], max_regions: int, categories: int) -> Dict[str, torch.Tensor]:
    """Question-level visual tensors [Q, R, ...] for `evaluate_batch` (one row per question; candidates share it)."""
    enc = [encode_regions(r, max_regions, categories) for r in records]
    return {k: torch.stack([e[k] for e in enc]) for k in enc[0]}


def question_batch(text: Dict[str, torch.Tensor], records: Sequence[dict], R: torch.Tensor, num_ans: torch.Tensor, gt_id: torch.Tensor,
                   needs_reg: torch.Tensor, tolerance_margin: torch.Tensor, params: dict) -> Dict[str, torch.Tensor]:
    """The de-duplicated evaluation batch (`cqa_crct_b200.evaluate`): per-candidate text tensors from the caller's
    tokeniser + the per-question visual tensors read from the Detector's records."""
    out = dict(text)
    out.update(visual_batch(records, params['max_vis_features'], params['categories']))
    out.update({'R': R, 'num_ans': num_ans, 'gt_id': gt_id, 'needs_reg': needs_reg, 'tolerance_margin': tolerance_margin,
                'id': torch.tensor([int(r['image_id']) for r in records])})
    return out


def write_feature_chunk(path: str, records: Iterable[dict]) -> None:
    """Same call the Detector makes (extract_features.py: `np.save(out, feature_lst)`); used by the tests and the demo."""
    np.save(path, np.array(list(records), dtype=object), allow_pickle=True)
