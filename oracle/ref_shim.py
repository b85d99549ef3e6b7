"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference model from /root/reference.

The reference is a Python repo, so "running the reference" means importing
`CRCT/backbone/vilbert.py` where it lies.  It needs one un-vendored helper
(`pytorch_pretrained_bert.file_utils.cached_path`, vilbert.py:31 — a download helper,
no arithmetic), which is stubbed in `sys.modules`.  The model is built directly with
`BertForMultiModalPreTraining(config, params=params)` because
`VisualDialogEncoder.__init__` -> `from_pretrained('bert-base-uncased')`
(encoder_decorator.py:16, vilbert.py:1154-1171) needs the network.

/root/reference exists only in the build container: nothing under `-m gpu`, `smoke()`
or `bench.py`'s GPU arm may import this module.  It is used to (1) validate the
restatement in `oracle/crct_oracle.py`, (2) generate `tests/golden/*` via
`oracle/make_golden.py`.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get('CRCT_REFERENCE_ROOT', '/root/reference')


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, 'CRCT', 'backbone', 'vilbert.py'))


def import_reference():
    """Returns the reference's `backbone.vilbert` and `backbone.encoder_decorator` modules."""
    if not reference_available():
        raise RuntimeError(f'reference not present at {REFERENCE_ROOT}')
    if 'pytorch_pretrained_bert' not in sys.modules:
        pkg = types.ModuleType('pytorch_pretrained_bert')
        fu = types.ModuleType('pytorch_pretrained_bert.file_utils')
        fu.cached_path = lambda *a, **k: (_ for _ in ()).throw(RuntimeError('offline'))
        pkg.file_utils = fu
        sys.modules['pytorch_pretrained_bert'] = pkg
        sys.modules['pytorch_pretrained_bert.file_utils'] = fu
    crct_dir = os.path.join(REFERENCE_ROOT, 'CRCT')
    if crct_dir not in sys.path:
        sys.path.insert(0, crct_dir)
    import backbone.vilbert as vilbert                      # noqa: E402
    import backbone.encoder_decorator as encoder_decorator  # noqa: E402
    return vilbert, encoder_decorator


def build_reference_model(model_config: str, params: dict, seed: int = 0):
    """Reference `BertForMultiModalPreTraining` with its own constructor init under `seed`."""
    import torch
    vilbert, _ = import_reference()
    cfg = vilbert.BertConfig.from_json_file(model_config)
    torch.manual_seed(seed)
    model = vilbert.BertForMultiModalPreTraining(cfg, params=params)
    return model


class RefEncoder:
    """What `VisualDialogEncoder` would be if `from_pretrained` worked offline: same forward
    (encoder_decorator.py:19-54) around a directly-constructed `bert_pretrained`."""

    def __init__(self, model_config: str, params: dict, seed: int = 0):
        import torch
        _, dec = import_reference()
        enc = dec.VisualDialogEncoder.__new__(dec.VisualDialogEncoder)
        torch.nn.Module.__init__(enc)
        enc.bert_pretrained = build_reference_model(model_config, params, seed)
        self.module = enc
        self.glue_forward = dec.forward
