"""CPU-side checks: C-ABI exports, host logic (arena, module tree, optimizer grouping), loud failure without CUDA."""
import ctypes
import os
import re

import pytest
import torch

from cqa_crct_b200 import _lib as L
from cqa_crct_b200.spec import ModelConfig, param_spec, arena_order, arena_offsets, fused_groups, synth_state_dict
from cqa_crct_b200.synthetic import default_params, make_batch
from tests.helpers import CONFIG_DIR, ROOT


def test_library_loads_and_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'crct_b200.h')).read()
    declared = set(re.findall(r'\b(crct_[a-z0-9_]+)\s*\(', hdr))
    lib = ctypes.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(L.EXPORTS)
    assert L.lib().crct_version() >= 100


def test_no_cuda_means_loud_failure_not_fallback():
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    params = default_params(cfg_path, max_seq_len=16, max_vis_features=6)
    m = VisualDialogEncoder(params)
    batch = make_batch(2, 16, 6, 128, vocab_size=2048)
    with pytest.raises(L.CrctError):
        glue_forward(m, batch, params, evaluation=True)
    with pytest.raises(L.CrctError):
        L.device_check()


@pytest.mark.parametrize('cfg_file', ['tiny.json', 'vilbert.json'])
def test_arena_layout(cfg_file):
    cfg = ModelConfig(os.path.join(CONFIG_DIR, cfg_file))
    spec = param_spec(cfg)
    order = arena_order(cfg, spec)
    off, live_end, total = arena_offsets(order)
    assert sorted(p.name for p in order) == sorted(p.name for p in spec)
    assert all(o % 64 == 0 for o in off.values())
    seen_dead = False
    for p in order:                                   # live tensors first, dead tensors last
        seen_dead |= not p.live
        assert not (seen_dead and p.live)
    by = {p.name: p for p in spec}
    for grp in fused_groups(cfg):                     # q|k|v weights and biases back to back
        for suffix in ('.weight', '.bias'):
            cur = off[grp[0] + suffix]
            for mname in grp:
                assert off[mname + suffix] == cur
                cur += by[mname + suffix].numel
    # forward-execution order: backward finishes the arena from the tail towards the head
    first = [off[f'bert.encoder.{k}.{i}.attention.self.query.weight' if k != 'c_layer' else f'bert.encoder.c_layer.{i}.biattention.query1.weight']
             for k, i in [({'t': 'layer', 'v': 'v_layer', 'c': 'c_layer'}[kk], ii) for kk, ii in cfg.schedule()]]
    assert first == sorted(first)


def test_module_tree_and_state_dict_roundtrip():
    from cqa_crct_b200.encoder import VisualDialogEncoder
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    m = VisualDialogEncoder(default_params(cfg_path))
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, 4, 'trained')
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    names = [k for k, _ in m.named_parameters()]
    assert names == ['bert_pretrained.' + p.name for p in param_spec(cfg)]
    out = m.state_dict()
    assert set(out) == {'bert_pretrained.' + k for k in sd}
    for k, v in sd.items():
        assert torch.equal(out['bert_pretrained.' + k], v)
    assert all(p.requires_grad for _, p in m.named_parameters())          # like the reference; dead tensors just never get .grad
    # parameters are views of one flat arena
    base = m.arena.w32.data_ptr()
    for k, p in m.named_parameters():
        assert base <= p.data_ptr() < base + m.arena.total * 4
    with pytest.raises(TypeError):
        m.half()


def test_reference_init_distributions():
    from cqa_crct_b200.encoder import VisualDialogEncoder
    torch.manual_seed(0)
    m = VisualDialogEncoder(default_params(os.path.join(CONFIG_DIR, 'tiny.json')))
    sd = m.state_dict()
    w = sd['bert_pretrained.bert.encoder.layer.0.intermediate.dense.weight']
    assert abs(float(w.std()) - 0.02) < 2e-3 and abs(float(w.mean())) < 1e-3            # vilbert.py:1105
    assert float(sd['bert_pretrained.bert.encoder.layer.0.intermediate.dense.bias'].abs().sum()) == 0
    assert torch.equal(sd['bert_pretrained.bert.embeddings.LayerNorm.weight'], torch.ones(192))
    rw = sd['bert_pretrained.regressor.fusion.0.weight']
    assert float(rw.abs().max()) <= 1 / 512 ** 0.5 + 1e-6 and float(rw.std()) > 0.02     # nn.Linear default, fan_in 512


def test_dropout_hash_reference_values_and_rate():
    """The counter-based keep/drop decision is a pure function of (seed, index) — csrc/common.cuh `crct_keep`: one
    32-bit hash per element pair, 16 bits per element.  Restated here so forward and backward kernels (and future
    refactors) cannot drift apart; the GPU tests check fwd/bwd mask agreement on the device."""
    M = 0xFFFFFFFF

    def hpair(seed, idx):
        pair = idx >> 1
        h = ((pair & M) * 0x9E3779B1 + (seed & M)) & M
        h ^= (((pair >> 32) & M) * 0x85EBCA77 + ((seed >> 32) & M) * 0x27D4EB2F) & M
        h ^= h >> 15; h = h * 0x85EBCA6B & M
        h ^= h >> 13; h = h * 0xC2B2AE35 & M
        h ^= h >> 16
        return h

    def keep(seed, idx, thr):
        h = hpair(seed, idx)
        return ((h >> 16) if idx & 1 else (h & 0xFFFF)) >= thr

    for p in (0.1, 0.5):
        thr = int(p * 65536 + 0.5)
        for seed in (12345, 0xDEADBEEFCAFE1234):
            n = 40000
            rate = sum(keep(seed, i, thr) for i in range(n)) / n
            assert abs(rate - (1 - p)) < 0.01, (p, seed, rate)
    # the two halves of a pair and neighbouring pairs are uncorrelated enough for dropout
    thr = int(0.5 * 65536 + 0.5)
    both = sum(keep(7, 2 * i, thr) and keep(7, 2 * i + 1, thr) for i in range(20000)) / 20000
    assert abs(both - 0.25) < 0.02


def test_question_batch_layout_and_expansion():
    """f3 host logic: one row per candidate for the text keys, one row per question for the visual keys; the expanded batch
    is the reference's layout after cut_batch_padding (CRCT/fig_dataloader.py:690-703)."""
    from cqa_crct_b200.evaluate import candidate_groups, expand_question_batch, TEXT_KEYS, VIS_KEYS
    from cqa_crct_b200.synthetic import make_question_batch
    qb = make_question_batch(7, 32, 12, 64, seed=9, vocab_size=2048, max_ans=11)
    N, Q = int(qb['num_ans'].sum()), 7
    for k in TEXT_KEYS:
        assert qb[k].shape[0] == N, k
    for k in VIS_KEYS:
        assert qb[k].shape[0] == Q, k
    grp = candidate_groups(qb['num_ans'])
    assert grp.shape == (N,) and torch.equal(torch.bincount(grp, minlength=Q), qb['num_ans'])
    assert torch.equal(grp, torch.sort(grp).values)                      # candidates of a question are contiguous
    full = expand_question_batch(qb)
    for k in VIS_KEYS:
        assert full[k].shape[0] == N and torch.equal(full[k], qb[k][grp]), k
    # candidates of one question share the chart text and the question, differ in the answer span only
    off = 0
    for n in qb['num_ans'].tolist():
        seg = qb['segments'][off]
        same = seg != 1
        assert all(torch.equal(qb['tokens'][off + i][same], qb['tokens'][off][same]) for i in range(n))
        off += n
    assert ((qb['gt_id'] >= -1) & (qb['gt_id'] < qb['num_ans'])).all()
    qb512 = make_question_batch(16, 124, 44, 32, seed=1, total=512)
    assert int(qb512['num_ans'].sum()) == 512 and qb512['tokens'].shape == (512, 124)


def test_optimizer_range_and_bucket_bookkeeping():
    """Ranges reported by the backward are 64-element aligned and tile the live arena from its tail; the per-bucket report
    threshold only merges neighbours (cqa_crct_b200/encoder.py `_backward_stages`, parallel.py)."""
    from cqa_crct_b200.encoder import VisualDialogEncoder
    m = VisualDialogEncoder(default_params(os.path.join(CONFIG_DIR, 'vilbert.json')))
    a = m.arena
    blocks = [m._block_range(p) for p in ['bert.embeddings', 'bert.v_embeddings'] +
              [{'t': f'bert.encoder.layer.{i}', 'v': f'bert.encoder.v_layer.{i}', 'c': f'bert.encoder.c_layer.{i}'}[k] for k, i in m.cfg.schedule()]]
    assert blocks[0][0] == 0
    for (lo, hi), (lo2, hi2) in zip(blocks, blocks[1:]):
        assert hi == lo2 and lo % 64 == 0 and hi % 64 == 0
    assert blocks[-1][1] == a.offsets['bert.t_pooler.dense.weight'] and a.live_end % 64 == 0
    assert a.live_end <= a.total and sum(p.numel for p in a.spec if p.live) <= a.live_end


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """Every argument struct of include/crct_b200.h against its ctypes mirror in cqa_crct_b200/_lib.py: same size, same field
    offsets in order — compiled from the header with gcc, so a field added on one side only cannot pass silently."""
    import subprocess
    pairs = {'crct_gemm_t': L.GemmArgs, 'crct_ln_bwd_t': L.LnBwdArgs, 'crct_embed_text_t': L.EmbedTextArgs,
             'crct_embed_text_bwd_t': L.EmbedTextBwdArgs, 'crct_embed_vis_t': L.EmbedVisArgs, 'crct_embed_vis_bwd_t': L.EmbedVisBwdArgs,
             'crct_attn_fwd_t': L.AttnFwdArgs, 'crct_attn_bwd_t': L.AttnBwdArgs, 'crct_linear_t': L.LinearArgs, 'crct_loss_t': L.LossArgs,
             'crct_adamw_t': L.AdamWArgs, 'crct_select_t': L.SelectArgs, 'crct_score_t': L.ScoreArgs, 'crct_config_t': L.ConfigArgs,
             'crct_batch_t': L.BatchArgs, 'crct_out_t': L.OutArgs}
    hdr_path = os.path.join(ROOT, 'include', 'crct_b200.h')
    text = re.sub(r'/\*.*?\*/', '', open(hdr_path).read(), flags=re.S)
    structs = dict((name, body) for body, name in re.findall(r'typedef struct \{(.*?)\}\s*(\w+);', text, flags=re.S))
    assert set(pairs) <= set(structs), set(pairs) - set(structs)
    assert {n for n in structs if n.startswith('crct_')} == set(pairs), 'a header struct has no ctypes mirror in this test'
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{hdr_path}"', 'int main(void) {']
    for name in pairs:
        fields = []
        for decl in structs[name].split(';'):
            decl = decl.strip()
            if not decl:
                continue
            for d in decl.split(','):
                fields.append(re.sub(r'\[.*?\]', '', d).replace('*', ' ').split()[-1])
        lines.append(f'  printf("{name} %zu", sizeof({name}));')
        lines += [f'  printf(" %zu", offsetof({name}, {f}));' for f in fields]
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c11', '-o', str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for row in out.strip().splitlines():
        name, size, *offs = row.split()
        cls = pairs[name]
        assert int(size) == __import__('ctypes').sizeof(cls), name
        assert [int(o) for o in offs] == [getattr(cls, f).offset for f, _ in cls._fields_], name


def test_library_is_callable_from_plain_c(tmp_path):
    """The boundary is a C ABI, not a Python extension: a C program that includes include/crct_b200.h and links
    libcrct_b200.so calls it without torch.  Without a GPU the device check must fail with a status and a message."""
    import subprocess
    src = tmp_path / 'use.c'
    src.write_text('#include <stdio.h>\n#include "crct_b200.h"\n'
                   'int main(void) {\n'
                   '  int v = crct_version();\n'
                   '  int rc = crct_device_check();\n'
                   '  crct_gemm_t g = {0};\n'
                   '  int rc2 = crct_gemm_bf16(&g, 0);\n'
                   '  printf("%d|%d|%d|%s\\n", v, rc, rc2, crct_last_error());\n'
                   '  crct_config_t c = {0};\n'          # whole-model entry points from plain C: create / size / destroy (tiny.json numbers)
                   '  c.hidden_size = 192; c.num_hidden_layers = 3; c.num_attention_heads = 4; c.intermediate_size = 256;\n'
                   '  c.v_hidden_size = 128; c.v_num_hidden_layers = 2; c.v_num_attention_heads = 2; c.v_intermediate_size = 128; c.v_feature_size = 128;\n'
                   '  c.bi_hidden_size = 128; c.bi_num_attention_heads = 4; c.max_position_embeddings = 64; c.num_connections = 2;\n'
                   '  c.v_biattention_id[1] = 1; c.t_biattention_id[0] = 1; c.t_biattention_id[1] = 2; c.l1 = 1; c.tol_margin = 0.01f;\n'
                   '  crct_handle_t h = 0;\n'
                   '  int rc3 = crct_create(&c, &h);\n'
                   '  size_t ws = crct_workspace_bytes(h, 6, 6, 32, 12);\n'
                   '  printf("%d|%zu|%d\\n", rc3, ws, crct_destroy(h));\n'
                   '  return 0;\n}\n')
    exe = tmp_path / 'use'
    libdir = os.path.dirname(L.LIB_PATH)
    subprocess.run(['gcc', '-std=c11', '-I', os.path.join(ROOT, 'include'), '-o', str(exe), str(src), '-L', libdir, '-l:libcrct_b200.so',
                    f'-Wl,-rpath,{libdir}'], check=True)
    out, out2 = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    rc3, ws, rc4 = out2.split('|')
    assert int(rc3) == 0 and int(ws) > 0 and int(rc4) == 0
    v, rc, rc2, msg = out.split('|', 3)
    assert int(v) >= 100
    assert int(rc2) != 0 and msg                      # null operands are rejected with a message, never dereferenced
    if not torch.cuda.is_available():
        assert int(rc) != 0


def test_dropout8_equals_elementwise_keep_native(tmp_path):
    """csrc/common.cuh: `dropout8` (what the GEMM epilogue, LayerNorm and embedding kernels call on 8-element chunks) against
    `crct_keep` (what the attention kernels and the fp32 check mode evaluate per element) — the forward/backward and the
    bf16/fp32 mask agreement rests on these two being the same function.  Built and run on the host with nvcc."""
    import shutil
    import subprocess
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        pytest.skip('nvcc not available')
    exe = tmp_path / 'dropout8_check'
    subprocess.run([nvcc, '-std=c++17', '-O2', '--expt-relaxed-constexpr', '-gencode', 'arch=compute_100a,code=sm_100a',
                    '-I', os.path.join(ROOT, 'cqa_crct_b200', 'csrc'), '-o', str(exe), os.path.join(ROOT, 'tests', 'native', 'dropout8_check.cu')],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith('0 mismatches'), out.stdout


def test_argument_validation_returns_status_and_message():
    """Error behaviour of the boundary: bad arguments are rejected on the host with a negative crct_status_t and
    a message in crct_last_error() before anything is launched — no exception, no crash, no fallback.  Runs without a GPU."""
    import ctypes as C
    lib = L.lib()
    msg = lambda: lib.crct_last_error().decode()
    buf = (C.c_char * 4096)()
    p = (C.addressof(buf) + 255) & ~255
    ARG, SHAPE = -1, -4
    g = L.GemmArgs()
    assert lib.crct_gemm_bf16(C.byref(g), None) == ARG and 'null pointer' in msg()
    g.A = g.B = g.D = p
    g.M, g.N, g.K, g.lda, g.ldb, g.ldd = 128, 100, 64, 64, 64, 104
    assert lib.crct_gemm_bf16(C.byref(g), None) == SHAPE and 'multiple of 8' in msg()
    g.N, g.ldd, g.epilogue = 128, 128, 9
    assert lib.crct_gemm_bf16(C.byref(g), None) == ARG and 'unknown epilogue' in msg()
    g.epilogue = L.EPI_MUL
    assert lib.crct_gemm_bf16(C.byref(g), None) == ARG and 'needs aux' in msg()
    g.epilogue, g.A = L.EPI_BIAS, p + 2
    assert lib.crct_gemm_bf16(C.byref(g), None) == ARG and '16-byte aligned' in msg()
    a = L.AttnFwdArgs()
    a.q = a.k = a.v = a.mask_add = a.out = p
    a.B, a.nh, a.dh, a.Lq, a.Lk = 1, 2, 40, 8, 8
    a.ldq = a.ldk = a.ldv = a.ldo = 80
    assert lib.crct_attn_fwd(C.byref(a), None) == SHAPE and 'head dim 40' in msg()
    assert lib.crct_layernorm_fwd(p, p, p, p, None, None, None, 4, 2048, 0, None, None) == SHAPE and 'row width 2048' in msg()
    assert lib.crct_layernorm_fwd(p, p, p, p, None, p, None, 4, 768, 0, None, None) == ARG
    assert lib.crct_expand_blocks(p, p, p, 4, 6, None) == ARG
    assert lib.crct_select_answers(C.byref(L.SelectArgs()), None) == ARG and 'null pointer' in msg()
    lin = L.LinearArgs()
    lin.A = lin.B = lin.C = p
    lin.M = lin.N = lin.K = 4
    lin.act = 7
    assert lib.crct_linear_f32(C.byref(lin), None) == ARG and 'unknown activation' in msg()
    assert lib.crct_linear_f32_batched(C.byref(lin), 13, None) == ARG
    f = L.GemmArgs()
    assert lib.crct_f32_gemm(C.byref(f), None) == ARG and 'crct_f32_gemm' in msg()


def test_whole_model_entry_points_validate_on_the_host(tmp_path):
    """crct_create / crct_bind_params / crct_workspace_bytes / crct_forward without a GPU: configuration and binding
    errors come back as status codes with messages; the forward refuses to run without an sm_100 device (no fallback)."""
    import ctypes as C
    from cqa_crct_b200.capi import config_args
    lib = L.lib()
    cfg = ModelConfig(os.path.join(CONFIG_DIR, 'tiny.json'))
    params = default_params(os.path.join(CONFIG_DIR, 'tiny.json'))
    h = C.c_void_p()
    bad = config_args(cfg, params)
    bad.num_attention_heads = 5                                   # 192 % 5 != 0   (vilbert.py:364-368)
    assert lib.crct_create(C.byref(bad), C.byref(h)) == -4 and b'heads' in lib.crct_last_error()
    bad = config_args(cfg, params)
    bad.t_biattention_id[1] = 7                                   # beyond the layer count (vilbert.py:193-194)
    assert lib.crct_create(C.byref(bad), C.byref(h)) == -1 and b'biattention' in lib.crct_last_error()
    assert lib.crct_create(C.byref(config_args(cfg, params)), C.byref(h)) == 0 and h.value
    assert lib.crct_workspace_bytes(h, 6, 6, 32, 12) > 0 and lib.crct_workspace_bytes(h, 0, 6, 32, 12) == 0
    assert lib.crct_workspace_bytes(h, 12, 6, 32, 12) > lib.crct_workspace_bytes(h, 6, 6, 32, 12)
    # binding: host arithmetic only (pointers are not dereferenced) — adjacency of the fused projections is checked
    spec = [p for p in arena_order(cfg, param_spec(cfg)) if p.live]
    off, _, _ = arena_offsets(arena_order(cfg, param_spec(cfg)))
    base32, base16 = 0x10000000, 0x40000000
    n = len(spec)
    names = (C.c_char_p * n)(*[p.name.encode() for p in spec])
    numel = (C.c_size_t * n)(*[p.numel for p in spec])
    w32 = (C.c_void_p * n)(*[base32 + 4 * off[p.name] for p in spec])
    w16 = (C.c_void_p * n)(*[base16 + 2 * off[p.name] for p in spec])
    assert lib.crct_bind_params(h, names, w32, w16, numel, n) == 0, lib.crct_last_error()
    k = [p.name for p in spec].index('bert.encoder.layer.1.attention.self.key.weight')
    w16b = (C.c_void_p * n)(*[base16 + 2 * off[p.name] + (64 if i == k else 0) for i, p in enumerate(spec)])
    assert lib.crct_bind_params(h, names, w32, w16b, numel, n) == -1 and b'adjacent' in lib.crct_last_error()
    assert lib.crct_bind_params(h, names, w32, w16, numel, n - 1) == -1 and b'missing' in lib.crct_last_error()     # last tensor dropped
    if not torch.cuda.is_available():
        ba, o = L.BatchArgs(), L.OutArgs()
        assert lib.crct_forward(h, C.byref(ba), C.byref(o), C.c_void_p(16), 16, None) in (-2, -3)       # no device: status, not a crash
        assert lib.crct_last_error()
    assert lib.crct_destroy(h) == 0
