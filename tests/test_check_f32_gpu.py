"""fp32 CHECK MODE on the B200 (`VisualDialogEncoder(params, precision='fp32')`, csrc/check_f32.cu).

(1) check mode vs the fp64 oracle and the reference's golden vectors — SURVEY.md §8c tolerances for the check mode:
        class logits / regression output   <= 2e-5 of the tensor's scale  (measured 0.9e-6 .. 4.5e-6; SURVEY bar 1e-4)
        loss                               <= 1e-5 absolute
        gradients                          per tensor ||got - ref|| <= 1e-3 ||ref|| + 1e-6 max_t ||ref_t|| (worst 0.26 of it),
                                           all tensors together <= 2e-5 relative L2 (measured 2.5e-7 .. 4.2e-6)
    (fp32 storage and arithmetic; the oracle runs in fp64, the goldens come from the reference in fp32);
(2) the bf16 production path vs check mode ON THE DEVICE, dropout ON (both draw the same counter-based masks), at a size
    the CPU oracle does not reach in seconds: logits <= 1e-2 of scale (measured 6.8e-3 .. 7.6e-3), loss <= 1e-3 (6e-5 .. 9e-5),
    regression <= 1e-3 (2.6e-4), gradients global relative L2 <= 7.5e-2 (5.8e-2; 2.7e-2 .. 6.3e-2 at the stress shape, depending on
    which attention kernels run: the same chaotic floor) — bars = measured x 1.3, see test_model_gpu.py / DESIGN.md for the floor they sit on;
(3) the check-mode GEMM and attention kernels on their own against torch (ragged shapes, every operand major / epilogue)."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import _lib as L                                   # noqa: E402
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward   # noqa: E402
from cqa_crct_b200.spec import ModelConfig, synth_state_dict          # noqa: E402
from cqa_crct_b200.synthetic import default_params, make_batch        # noqa: E402
from oracle import crct_oracle as O                                   # noqa: E402
from tests.helpers import CONFIG_DIR, load_golden, golden_inputs      # noqa: E402

DEV = 'cuda'


def build(rec, precision):
    cfg_path, cfg, sd, batch = golden_inputs(rec)
    params = default_params(cfg_path, device='cuda', max_seq_len=rec['T'], max_vis_features=rec['R'], L1=rec['l1'])
    m = VisualDialogEncoder(params, precision=precision)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to('cuda').eval()
    return m, params, cfg, sd, batch, {k: v.to('cuda') for k, v in batch.items()}


def scale_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize('name', ['tiny_eval', 'full_eval_b8', 'full_eval_b8_mild'])
def test_check_mode_eval_matches_oracle_and_golden(name):
    rec = load_golden(name)
    m, params, cfg, sd, batch, gb = build(rec, 'fp32')
    with torch.no_grad():
        _, _, _, _, scores, reg = glue_forward(m, gb, params, evaluation=True)
    out, _ = O.forward(sd, O.Config(cfg.__dict__), batch, train=False, l1=rec['l1'], keep_cache=False, dtype=torch.float64)
    e_logit, e_reg = scale_err(scores, out['logits']), scale_err(reg[0], out['reg_pred'])
    print(f'{name}: logits {e_logit:.2e} reg {e_reg:.2e} golden {scale_err(scores, rec["logits"]):.2e}')
    assert e_logit < 2e-5 and e_reg < 2e-5
    assert scale_err(scores, rec['logits']) < 2e-5 and scale_err(reg[0], rec['reg_pred']) < 2e-5      # the reference itself (fp32)
    assert torch.equal(scores.cpu().argmax(1), rec['logits'].argmax(1)) or \
        float((rec['logits'][:, 0] - rec['logits'][:, 1]).abs().min()) < 1e-4 * float(rec['logits'].abs().max())


@pytest.mark.parametrize('name', ['tiny_train_l1', 'tiny_train_smooth', 'tiny_ragged', 'full_train_b4', 'full_train_b4_mild'])
def test_check_mode_gradients_match_oracle(name):
    rec = load_golden(name)
    m, params, cfg, sd, batch, gb = build(rec, 'fp32')
    m.zero_grad()
    loss, _, nsp, _, scores, reg, _ = glue_forward(m, gb, params)
    loss.backward()
    torch.cuda.synchronize()
    out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=rec['l1'], dtype=torch.float64)
    g = O.backward(cache)
    assert abs(float(loss) - float(out['loss'])) < 1e-5 and abs(float(loss) - rec['loss']) < 1e-5
    assert scale_err(scores, out['logits']) < 2e-5
    named = dict(m.bert_pretrained.named_parameters())
    gnorm = max(float(v.norm()) for v in g.values())
    worst, num, den = (0.0, None), 0.0, 0.0
    for k, ref in g.items():
        got, ref = named[k].grad.double().cpu(), ref.double()
        err, rn = float((got - ref).norm()), float(ref.norm())
        num, den = num + err * err, den + rn * rn
        excess = err / (1e-3 * rn + 1e-6 * gnorm)
        if excess > worst[0]:
            worst = (excess, k)
    print(f'{name}: global rel {(num / den) ** 0.5:.2e}, worst tensor at {worst[0]:.2f} of its bound ({worst[1]})')
    assert worst[0] < 1.0, worst
    assert (num / den) ** 0.5 < 2e-5
    for k, p in named.items():
        if k not in g:
            assert p.grad is None or float(p.grad.abs().sum()) == 0.0, k


def test_bf16_path_against_check_mode_with_dropout_on_full_model():
    """B = 12 sequences of the full model, train mode (every dropout site active): the production path and the check mode
    share seeds and element counters, so they apply the same masks and differ only by bf16 rounding."""
    cfg_path = os.path.join(CONFIG_DIR, 'vilbert.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, 2, 'mild')
    batch = make_batch(12, 124, 44, cfg.v_feature_size, seed=77)
    gb = {k: v.to(DEV) for k, v in batch.items()}
    res = {}
    for precision in ('fp32', 'bf16'):
        torch.manual_seed(5)                                   # same dropout salt in both runs
        params = default_params(cfg_path, device='cuda', L1=True)
        m = VisualDialogEncoder(params, precision=precision)
        m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
        m.to(DEV).train()
        m.zero_grad()
        loss, _, _, _, scores, reg, _ = glue_forward(m, gb, params)
        loss.backward()
        torch.cuda.synchronize()
        res[precision] = (float(loss), scores.detach().clone(), reg[0].detach().clone(), m.arena.g32[:m.arena.live_end].clone())
        del m
    (l32, s32, r32, g32), (l16, s16, r16, g16) = res['fp32'], res['bf16']
    e_logit, e_grad = scale_err(s16, s32), float((g16 - g32).norm() / g32.norm())
    print(f'bf16 vs fp32 check, dropout on: loss {l16:.5f} / {l32:.5f}, logits {e_logit:.2e}, reg {scale_err(r16, r32):.2e}, gradients {e_grad:.2e}')
    assert abs(l16 - l32) < 1e-3 and e_logit < 1e-2 and scale_err(r16, r32) < 1e-3 and e_grad < 7.5e-2
    # a different mask stream would not be a rounding-sized difference: same model, other salt
    torch.manual_seed(6)
    params = default_params(cfg_path, device='cuda', L1=True)
    m = VisualDialogEncoder(params, precision='bf16')
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to(DEV).train()
    m.zero_grad()
    glue_forward(m, gb, params)[0].backward()
    other = float((m.arena.g32[:m.arena.live_end] - g32).norm() / g32.norm())
    assert other > 2 * e_grad, (other, e_grad)


EPIS = [(L.EPI_BIAS, 'bias'), (L.EPI_BIAS_GELU, 'gelu'), (L.EPI_BIAS_RES, 'res'), (L.EPI_MUL, 'mul'), (L.EPI_F32, 'f32')]


@pytest.mark.parametrize('a_major,b_major', [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize('epi,_n', EPIS)
def test_check_gemm_against_torch(a_major, b_major, epi, _n):
    M, N, K = 131, 77, 203
    g = torch.Generator().manual_seed(M + epi)
    A = torch.randn((K, M) if a_major else (M, K), generator=g).to(DEV)
    B = torch.randn((K, N) if b_major else (N, K), generator=g).to(DEV)
    bias, aux = torch.randn(N, generator=g).to(DEV), torch.randn(M, N, generator=g).to(DEV)
    Am = (A.t() if a_major else A).double()
    Bm = (B.t() if b_major else B).double()
    acc = Am @ Bm.t()
    D = torch.full((M, N), 3.0, device=DEV)
    D2 = torch.empty(M, N, device=DEV) if epi == L.EPI_BIAS_GELU else None
    L.gemm(A, B, D, M=M, N=N, K=K, a_major=a_major, b_major=b_major, epilogue=epi, bias=None if epi in (L.EPI_MUL, L.EPI_F32) else bias,
           aux=aux if epi in (L.EPI_BIAS_RES, L.EPI_MUL) else None, D2=D2, accumulate=1 if epi == L.EPI_F32 else 0)
    if epi == L.EPI_BIAS:
        ref = acc + bias.double()
    elif epi == L.EPI_BIAS_GELU:
        u = acc + bias.double()
        ref = 0.5 * u * (1 + torch.erf(u / math.sqrt(2)))
        d = 0.5 * (1 + torch.erf(u / math.sqrt(2))) + u * torch.exp(-0.5 * u * u) / math.sqrt(2 * math.pi)
        assert float((D2.double() - d).abs().max()) < 1e-5
    elif epi == L.EPI_BIAS_RES:
        ref = acc + bias.double() + aux.double()
    elif epi == L.EPI_MUL:
        ref = acc * aux.double()
    else:
        ref = acc + 3.0
    assert float((D.double() - ref).abs().max()) < 2e-4 * float(ref.abs().max())


@pytest.mark.parametrize('nh,dh,Lq,Lk,p', [(3, 48, 37, 37, 0.0), (2, 32, 20, 51, 0.0), (2, 64, 9, 13, 0.0), (2, 32, 33, 40, 0.3)])
def test_check_attention_against_torch(nh, dh, Lq, Lk, p):
    B, H = 2, nh * dh
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    q, k, v = [torch.randn(B * n, H, generator=g).to(DEV) for n in (Lq, Lk, Lk)]
    dout = torch.randn(B * Lq, H, generator=g).to(DEV)
    mask = torch.zeros(B, Lk)
    mask[1, Lk - 3:] = -10000.0
    mask = mask.to(DEV)
    out, lse = torch.empty(B * Lq, H, device=DEV), torch.empty(B, nh, Lq, device=DEV)
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    L.SALT = None
    kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H, dropout_p=p, seed=11)
    L.attn_fwd(q, k, v, mask, out, lse, **kw)
    L.attn_bwd(q, k, v, mask, out, dout, lse, dq, dk, dv, lddo=H, lddq=H, lddk=H, lddv=H, **kw)
    qd, kd, vd = [t.double().view(B, -1, nh, dh).transpose(1, 2).requires_grad_() for t in (q, k, v)]
    s = qd @ kd.transpose(-1, -2) / math.sqrt(dh) + mask.double()[:, None, None, :]
    pr = torch.softmax(s, -1)
    if p > 0:      # the kernel's own mask: recovered from a forward with v = identity-like probe is overkill; use the hash restated on the host
        from tests.test_host_cpu import test_dropout_hash_reference_values_and_rate  # noqa: F401  (documents the hash)
        idx = torch.arange(B * nh * Lq * Lk).view(B, nh, Lq, Lk)
        keep = torch.tensor([_keep(11, int(i), int(p * 65536 + 0.5)) for i in idx.flatten().tolist()]).view_as(idx).to(DEV)
        pr = pr * keep.double() / (1 - p)
    o = (pr @ vd).transpose(1, 2).reshape(B * Lq, H)
    o.backward(dout.double())
    assert float((out.double() - o.detach()).abs().max()) < 1e-5
    for got, ref in ((dq, qd.grad), (dk, kd.grad), (dv, vd.grad)):
        ref = ref.transpose(1, 2).reshape(-1, H)
        assert float((got.double() - ref).abs().max()) < 2e-5 * max(1.0, float(ref.abs().max()))


def _keep(seed, idx, thr):
    Mk = 0xFFFFFFFF
    pair = idx >> 1
    h = ((pair & Mk) * 0x9E3779B1 + (seed & Mk)) & Mk
    h ^= (((pair >> 32) & Mk) * 0x85EBCA77 + ((seed >> 32) & Mk) * 0x27D4EB2F) & Mk
    h ^= h >> 15; h = h * 0x85EBCA6B & Mk
    h ^= h >> 13; h = h * 0xC2B2AE35 & Mk
    h ^= h >> 16
    return ((h >> 16) if idx & 1 else (h & 0xFFFF)) >= thr


@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_edge_case_batches(precision):
    """The degenerate batches of tests/test_oracle_vs_reference.py::test_edge_case_batches_against_reference (B = 1, no row /
    every row needing regression, only the <IMG> region visible, zero target) through the CUDA path, forward and backward,
    against the fp64 oracle: check-mode bars for fp32, the bf16 floor for the production kernels."""
    from tests.test_oracle_vs_reference import _edge_batches
    cfg_path = os.path.join(CONFIG_DIR, 'tiny.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, seed=5, style='trained')
    params = default_params(cfg_path, device='cuda', max_seq_len=24, max_vis_features=9, L1=True)
    m = VisualDialogEncoder(params, precision=precision)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to(DEV).eval()
    tol_logit, tol_loss, tol_grad = (2e-5, 1e-5, 2e-5) if precision == 'fp32' else (1e-2, 1e-3, 5e-2)      # bf16 measured: logits <= 4.2e-3, loss <= 1e-4, gradients <= 3.1e-2
    for name, batch in _edge_batches(cfg).items():
        gb = {k: v.to(DEV) for k, v in batch.items()}
        m.zero_grad()
        loss, _, _, _, scores, reg, _ = glue_forward(m, gb, params)
        loss.backward()
        torch.cuda.synchronize()
        out, cache = O.forward(sd, O.Config(cfg.__dict__), batch, train=True, l1=True, dtype=torch.float64)
        g = O.backward(cache)
        assert scale_err(scores, out['logits']) < tol_logit, name
        assert abs(float(loss) - float(out['loss'])) < tol_loss, name
        assert (int(reg[3][0]), int(reg[3][1])) == tuple(out['reg_right']) or precision == 'bf16', name
        named = dict(m.bert_pretrained.named_parameters())
        num = sum(float((named[k].grad.double().cpu() - v).norm() ** 2) for k, v in g.items())
        den = sum(float(v.norm() ** 2) for v in g.values())
        print(f'edge case {name} [{precision}]: logits {scale_err(scores, out["logits"]):.2e}, loss {abs(float(loss) - float(out["loss"])):.1e}, '
              f'gradients {(num / den) ** 0.5:.2e}')
        assert (num / den) ** 0.5 < tol_grad, (name, (num / den) ** 0.5)
        assert torch.isfinite(m.arena.g32).all(), name


def test_full_size_b80_properties():
    """BASELINE.json's full train size (B = 80, T = 124, R = 44, full model), where the CPU oracle is out of reach in seconds:
    (1) batch independence — every sequence's class logits and regression output from the B = 80 eval forward equal those
        of the same sequences run in chunks of 8 (each output element is produced by one CTA in a fixed K order, so the
        result could only change where the tile policy picked another kernel for the smaller problem: measured exactly 0,
        bar 1e-6 of scale);
    (2) the bf16 train step (dropout ON) against the fp32 check mode with the same masks: loss, logits, all gradients
        (measured: loss 0.77317 vs 0.77321, logits 2.0e-2, gradients 8.0e-2)."""
    cfg_path = os.path.join(CONFIG_DIR, 'vilbert.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, 2, 'mild')
    batch = make_batch(80, 124, 44, cfg.v_feature_size, seed=1234)
    gb = {k: v.to(DEV) for k, v in batch.items()}
    res = {}
    for precision in ('bf16', 'fp32'):
        torch.manual_seed(5)
        params = default_params(cfg_path, device='cuda', L1=True)
        m = VisualDialogEncoder(params, precision=precision)
        m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
        m.to(DEV)
        if precision == 'bf16':
            m.eval()
            with torch.no_grad():
                full = glue_forward(m, gb, params, evaluation=True)
                parts = [glue_forward(m, gb, params, evaluation=True, sample_ids=slice(i, i + 8)) for i in range(0, 80, 8)]
            s_full, r_full = full[4], full[5][0]
            s_part = torch.cat([p[4] for p in parts])
            r_part = torch.cat([p[5][0] for p in parts])
            d_s, d_r = scale_err(s_part, s_full), scale_err(r_part, r_full)
            print(f'B=80 vs 10 x B=8: logits {d_s:.2e}, regression {d_r:.2e}')
            assert d_s < 1e-6 and d_r < 1e-6
        m.train()
        m.zero_grad()
        loss, _, _, _, scores, reg, _ = glue_forward(m, gb, params)
        loss.backward()
        torch.cuda.synchronize()
        res[precision] = (float(loss), scores.detach().clone(), m.arena.g32[:m.arena.live_end].clone())
        del m
        torch.cuda.empty_cache()
    (l16, s16, g16), (l32, s32, g32) = res['bf16'], res['fp32']
    e_logit, e_grad = scale_err(s16, s32), float((g16 - g32).norm() / g32.norm())
    print(f'B=80 train, dropout on, bf16 vs fp32 check: loss {l16:.5f} / {l32:.5f}, logits {e_logit:.2e}, gradients {e_grad:.2e}')
    assert abs(l16 - l32) < 1e-3 and e_logit < 1e-2 and e_grad < 7.5e-2


def test_stress_shape_bf16_vs_check_mode():
    """BASELINE config 5 (2x regions, 2x tokens: T = 248, R = 88), full model, B = 3, dropout on: the production path (split
    two-pass attention backward — dS^T of a 248 x 248 head does not fit next to the operand tiles) against the check mode."""
    cfg_path = os.path.join(CONFIG_DIR, 'vilbert.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, 2, 'mild')
    batch = make_batch(3, 248, 88, cfg.v_feature_size, seed=99)
    gb = {k: v.to(DEV) for k, v in batch.items()}
    res = {}
    for precision in ('fp32', 'bf16'):
        torch.manual_seed(9)
        params = default_params(cfg_path, device='cuda', max_seq_len=248, max_vis_features=88, L1=False)
        m = VisualDialogEncoder(params, precision=precision)
        m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
        m.to(DEV).train()
        m.zero_grad()
        loss, _, _, _, scores, reg, _ = glue_forward(m, gb, params)
        loss.backward()
        torch.cuda.synchronize()
        res[precision] = (float(loss.detach()), scores.detach().clone(), m.arena.g32[:m.arena.live_end].clone())
        del m
        torch.cuda.empty_cache()
    (l32, s32, g32), (l16, s16, g16) = res['fp32'], res['bf16']
    e_logit, e_grad = scale_err(s16, s32), float((g16 - g32).norm() / g32.norm())
    print(f'stress shape, bf16 vs fp32 check: loss {l16:.5f} / {l32:.5f}, logits {e_logit:.2e}, gradients {e_grad:.2e}')
    assert abs(l16 - l32) < 1e-3 and e_logit < 1e-2 and e_grad < 7.5e-2


def test_question_batch_full_model_bit_identical_to_replicated_layout():
    """f3 at the real model size: 6 questions / 96 candidate sequences; de-duplicated visual inputs (embedding once per
    question, fanned out on the device) give the same logits bit for bit as the reference's replicated layout."""
    from cqa_crct_b200.evaluate import evaluate_batch, expand_question_batch
    from cqa_crct_b200.synthetic import make_question_batch
    cfg_path = os.path.join(CONFIG_DIR, 'vilbert.json')
    cfg = ModelConfig(cfg_path)
    sd = synth_state_dict(cfg, 228, 2, 'mild')
    params = default_params(cfg_path, device='cuda', L1=True)
    m = VisualDialogEncoder(params)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to(DEV).eval()
    qb = make_question_batch(6, 124, 44, cfg.v_feature_size, seed=3, total=96)
    out = evaluate_batch(m, qb, params, eval_batch_size=40)               # chunks cut through questions
    full = {k: v.to(DEV) for k, v in expand_question_batch(qb).items()}
    with torch.no_grad():
        scores, reg = glue_forward(m, full, params, evaluation=True)[4:6]
    assert torch.equal(scores, out['logits'])
    sel = out['answers'].cpu()
    off = torch.cumsum(qb['num_ans'], 0) - qb['num_ans']
    assert torch.equal(out['reg_output'].cpu(), reg[0].cpu()[off + sel])
