"""Var-len ("packed") rows and the fp32 pre-LayerNorm sum on the B200.

The reference pads to T tokens / R regions and masks additively (CRCT/backbone/vilbert.py:1380-1396, CRCT/utils.py:152,178);
SURVEY.md §2.3: padded rows influence nothing.  The production path therefore runs only the valid rows (csrc/varlen.cu) with
device-side row counts.  These tests pin (1) the row maps, (2) every kernel's `rows_dev` / `cu` form against its padded form,
(3) the whole model: packed == padded, bit for bit on the forward outputs, to fp32 summation order on the gradients."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import _lib as L                                     # noqa: E402
from cqa_crct_b200.encoder import VisualDialogEncoder, glue_forward     # noqa: E402
from cqa_crct_b200.synthetic import default_params                      # noqa: E402
from tests.helpers import load_golden, golden_inputs                    # noqa: E402

DEV = 'cuda'


def bf(x):
    return x.to(torch.bfloat16)


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def ref_row_map(mask):
    m = mask != 0
    m = m.clone()
    m[~m.any(1), 0] = True                    # a sample without valid rows keeps its row 0
    lens = m.sum(1)
    cu = torch.zeros(mask.shape[0] + 1, dtype=torch.int64)
    cu[1:] = lens.cumsum(0)
    src = m.flatten().nonzero().flatten()
    return cu, src


@pytest.mark.parametrize('B,Lm', [(1, 5), (7, 44), (80, 124), (512, 124), (300, 33)])
@pytest.mark.parametrize('dtype', [torch.bool, torch.int64, torch.float32])
def test_row_map_matches_torch(B, Lm, dtype):
    g = torch.Generator().manual_seed(B * 131 + Lm)
    lens = torch.randint(0 if B > 1 else 1, Lm + 1, (B,), generator=g)
    prefix = torch.arange(Lm).unsqueeze(0) < lens.unsqueeze(1)
    holes = torch.rand(B, Lm, generator=g) < 0.7              # second case: arbitrary (non-prefix) masks
    for mask in (prefix, prefix & holes):
        cu_ref, src_ref = ref_row_map(mask)
        cu = torch.empty(B + 1, dtype=torch.int32, device=DEV)
        src = torch.full((B * Lm,), -1, dtype=torch.int32, device=DEV)
        L.row_map(mask.to(dtype).to(DEV), cu, src)
        assert torch.equal(cu.cpu().long(), cu_ref)
        n = int(cu_ref[-1])
        assert torch.equal(src[:n].cpu().long(), src_ref)


def test_group_map_and_gather_rows():
    g = torch.Generator().manual_seed(5)
    Q, R, H, N = 9, 12, 64, 40
    mask = torch.arange(R).unsqueeze(0) < torch.randint(1, R + 1, (Q, 1), generator=g)
    cu_q = torch.empty(Q + 1, dtype=torch.int32, device=DEV)
    src_q = torch.empty(Q * R, dtype=torch.int32, device=DEV)
    L.row_map(mask.to(DEV), cu_q, src_q)
    group = torch.randint(0, Q, (N,), generator=g).sort().values.to(DEV)
    cu = torch.empty(N + 1, dtype=torch.int32, device=DEV)
    src = torch.empty(N * R, dtype=torch.int32, device=DEV)
    L.group_map(cu_q, group, cu, src)
    lens = (cu_q[1:] - cu_q[:-1])[group]
    assert torch.equal((cu[1:] - cu[:-1]), lens) and int(cu[0]) == 0
    vq = bf(torch.randn(Q * R, H, device=DEV))
    v = torch.zeros(N * R, H, dtype=torch.bfloat16, device=DEV)
    L.gather_rows(vq, src, v, rows_dev=cu[N:])
    for n in range(N):
        q = int(group[n])
        assert torch.equal(v[int(cu[n]):int(cu[n + 1])], vq[int(cu_q[q]):int(cu_q[q + 1])])
    assert float(v[int(cu[N]):].abs().sum()) == 0.0          # rows past the device-side count are not written


@pytest.mark.parametrize('M,N,K,rows', [(1000, 768, 768, 617), (9920, 2304, 768, 6899), (3520, 1024, 1024, 1), (512, 128, 64, 512)])
def test_gemm_device_row_count_forward_and_dgrad(M, N, K, rows):
    """a_rows_dev with a_major = 0: tiles past the count are skipped, rows past it keep their old contents (NaN garbage in A
    past the count must not leak)."""
    torch.manual_seed(M + rows)
    A, B = bf(torch.randn(M, K, device=DEV) * 0.5), bf(torch.randn(N, K, device=DEV) * 0.5)
    A[rows:] = float('nan')
    bias, aux = torch.randn(N, device=DEV), bf(torch.randn(M, N, device=DEV))
    n = torch.tensor([rows], dtype=torch.int32, device=DEV)
    ref = A[:rows].float() @ B.float().t() + bias
    for epi, want in ((L.EPI_BIAS, ref), (L.EPI_BIAS_RES, ref + aux[:rows].float()), (L.EPI_BIAS_GELU, torch.nn.functional.gelu(ref))):
        D = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
        L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, epilogue=epi, aux=aux if epi == L.EPI_BIAS_RES else None, rows_dev=n)
        assert relmax(D[:rows].float(), want) < 6e-3
        assert bool((D[rows:] == 7.0).all())
    Z = torch.full((M, N), 7.0, device=DEV)                                   # fp32 pre-LayerNorm sum, fp32 residual
    aux32 = torch.randn(M, N, device=DEV)
    aux32[rows:] = float('nan')
    L.gemm(A, B, Z, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES_F32, aux=aux32, rows_dev=n)
    assert relmax(Z[:rows], ref + aux32[:rows]) < 1e-5
    assert bool((Z[rows:] == 7.0).all())
    W = bf(torch.randn(K, N, device=DEV) * 0.5)                               # dgrad form
    D = torch.full((M, N), 7.0, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, W, D, M=M, N=N, K=K, b_major=1, epilogue=L.EPI_MUL, aux=aux, rows_dev=n)
    assert relmax(D[:rows].float(), (A[:rows].float() @ W.float()) * aux[:rows].float()) < 6e-3
    assert bool((D[rows:] == 7.0).all())


def test_gemm_res_f32_matches_bf16_epilogue_and_dropout_stream():
    torch.manual_seed(3)
    M, N, K = 1024, 768, 512
    A, B = bf(torch.randn(M, K, device=DEV) * 0.5), bf(torch.randn(N, K, device=DEV) * 0.5)
    bias, aux = torch.randn(N, device=DEV), bf(torch.randn(M, N, device=DEV))
    for cg in (1, 2):
        for bn in (128, 256):
            Z = torch.empty(M, N, device=DEV)
            D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
            L.gemm(A, B, Z, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES_F32, aux=aux.float(), dropout_p=0.1, seed=9, cta_group=cg, block_n=bn)
            L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES, aux=aux, dropout_p=0.1, seed=9, cta_group=cg, block_n=bn)
            assert torch.equal(bf(Z), D)       # same accumulators, same dropout decisions: the bf16 form is the rounded fp32 form
    Mr, Nr = 200, 72                           # ragged: partial tile rows and a partial 16-column step
    A2, B2 = bf(torch.randn(Mr, K, device=DEV) * 0.5), bf(torch.randn(Nr, K, device=DEV) * 0.5)
    aux2 = torch.randn(Mr, Nr, device=DEV)
    Z2 = torch.empty(Mr, Nr, device=DEV)
    L.gemm(A2, B2, Z2, M=Mr, N=Nr, K=K, bias=bias[:Nr].contiguous(), epilogue=L.EPI_BIAS_RES_F32, aux=aux2)
    assert relmax(Z2, A2.float() @ B2.float().t() + bias[:Nr] + aux2) < 1e-5


@pytest.mark.parametrize('rows,No,Ki,valid', [(1000, 576, 192, 617), (9920, 768, 768, 6899), (9920, 3072, 768, 6848), (3520, 1024, 1024, 1903),
                                              (992, 768, 3072, 64), (992, 256, 128, 3)])
@pytest.mark.parametrize('split', [0, 1, 7])
def test_gemm_device_row_count_wgrad(rows, No, Ki, valid, split):
    """a_rows_dev in the wgrad form: GEMM-K = the device-side row count; NaN rows past it contribute nothing."""
    torch.manual_seed(valid)
    dy, x = bf(torch.randn(rows, No, device=DEV) * 0.5), bf(torch.randn(rows, Ki, device=DEV) * 0.5)
    dy[valid:] = float('nan')
    x[valid:] = float('nan')
    n = torch.tensor([valid], dtype=torch.int32, device=DEV)
    ref = dy[:valid].float().t() @ x[:valid].float()
    dW = torch.full((No, Ki), 1.0, device=DEV)
    L.gemm(dy, x, dW, M=No, N=Ki, K=rows, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, split_k=split, rows_dev=n)
    assert bool(torch.isfinite(dW).all())
    assert relmax(dW - 1.0, ref) < 1e-4
    gb = torch.zeros(No, device=DEV)
    L.colsum_bf16(dy, gb, rows_dev=n)
    assert relmax(gb, dy[:valid].float().sum(0)) < 1e-4


GROUPS = {
    'text_layer': [(9920, 2304, 768, 6899), (9920, 768, 768, 6899), (9920, 3072, 768, 6899), (9920, 768, 3072, 6899)],
    'visual_layer': [(3520, 3072, 1024, 1903), (3520, 1024, 1024, 1903), (3520, 1024, 1024, None), (3520, 1024, 1024, 1903)],
    'co_block': [(9920, 1024, 768, 6848), (3520, 2048, 1024, 1903), (9920, 768, 1024, 6848), (3520, 1024, 1024, 1903),
                 (9920, 3072, 768, 6848), (9920, 768, 3072, 6848), (3520, 1024, 1024, 1903), (3520, 1024, 1024, 1903)],
    'ragged': [(992, 256, 128, 3), (1000, 576, 192, 617), (128, 128, 64, None), (37, 64, 72, 1)],
    'single': [(992, 768, 3072, 64)],
}


@pytest.mark.parametrize('name', sorted(GROUPS))
@pytest.mark.parametrize('split', [0, 3])
def test_gemm_wgrad_grouped_matches_separate_launches(name, split):
    """crct_gemm_wgrad_grouped: up to 8 weight-gradient problems of different shapes / device-side row counts in ONE launch give what
    the same problems give one by one (fp32 accumulation on top of existing contents, NaN rows past the count contribute nothing)."""
    torch.manual_seed(len(name))
    probs, keep = [], []
    for i, (rows, No, Ki, valid) in enumerate(GROUPS[name]):
        dy, x = bf(torch.randn(rows, No, device=DEV) * 0.5), bf(torch.randn(rows, Ki, device=DEV) * 0.5)
        n = None
        if valid is not None:
            dy[valid:] = float('nan')
            x[valid:] = float('nan')
            n = torch.tensor([valid], dtype=torch.int32, device=DEV)
            n.hint = valid
        v = rows if valid is None else valid
        ref = dy[:v].float().t() @ x[:v].float()
        dW = torch.full((No, Ki), float(i + 1), device=DEV)
        probs.append(L.gemm_args(dy, x, dW, M=No, N=Ki, K=rows, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, split_k=split,
                                 lda=No, ldb=Ki, ldd=Ki, rows_dev=n))
        keep.append((dy, x, n, dW, ref, i + 1))
    L.gemm_wgrad_grouped(probs)
    for dy, x, n, dW, ref, base in keep:
        assert bool(torch.isfinite(dW).all())
        assert relmax(dW - base, ref) < 1e-4


def test_gemm_wgrad_grouped_rejects_other_forms():
    dy, x = bf(torch.randn(256, 128, device=DEV)), bf(torch.randn(256, 64, device=DEV))
    dW = torch.zeros(128, 64, device=DEV)
    ok = L.gemm_args(dy, x, dW, M=128, N=64, K=256, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, lda=128, ldb=64, ldd=64)
    with pytest.raises(L.CrctError):
        L.gemm_wgrad_grouped([ok] * 9)
    bad = L.gemm_args(dy, x, dW, M=128, N=64, K=256, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=0, lda=128, ldb=64, ldd=64)
    with pytest.raises(L.CrctError):
        L.gemm_wgrad_grouped([bad])


@pytest.mark.parametrize('rows,H,valid', [(9920, 768, 6899), (3520, 1024, 1903), (37, 192, 5)])
def test_layernorm_fp32_z_and_device_row_count(rows, H, valid):
    torch.manual_seed(rows)
    z = torch.randn(rows, H, device=DEV) * 2 + 0.3
    z[valid:] = float('nan')
    gamma, beta = torch.randn(H, device=DEV), torch.randn(H, device=DEV)
    n = torch.tensor([valid], dtype=torch.int32, device=DEV)
    y = torch.full((rows, H), 7.0, device=DEV, dtype=torch.bfloat16)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    y32 = torch.full((rows, H), 7.0, device=DEV)
    L.layernorm_fwd(z, gamma, beta, y, mean, rstd, rows_dev=n, y32=y32)
    assert torch.equal(bf(y32[:valid]), y[:valid]) and bool((y32[valid:] == 7.0).all())
    zz = z[:valid].double()
    mu, var = zz.mean(1, keepdim=True), zz.var(1, unbiased=False, keepdim=True)
    ref = (zz - mu) / torch.sqrt(var + 1e-12) * gamma.double() + beta.double()
    assert relmax(y[:valid].float(), ref) < 6e-3 and bool((y[valid:] == 7.0).all())
    assert relmax(mean[:valid], mu.flatten()) < 1e-5
    # backward, split form, against autograd in fp64
    dy = bf(torch.randn(rows, H, device=DEV))
    dy[valid:] = float('nan')
    dz = torch.full((rows, H), 7.0, device=DEV, dtype=torch.bfloat16)
    L.layernorm_bwd(dy, z, mean, rstd, gamma, dz, rows_dev=n)
    dg, db, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    L.layernorm_bwd_params(dy, z, mean, rstd, dz, dg, db, dbias=dbias, rows_dev=n)
    zr = zz.clone().requires_grad_(True)
    gr = gamma.double().clone().requires_grad_(True)
    out = (zr - zr.mean(1, keepdim=True)) / torch.sqrt(zr.var(1, unbiased=False, keepdim=True) + 1e-12) * gr + beta.double()
    out.backward(dy[:valid].double())
    assert relmax(dz[:valid].float(), zr.grad) < 8e-3 and bool((dz[valid:] == 7.0).all())
    assert relmax(dg, gr.grad) < 2e-3
    assert relmax(db, dy[:valid].double().sum(0)) < 1e-4
    assert relmax(dbias, dz[:valid].double().sum(0)) < 1e-4


@pytest.mark.parametrize('B,nh,dh,Lq,Lk', [(5, 16, 48, 124, 124), (4, 16, 64, 44, 44), (6, 32, 32, 124, 44), (6, 32, 32, 44, 124), (3, 4, 48, 248, 248)])
@pytest.mark.parametrize('p', [0.0, 0.1])
def test_attention_packed_rows_equal_padded_rows(B, nh, dh, Lq, Lk, p):
    """cu_q / cu_k: sample b attends over its packed rows only == the padded call with the additive mask, on the valid rows
    (forward bit for bit without dropout; the dropout counters are laid out by (b, h, i, j) in both forms)."""
    torch.manual_seed(B * 7 + dh)
    H = nh * dh
    self_att = Lq == Lk
    g = torch.Generator().manual_seed(11)
    lq = torch.randint(1, Lq + 1, (B,), generator=g)
    lk = lq if self_att else torch.randint(1, Lk + 1, (B,), generator=g)
    cuq = torch.zeros(B + 1, dtype=torch.int32); cuq[1:] = lq.cumsum(0)
    cuk = torch.zeros(B + 1, dtype=torch.int32); cuk[1:] = lk.cumsum(0)
    q_pad, k_pad, v_pad = (bf(torch.randn(B * L_, H, device=DEV)) for L_ in (Lq, Lk, Lk))
    do_pad = bf(torch.randn(B * Lq, H, device=DEV))
    mask = torch.zeros(B, Lk, device=DEV)
    for b in range(B):
        mask[b, int(lk[b]):] = -10000.0
        do_pad[b * Lq + int(lq[b]):(b + 1) * Lq] = 0          # padded queries carry no gradient in the model

    def pack(x, lens, Lmax):
        return torch.cat([x[b * Lmax:b * Lmax + int(lens[b])] for b in range(B)] + [torch.full((B * Lmax - int(lens.sum()), H), float('nan'), device=DEV, dtype=x.dtype)])

    q_pk, k_pk, v_pk, do_pk = pack(q_pad, lq, Lq), pack(k_pad, lk, Lk), pack(v_pad, lk, Lk), pack(do_pad, lq, Lq)
    outs = {}
    for name, (q, k, v, do, m, cq, ck) in {'pad': (q_pad, k_pad, v_pad, do_pad, mask, None, None),
                                           'pk': (q_pk, k_pk, v_pk, do_pk, None, cuq.to(DEV), cuk.to(DEV))}.items():
        o = torch.zeros(B * Lq, H, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, nh, Lq, device=DEV)
        L.attn_fwd(q, k, v, m, o, lse, B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H, dropout_p=p, seed=5, cu_q=cq, cu_k=ck)
        dq, dk, dv = (torch.zeros(B * L_, H, device=DEV, dtype=torch.bfloat16) for L_ in (Lq, Lk, Lk))
        L.attn_bwd(q, k, v, m, o, do, lse, dq, dk, dv, B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H, lddo=H, lddq=H,
                   lddk=H, lddv=H, dropout_p=p, seed=5, cu_q=cq, cu_k=ck)
        outs[name] = (o, dq, dk, dv)
    for b in range(B):
        nq, nk = int(lq[b]), int(lk[b])
        sl_q_pad, sl_q_pk = slice(b * Lq, b * Lq + nq), slice(int(cuq[b]), int(cuq[b + 1]))
        sl_k_pad, sl_k_pk = slice(b * Lk, b * Lk + nk), slice(int(cuk[b]), int(cuk[b + 1]))
        assert torch.equal(outs['pad'][0][sl_q_pad], outs['pk'][0][sl_q_pk]), b
        for i, (sp, sk) in ((1, (sl_q_pad, sl_q_pk)), (2, (sl_k_pad, sl_k_pk)), (3, (sl_k_pad, sl_k_pk))):
            a, c = outs['pad'][i][sp].float(), outs['pk'][i][sk].float()
            assert bool(torch.isfinite(c).all())
            assert relmax(c, a) < 4e-3, (b, i)         # at most one bf16 rounding apart (zero terms summed in another order)


def build(name, varlen, **over):
    rec = load_golden(name)
    cfg_path, cfg, sd, batch = golden_inputs(rec)
    params = default_params(cfg_path, device='cuda', max_seq_len=rec['T'], max_vis_features=rec['R'], L1=rec['l1'], varlen=varlen, **over)
    m = VisualDialogEncoder(params)
    m.load_state_dict({'bert_pretrained.' + k: v for k, v in sd.items()}, strict=True)
    m.to('cuda').eval()
    return rec, m, params, {k: v.to('cuda') for k, v in batch.items()}


@pytest.mark.parametrize('name', ['tiny_eval', 'full_eval_b8_mild'])
def test_packed_forward_is_bitwise_the_padded_forward(name):
    outs = []
    for varlen in (True, False):
        rec, m, params, gb = build(name, varlen)
        assert m.varlen == varlen
        with torch.no_grad():
            _, _, _, _, scores, reg = glue_forward(m, gb, params, evaluation=True)
        outs.append((scores.clone(), reg[0].clone(), reg[2].clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b), float((a - b).abs().max())


@pytest.mark.parametrize('name', ['tiny_train_l1', 'tiny_ragged', 'full_train_b4_mild'])
def test_packed_gradients_match_padded_gradients(name):
    """Same bf16 values on both paths; only fp32 summation order differs (k-block grouping of the weight-gradient GEMMs,
    atomics): global relative L2 <= 2e-4, every tensor <= 1e-3 of the largest gradient norm."""
    grads, losses = [], []
    for varlen in (True, False):
        rec, m, params, gb = build(name, varlen)
        m.zero_grad()
        loss = glue_forward(m, gb, params)[0]
        loss.backward()
        torch.cuda.synchronize()
        losses.append(float(loss))
        grads.append({k: p.grad.double().clone() for k, p in m.bert_pretrained.named_parameters() if p.grad is not None})
    assert losses[0] == losses[1]
    num = sum(float((grads[0][k] - v).norm() ** 2) for k, v in grads[1].items())
    den = sum(float(v.norm() ** 2) for v in grads[1].values())
    gmax = max(float(v.norm()) for v in grads[1].values())
    assert (num / den) ** 0.5 < 2e-4, (num / den) ** 0.5
    for k, v in grads[1].items():
        assert float((grads[0][k] - v).norm()) <= 1e-3 * gmax, k


def test_arbitrary_masks_pack_exactly():
    """The public forward accepts any 0/1 attention masks (not only prefixes): packing compacts whichever rows are on."""
    rec, m, params, gb = build('tiny_eval', True)
    _, m2, params2, _ = build('tiny_eval', False)
    B, T = gb['tokens'].shape
    R = gb['image_feat'].shape[1]
    g = torch.Generator().manual_seed(3)
    am = (torch.rand(B, T, generator=g) < 0.6).to('cuda')
    am[:, 0] = True
    im = (torch.rand(B, R, generator=g) < 0.6).long().to('cuda')
    im[:, 0] = 1
    outs = []
    for model in (m, m2):
        with torch.no_grad():
            out = model(gb['tokens'], gb['loc'], gb['image_feat'], gb['image_loc'], token_type_ids=gb['segments'], attention_mask=am,
                        image_attention_mask=im, image_target=gb['image_target'], gt_reg=[gb['R'], 'L1'])
        outs.append((out[3].clone(), out[4][0].clone()))
    # with holes the packed keys sit at other positions of the attention tiles than the padded ones: same terms, another fp32
    # summation order (bit-identical only for prefix masks, test above)
    assert relmax(outs[0][0], outs[1][0]) < 1e-2 and relmax(outs[0][1], outs[1][1]) < 1e-3


@pytest.mark.parametrize('name', ['tiny_train_l1', 'full_train_b4_mild'])
def test_packed_training_step_with_dropout_draws_the_padded_masks(name):
    """Dropout ON: the counters of a packed row are those of its padded position (drop_rows = src_row; attention counters are
    (b, h, i, j) in both layouts), so a packed and a padded training pass draw identical masks and agree like the
    dropout-free passes do."""
    grads, losses = [], []
    for varlen in (True, False):
        torch.manual_seed(1234)
        rec, m, params, gb = build(name, varlen)
        m.train()
        m.zero_grad()
        loss = glue_forward(m, gb, params)[0]
        loss.backward()
        torch.cuda.synchronize()
        losses.append(float(loss.detach()))
        grads.append(m.arena.g32[:m.arena.live_end].double().clone())
    assert abs(losses[0] - losses[1]) < 1e-6, losses
    assert float((grads[0] - grads[1]).norm() / grads[1].norm()) < 2e-4


@pytest.mark.parametrize('M,N,K', [(200, 192, 192), (1000, 768, 3072), (6899, 768, 768), (1903, 1024, 1024), (9920, 2304, 768), (1000, 200, 128)])
def test_gemm_192_wide_tiles_all_forms(M, N, K):
    """block_n = 192 (single-CTA tile; the tile shape the policy picks for ~1.1-wave problems on 256-wide tiles)."""
    torch.manual_seed(M + N)
    A, B = bf(torch.randn(M, K, device=DEV) * 0.5), bf(torch.randn(N, K, device=DEV) * 0.5)
    bias, aux = torch.randn(N, device=DEV), bf(torch.randn(M, N, device=DEV))
    ref = A.float() @ B.float().t() + bias
    D, D2 = torch.empty(M, N, device=DEV, dtype=torch.bfloat16), torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, block_n=192)
    assert relmax(D.float(), ref) < 6e-3
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_GELU, D2=D2, block_n=192)
    assert relmax(D.float(), torch.nn.functional.gelu(ref)) < 6e-3
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES, aux=aux, block_n=192, dropout_p=0.1, seed=3)
    L.gemm(A, B, D2, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES, aux=aux, block_n=256, dropout_p=0.1, seed=3)
    assert torch.equal(D, D2)                     # same accumulation order per element, same dropout counters: tile shape is invisible
    Z = torch.empty(M, N, device=DEV)
    L.gemm(A, B, Z, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES_F32, aux=aux.float(), block_n=192)
    assert relmax(Z, ref + aux.float()) < 1e-5
    W = bf(torch.randn(K, N, device=DEV) * 0.5)   # dgrad form (MN-major B)
    L.gemm(A, W, D, M=M, N=N, K=K, b_major=1, epilogue=L.EPI_MUL, aux=aux, block_n=192)
    assert relmax(D.float(), (A.float() @ W.float()) * aux.float()) < 6e-3
    dW = torch.zeros(K, N, device=DEV)            # wgrad form: dW[K,N] += A^T aux  (both MN-major), GEMM-K = M rows
    L.gemm(A, aux, dW, M=K, N=N, K=M, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, block_n=192)
    assert relmax(dW, A.float().t() @ aux.float()) < 1e-4
    n = torch.tensor([max(1, M // 2)], dtype=torch.int32, device=DEV)          # hint-driven policy with a device-side count
    n.hint = int(n)
    D.fill_(7.0)
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, rows_dev=n)
    assert relmax(D[:int(n)].float(), ref[:int(n)]) < 6e-3 and bool((D[int(n):] == 7.0).all())
