"""Per-kernel parity on the B200: every C-ABI entry point against a plain torch fp32 statement of the
same op on the same (bf16-rounded) inputs.  Tolerances: a bf16 output carries one rounding (2^-9 relative);
fp32 outputs are compared at 1e-4 relative to the tensor scale."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import _lib as L   # noqa: E402
from oracle import crct_oracle as O   # noqa: E402

DEV = 'cuda'


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


def bf(x):
    return x.to(torch.bfloat16)


@pytest.fixture(autouse=True, scope='module')
def _device():
    L.device_check()


GEMM_SHAPES = [(128, 128, 64), (200, 192, 192), (992, 576, 192), (1000, 768, 3072), (9920, 2304, 768), (3520, 1024, 1024)]


@pytest.mark.parametrize('M,N,K', GEMM_SHAPES)
@pytest.mark.parametrize('bn', [0, 128, 256])
@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_forward_epilogues(M, N, K, bn, cg):
    torch.manual_seed(M + N + K)
    A, B = bf(torch.randn(M, K, device=DEV) * 0.5), bf(torch.randn(N, K, device=DEV) * 0.5)
    bias, aux = torch.randn(N, device=DEV), bf(torch.randn(M, N, device=DEV))
    ref = A.float() @ B.float().t() + bias
    D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, block_n=bn, cta_group=cg)
    assert relmax(D.float(), ref) < 6e-3
    D2 = torch.empty_like(D)
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_GELU, D2=D2, block_n=bn, cta_group=cg)
    assert relmax(D2.float(), O.gelu_grad(ref)) < 6e-3          # saved derivative for the backward
    assert relmax(D.float(), torch.nn.functional.gelu(ref)) < 6e-3
    D3 = torch.empty_like(D)
    L.gemm(A, B, D3, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_GELU, block_n=bn, cta_group=cg)      # eval form: no D2
    assert relmax(D3.float(), torch.nn.functional.gelu(ref)) < 6e-3
    L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, epilogue=L.EPI_BIAS_RES, aux=aux, block_n=bn, cta_group=cg)
    assert relmax(D.float(), ref + aux.float()) < 6e-3


@pytest.mark.parametrize('cg', [1, 2])
@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (1000, 192, 576), (9920, 768, 2304), (3520, 1024, 3072)])
def test_gemm_dgrad_forms(M, N, K, cg):
    """dx = dy W with W stored [K_gemm, N_gemm] = [out, in] (b_major = 1)."""
    torch.manual_seed(1)
    dy, W = bf(torch.randn(M, K, device=DEV) * 0.5), bf(torch.randn(K, N, device=DEV) * 0.5)
    aux = bf(torch.randn(M, N, device=DEV))
    ref = dy.float() @ W.float()
    D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.gemm(dy, W, D, M=M, N=N, K=K, b_major=1, cta_group=cg)
    assert relmax(D.float(), ref) < 6e-3
    L.gemm(dy, W, D, M=M, N=N, K=K, b_major=1, epilogue=L.EPI_BIAS_RES, aux=aux, cta_group=cg)
    assert relmax(D.float(), ref + aux.float()) < 6e-3
    L.gemm(dy, W, D, M=M, N=N, K=K, b_major=1, epilogue=L.EPI_MUL, aux=aux, cta_group=cg)
    assert relmax(D.float(), ref * aux.float()) < 6e-3


@pytest.mark.parametrize('rows,No,Ki', [(128, 128, 64), (1000, 576, 192), (9920, 768, 768), (9920, 3072, 768), (3520, 1024, 1024), (992, 768, 3072)])
@pytest.mark.parametrize('split', [0, 1, 5])
@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_wgrad_accumulates_fp32(rows, No, Ki, split, cg):
    """dW[out,in] += dy^T x with both operands MN-major; split-K partial sums meet through fp32 atomics."""
    torch.manual_seed(2)
    dy, x = bf(torch.randn(rows, No, device=DEV) * 0.5), bf(torch.randn(rows, Ki, device=DEV) * 0.5)
    ref = dy.float().t() @ x.float()
    dW = torch.full((No, Ki), 1.0, device=DEV)
    L.gemm(dy, x, dW, M=No, N=Ki, K=rows, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, split_k=split, cta_group=cg)
    assert relmax(dW - 1.0, ref) < 1e-4


def test_gemm_persistent_schedule_and_dropout():
    """max_ctas forces several tiles per CTA (smem ring + TMEM double buffer wrap); dropout keeps ~1-p and rescales."""
    torch.manual_seed(3)
    M, N, K = 1024, 1024, 512
    A, B = bf(torch.randn(M, K, device=DEV) * 0.5), bf(torch.randn(N, K, device=DEV) * 0.5)
    ref = A.float() @ B.float().t()
    for bn in (128, 256):
        D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        L.gemm(A, B, D, M=M, N=N, K=K, block_n=bn, max_ctas=3)
        assert relmax(D.float(), ref) < 6e-3
        L.gemm(A, B, D, M=M, N=N, K=K, block_n=bn, max_ctas=4, cta_group=2)       # 2 clusters, several tiles each
        assert relmax(D.float(), ref) < 6e-3
    D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, B, D, M=M, N=N, K=K, epilogue=L.EPI_BIAS_RES, dropout_p=0.25, seed=77)
    kept = D.float() != 0
    assert abs(float(kept.float().mean()) - 0.75) < 0.01
    assert relmax(D.float()[kept], (ref / 0.75)[kept]) < 6e-3
    D2 = torch.empty_like(D)
    L.gemm(A, B, D2, M=M, N=N, K=K, epilogue=L.EPI_BIAS_RES, dropout_p=0.25, seed=77)
    assert torch.equal(D, D2)                       # counter-based: same seed, same mask


def test_gemm_rejects_bad_arguments():
    A = torch.zeros(128, 64, device=DEV, dtype=torch.bfloat16)
    D = torch.zeros(128, 132, device=DEV, dtype=torch.bfloat16)
    with pytest.raises(L.CrctError):
        L.gemm(A, A, D, M=128, N=132, K=64, ldd=132)          # N % 8
    with pytest.raises(L.CrctError):
        L.gemm(A, A, D, M=128, N=128, K=64, split_k=2)        # split-K needs the fp32 accumulate epilogue


@pytest.mark.parametrize('rows,H', [(9920, 768), (3520, 1024), (37, 192), (5, 128)])
def test_layernorm_fwd_bwd(rows, H):
    torch.manual_seed(4)
    z = bf(torch.randn(rows, H, device=DEV) * 2 + 0.3)
    gamma, beta = 1 + 0.1 * torch.randn(H, device=DEV), 0.1 * torch.randn(H, device=DEV)
    y = torch.empty_like(z)
    mean, rstd = torch.empty(rows, device=DEV), torch.empty(rows, device=DEV)
    L.layernorm_fwd(z, gamma, beta, y, mean, rstd)
    yr, mr, rr = O.ln_fwd(z.float(), gamma, beta)
    assert relmax(y.float(), yr) < 6e-3
    assert relmax(mean, mr.squeeze(-1)) < 1e-4 and relmax(rstd, rr.squeeze(-1)) < 1e-4
    dy = bf(torch.randn(rows, H, device=DEV))
    dz = torch.empty_like(z)
    dg, db, dbias = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    L.layernorm_bwd(dy, z, mean, rstd, gamma, dz, dg, db, dbias=dbias)
    dzr, dgr, dbr = O.ln_bwd(dy.float(), z.float(), mr, rr, gamma)
    assert relmax(dz.float(), dzr) < 6e-3
    assert relmax(dg, dgr) < 2e-4 and relmax(db, dbr) < 2e-4
    assert relmax(dbias, dzr.sum(0)) < 2e-4            # summed in fp32 before the bf16 store
    # with the residual-path dropout re-applied to the dense-output gradient
    dzm = torch.empty_like(z)
    dbias.zero_(); dg.zero_(); db.zero_()
    L.layernorm_bwd(dy, z, mean, rstd, gamma, dz, dg, db, dbias=dbias, dzm=dzm, p_out=0.1, seed_out=5)
    kept = dzm.float() != 0
    assert abs(float(kept.float().mean()) - 0.9) < max(0.02, 4.5 * (0.09 / kept.numel()) ** 0.5)      # 4.5 sigma of the keep-rate estimate
    assert relmax(dzm.float()[kept], (dz.float() / 0.9)[kept]) < 1e-2
    assert relmax(dbias, dzm.float().sum(0)) < 3e-3            # reference built from bf16-rounded dzm
    # split form (what the training step runs): dz/dzm alone, then the three column sums as a second launch
    dz2, dzm2 = torch.empty_like(z), torch.empty_like(z)
    dg2, db2, dbias2 = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
    L.layernorm_bwd(dy, z, mean, rstd, gamma, dz2, dzm=dzm2, p_out=0.1, seed_out=5)
    L.layernorm_bwd_params(dy, z, mean, rstd, dz2, dg2, db2, dbias=dbias2, dzm=dzm2, p_out=0.1)
    assert torch.equal(dz2, dz) and torch.equal(dzm2, dzm)
    assert relmax(dg2, dgr) < 2e-4 and relmax(db2, dbr) < 2e-4
    assert relmax(dbias2, dzm.float().sum(0)) < 2e-4
    dbias2.zero_()
    L.layernorm_bwd(dy, z, mean, rstd, gamma, dz2)            # no dropout: the dense-bias gradient sums dz itself
    L.layernorm_bwd_params(dy, z, mean, rstd, dz2, None, None, dbias=dbias2)
    assert relmax(dbias2, dz2.float().sum(0)) < 2e-4


def test_layernorm_bwd_mask_matches_gemm_epilogue_mask():
    """The dense-output gradient must be masked with exactly the mask the forward epilogue drew."""
    torch.manual_seed(5)
    M, N, K = 256, 192, 64
    A, B = bf(torch.randn(M, K, device=DEV)), bf(torch.randn(N, K, device=DEV))
    D = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    L.gemm(A, B, D, M=M, N=N, K=K, epilogue=L.EPI_BIAS_RES, dropout_p=0.3, seed=123)
    z = bf(torch.randn(M, N, device=DEV))
    mean, rstd, y = torch.empty(M, device=DEV), torch.empty(M, device=DEV), torch.empty_like(z)
    g = torch.ones(N, device=DEV)
    L.layernorm_fwd(z, g, torch.zeros(N, device=DEV), y, mean, rstd)
    dz, dzm = torch.empty_like(z), torch.empty_like(z)
    L.layernorm_bwd(bf(torch.randn(M, N, device=DEV)), z, mean, rstd, g, dz, torch.zeros(N, device=DEV), torch.zeros(N, device=DEV),
                    dzm=dzm, p_out=0.3, seed_out=123)
    both = (dz.float() != 0) & (A.float() @ B.float().t()).abs().gt(1e-3)
    assert torch.equal((D.float() != 0)[both], (dzm.float() != 0)[both])


def test_colsum_softmax_cast_mask():
    torch.manual_seed(6)
    x = bf(torch.randn(9920, 2304, device=DEV))
    out = torch.ones(2304, device=DEV)
    L.colsum_bf16(x, out)
    assert relmax(out - 1, x.float().sum(0)) < 2e-4
    f = torch.relu(torch.randn(3520, 1024, device=DEV)) * 1.5
    f[7] = 0
    p = torch.empty(3520, 1024, device=DEV, dtype=torch.bfloat16)
    L.softmax_rows(f, p)
    assert relmax(p.float(), torch.softmax(f, -1)) < 6e-3
    w = torch.randn(64 * 1001, device=DEV)
    w16 = torch.empty_like(w, dtype=torch.bfloat16)
    L.cast_f32_to_bf16(w, w16)
    assert torch.equal(w16, w.to(torch.bfloat16))
    for m in (torch.rand(80, 124, device=DEV) > 0.3, (torch.rand(80, 44, device=DEV) > 0.3).long()):
        o = torch.empty(m.shape, device=DEV)
        L.additive_mask(m, o)
        assert torch.equal(o, (1.0 - m.float()) * -10000.0)


ATT = [  # B, nh, dh, Lq, Lk   (text self, visual self, co-attention both ways, ragged, stress)
    (8, 16, 48, 124, 124), (8, 16, 64, 44, 44), (8, 32, 32, 124, 44), (8, 32, 32, 44, 124), (3, 4, 48, 19, 19),
    (2, 2, 64, 5, 5), (3, 4, 32, 19, 5), (2, 16, 48, 248, 248), (2, 32, 32, 88, 248)]


def _attn_case(B, nh, dh, Lq, Lk, seed=0):
    torch.manual_seed(seed)
    H = nh * dh
    q, k, v = (bf(torch.randn(B, L_, H, device=DEV)) for L_ in (Lq, Lk, Lk))
    valid = torch.arange(Lk, device=DEV)[None, :] < torch.randint(max(1, Lk // 3), Lk + 1, (B, 1), device=DEV)
    mask = (1.0 - valid.float()) * -10000.0
    return H, q, k, v, mask


@pytest.mark.parametrize('B,nh,dh,Lq,Lk', ATT)
def test_attention_forward(B, nh, dh, Lq, Lk):
    H, q, k, v, mask = _attn_case(B, nh, dh, Lq, Lk)
    out = torch.empty(B * Lq, H, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, nh, Lq, device=DEV)
    L.attn_fwd(q, k, v, mask, out, lse, B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H)
    ctx, p, _ = O.attn_fwd(q.float(), k.float(), v.float(), mask, nh)
    assert relmax(out.float().view(B, Lq, H), ctx) < 1e-2
    s = O.split_heads(q.float(), B, Lq, nh) @ O.split_heads(k.float(), B, Lk, nh).transpose(-1, -2) / math.sqrt(dh) + mask[:, None, None, :]
    assert relmax(lse, torch.logsumexp(s, -1)) < 1e-4


@pytest.mark.parametrize('B,nh,dh,Lq,Lk', ATT)
def test_attention_backward(B, nh, dh, Lq, Lk):
    H, q, k, v, mask = _attn_case(B, nh, dh, Lq, Lk, seed=1)
    out = torch.empty(B * Lq, H, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, nh, Lq, device=DEV)
    kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H)
    L.attn_fwd(q, k, v, mask, out, lse, **kw)
    dout = bf(torch.randn(B * Lq, H, device=DEV))
    dq, dk, dv = torch.full_like(q, 9.0), torch.full_like(k, 9.0), torch.full_like(v, 9.0)
    L.attn_bwd(q, k, v, mask, out, dout, lse, dq, dk, dv, lddo=H, lddq=H, lddk=H, lddv=H, **kw)
    ctx, p, pd = O.attn_fwd(q.float(), k.float(), v.float(), mask, nh)
    rq, rk, rv = O.attn_bwd(dout.float().view(B, Lq, H), q.float(), k.float(), v.float(), p, pd, nh)
    assert relmax(dq.float(), rq) < 2e-2
    assert relmax(dk.float(), rk) < 2e-2
    assert relmax(dv.float(), rv) < 2e-2


def test_attention_packed_qkv_strides_and_dropout():
    """q/k/v as column slices of one packed [rows, 3H] projection (how the engine calls it); dropout statistics and
    forward/backward mask agreement (dV of a kept-everything V = column sums of the dropped probabilities)."""
    torch.manual_seed(7)
    B, nh, dh, Lt = 4, 16, 48, 124
    H = nh * dh
    qkv = bf(torch.randn(B * Lt, 3 * H, device=DEV))
    mask = torch.zeros(B, Lt, device=DEV)
    out = torch.empty(B * Lt, H, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, nh, Lt, device=DEV)
    kw = dict(B=B, nh=nh, dh=dh, Lq=Lt, Lk=Lt, ldq=3 * H, ldk=3 * H, ldv=3 * H, ldo=H)
    L.attn_fwd(qkv, qkv[:, H:], qkv[:, 2 * H:], mask, out, lse, **kw)
    q, k, v = (qkv[:, i * H:(i + 1) * H].float().reshape(B, Lt, H) for i in range(3))
    ctx, p, _ = O.attn_fwd(q, k, v, mask, nh)
    assert relmax(out.float().view(B, Lt, H), ctx) < 1e-2
    # dropout: E[out_dropped] = out; compare forward with p=0.5 over many heads statistically
    outd = torch.empty_like(out)
    L.attn_fwd(qkv, qkv[:, H:], qkv[:, 2 * H:], mask, outd, lse, dropout_p=0.5, seed=11, **kw)
    assert abs(float(outd.float().mean() - out.float().mean())) < 0.02
    assert float((outd.float() - out.float()).abs().mean()) > 1e-3
    dqkv = torch.zeros_like(qkv)
    dout = bf(torch.randn(B * Lt, H, device=DEV))
    L.attn_bwd(qkv, qkv[:, H:], qkv[:, 2 * H:], mask, outd, dout, lse, dqkv, dqkv[:, H:], dqkv[:, 2 * H:], lddo=H,
               lddq=3 * H, lddk=3 * H, lddv=3 * H, dropout_p=0.5, seed=11, **kw)
    # recover the mask from the forward: out_dropped = (P*mask/0.5) V  -> check dV = (P*mask/0.5)^T dO with the same mask
    # by solving for it on one head through V = identity-like probe is overkill; instead check linearity of dV in dO:
    dqkv2 = torch.zeros_like(qkv)
    L.attn_bwd(qkv, qkv[:, H:], qkv[:, 2 * H:], mask, outd, bf(dout.float() * 2), lse, dqkv2, dqkv2[:, H:], dqkv2[:, 2 * H:],
               lddo=H, lddq=3 * H, lddk=3 * H, lddv=3 * H, dropout_p=0.5, seed=11, **kw)
    assert relmax(dqkv2[:, 2 * H:].float(), 2 * dqkv[:, 2 * H:].float()) < 2e-2


def test_attention_dropout_mask_consistency_fwd_bwd():
    """With V = I-like probe columns the forward output IS the dropped probability matrix, and dV with dO = e_j picks
    its columns: both passes must have drawn the same mask."""
    torch.manual_seed(8)
    B, nh, dh, Lq, Lk = 1, 1, 32, 16, 32
    H = nh * dh
    q = bf(torch.randn(B, Lq, H, device=DEV)); k = bf(torch.randn(B, Lk, H, device=DEV))
    v = torch.eye(Lk, dh, device=DEV).to(torch.bfloat16).view(B, Lk, H).contiguous()      # V = I (Lk == dh)
    mask = torch.zeros(B, Lk, device=DEV)
    out = torch.empty(B * Lq, H, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, nh, Lq, device=DEV)
    kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=H, ldk=H, ldv=H, ldo=H)
    L.attn_fwd(q, k, v, mask, out, lse, dropout_p=0.4, seed=3, **kw)
    pd_fwd = out.float()                                    # [Lq, Lk] = dropped probabilities
    dout = torch.ones(B * Lq, H, device=DEV, dtype=torch.bfloat16)
    dq, dk, dv = torch.zeros_like(q), torch.zeros_like(k), torch.zeros_like(v)
    L.attn_bwd(q, k, v, mask, out, dout, lse, dq, dk, dv, lddo=H, lddq=H, lddk=H, lddv=H, dropout_p=0.4, seed=3, **kw)
    # dV[j, :] = sum_i pd[i, j] * dO[i, :] = column sums of pd (dO = 1)
    assert relmax(dv.float().view(Lk, dh)[:, 0], pd_fwd.sum(0)) < 2e-2


def test_embeddings_fwd_bwd():
    from cqa_crct_b200.synthetic import make_batch
    torch.manual_seed(9)
    B, T, R, H, Hv, V, P = 6, 32, 12, 192, 128, 2048, 64
    batch = {k: v.to(DEV) for k, v in make_batch(B, T, R, 128, seed=5, vocab_size=V).items()}
    w = {'bert.embeddings.word_embeddings.weight': torch.randn(V, H, device=DEV) * 0.05,
         'bert.embeddings.position_embeddings.weight': torch.randn(P, H, device=DEV) * 0.05,
         'bert.embeddings.plotqa_type_embeddings.weight': torch.randn(12, H, device=DEV) * 0.05,
         'bert.embeddings.txt_location_embeddings.weight': torch.randn(H, 4, device=DEV) * 0.05,
         'bert.embeddings.txt_location_embeddings.bias': torch.randn(H, device=DEV) * 0.05,
         'bert.embeddings.LayerNorm.weight': 1 + 0.1 * torch.randn(H, device=DEV),
         'bert.embeddings.LayerNorm.bias': 0.1 * torch.randn(H, device=DEV)}
    pre = 'bert.embeddings.'
    y = torch.empty(B * T, H, device=DEV, dtype=torch.bfloat16); z = torch.empty_like(y)
    mean, rstd = torch.empty(B * T, device=DEV), torch.empty(B * T, device=DEV)
    L.embed_text_fwd(batch['tokens'], batch['segments'], batch['loc'], w[pre + 'word_embeddings.weight'],
                     w[pre + 'position_embeddings.weight'], w[pre + 'plotqa_type_embeddings.weight'],
                     w[pre + 'txt_location_embeddings.weight'], w[pre + 'txt_location_embeddings.bias'],
                     w[pre + 'LayerNorm.weight'], w[pre + 'LayerNorm.bias'], y, z, mean, rstd)
    wc = {k: v.cpu() for k, v in w.items()}
    bc = {k: v.cpu() for k, v in batch.items()}
    yr, c = O.embed_text_fwd(wc, bc['tokens'], bc['segments'], bc['loc'])
    assert relmax(z.float().cpu().view(B, T, H), c['z']) < 6e-3
    assert relmax(y.float().cpu().view(B, T, H), yr) < 8e-3
    # backward: LN backward + scatter
    dy = bf(torch.randn(B * T, H, device=DEV))
    dz = torch.empty_like(y)
    g = {k: torch.zeros_like(v) for k, v in w.items()}
    L.layernorm_bwd(dy, z, mean, rstd, w[pre + 'LayerNorm.weight'], dz, g[pre + 'LayerNorm.weight'], g[pre + 'LayerNorm.bias'])
    L.embed_text_bwd(batch['tokens'], batch['segments'], batch['loc'], dz, g[pre + 'word_embeddings.weight'],
                     g[pre + 'position_embeddings.weight'], g[pre + 'plotqa_type_embeddings.weight'],
                     g[pre + 'txt_location_embeddings.weight'], g[pre + 'txt_location_embeddings.bias'])
    gr = O._Grads()
    O.embed_text_bwd(wc, gr, dy.float().cpu().view(B, T, H), c, bc['tokens'], bc['loc'])
    for kname in w:
        assert relmax(g[kname].cpu(), gr[kname]) < 2e-2, kname

    # visual side
    wv = {'new_loc_emb.weight': torch.randn(Hv, 4, device=DEV) * 0.05, 'new_loc_emb.bias': torch.randn(Hv, device=DEV) * 0.05,
          'color_emb.weight': torch.randn(229, Hv, device=DEV) * 0.05, 'LayerNorm.weight': 1 + 0.1 * torch.randn(Hv, device=DEV),
          'LayerNorm.bias': 0.1 * torch.randn(Hv, device=DEV)}
    gimg = bf(torch.randn(B * R, Hv, device=DEV) * 0.1)
    box, cls = batch['image_loc'].view(-1, 4).contiguous(), batch['image_target'].view(-1).contiguous()
    yv = torch.empty(B * R, Hv, device=DEV, dtype=torch.bfloat16); zv = torch.empty_like(yv)
    mv, rv = torch.empty(B * R, device=DEV), torch.empty(B * R, device=DEV)
    L.embed_vis_fwd(gimg, box, cls, wv['new_loc_emb.weight'], wv['new_loc_emb.bias'], wv['color_emb.weight'],
                    wv['LayerNorm.weight'], wv['LayerNorm.bias'], yv, zv, mv, rv)
    zr = gimg.float() + box @ wv['new_loc_emb.weight'].t() + wv['new_loc_emb.bias'] + wv['color_emb.weight'][cls]
    assert relmax(zv.float(), zr) < 6e-3
    yr2, _, _ = O.ln_fwd(zv.float(), wv['LayerNorm.weight'], wv['LayerNorm.bias'])
    assert relmax(yv.float(), yr2) < 8e-3
    dzv = bf(torch.randn(B * R, Hv, device=DEV))
    gc, gw = torch.zeros(229, Hv, device=DEV), torch.zeros(Hv, 4, device=DEV)
    L.embed_vis_bwd(dzv, box, cls, gc, gw)
    assert relmax(gw, dzv.float().t() @ box) < 2e-4
    assert relmax(gc, torch.zeros_like(gc).index_add_(0, cls, dzv.float())) < 2e-4


def test_heads_linear_loss_adamw():
    torch.manual_seed(10)
    M, N, K = 80, 1000, 520
    x, W, b = torch.randn(M, K, device=DEV), torch.randn(N, K, device=DEV) * 0.05, torch.randn(N, device=DEV)
    y = torch.empty(M, N, device=DEV)
    for act, fn in ((L.ACT_NONE, lambda t: t), (L.ACT_RELU, torch.relu), (L.ACT_LEAKY, lambda t: torch.nn.functional.leaky_relu(t, 0.01)),
                    (L.ACT_TANH, torch.tanh)):
        L.linear_f32(x, K, 1, W, 1, K, y, N, M, N, K, bias=b, act=act)
        assert relmax(y, fn(x @ W.t() + b)) < 1e-5
    dy = torch.randn(M, N, device=DEV)
    post = torch.randn(M, K, device=DEV)
    dx = torch.empty(M, K, device=DEV)
    L.linear_f32(dy, N, 1, W, K, 1, dx, K, M, K, N, dmask=post, ldm=K, slope=0.01)
    assert relmax(dx, (dy @ W) * torch.where(post > 0, 1.0, 0.01)) < 1e-5
    dW = torch.ones(N, K, device=DEV)
    L.linear_f32(dy, 1, N, x, K, 1, dW, K, N, K, M, accumulate=1)
    assert relmax(dW - 1, dy.t() @ x) < 1e-5
    db = torch.zeros(N, device=DEV)
    L.colsum_f32(dy, db, M, N, N)
    assert relmax(db, dy.sum(0)) < 1e-5
    # loss
    for l1, kind in ((True, 'L1_smooth'), (False, 'L1_smooth'), (True, 'L1')):
        B = 80
        logits, reg = torch.randn(B, 2, device=DEV), torch.tanh(torch.randn(B, device=DEV))
        labels = torch.randint(0, 2, (B,), device=DEV); labels[3] = -1
        R = torch.stack([torch.randn(B, device=DEV) * 50, (torch.rand(B, device=DEV) < 0.4).float(), torch.full((B,), 0.01, device=DEV),
                         torch.rand(B, device=DEV) * 60 + 1], 1).contiguous()
        outs = [torch.empty(B, device=DEV) for _ in range(4)]
        sc, dl, dp = torch.empty(5, device=DEV), torch.empty(B, 2, device=DEV), torch.empty(B, device=DEV)
        L.hybrid_loss(logits, reg, labels, R, *outs, sc, dl, dp, l1=l1, zero_impossible=(kind != 'L1'), tol_margin=0.01,
                      nsp_coeff=0.7, reg_coeff=1.3)
        ref = O.losses_fwd(logits.cpu(), reg.cpu(), labels.cpu(), R.cpu(), kind, l1, 0.01)
        for o, kname in zip(outs, ('reg_pred', 'reg_loss', 'reg_l1', 'reg_dist')):
            assert torch.allclose(o.cpu(), ref[kname], rtol=1e-5, atol=1e-6), kname
        assert abs(float(sc[1]) - float(ref['nsp_loss'])) < 1e-5
        assert abs(float(sc[0]) - float(0.7 * ref['nsp_loss'] + 1.3 * ref['reg_loss'].mean())) < 1e-5
        assert (int(sc[3]), int(sc[4])) == ref['reg_right']
        assert torch.allclose(dl.cpu(), 0.7 * ref['dlogits'], atol=1e-6)
        assert torch.allclose(dp.cpu(), 1.3 * ref['dreg'] * (1 - reg.cpu() ** 2), atol=1e-6)
    # AdamW against torch.optim.AdamW over two steps, two groups
    n = 64 * 50
    w0, g0 = torch.randn(n, device=DEV), torch.randn(n, device=DEV)
    pa, pb = torch.nn.Parameter(w0[:n // 2].clone()), torch.nn.Parameter(w0[n // 2:].clone())
    opt = torch.optim.AdamW([{'params': [pa], 'lr': 2e-3, 'weight_decay': 0.01}, {'params': [pb], 'lr': 1e-3, 'weight_decay': 0.0}])
    w, m, v = w0.clone(), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    w16 = torch.empty(n, device=DEV, dtype=torch.bfloat16)
    group = torch.zeros(n // 64, dtype=torch.uint8, device=DEV); group[n // 128:] = 1
    for step in (1, 2):
        pa.grad, pb.grad = g0[:n // 2].clone() * step, g0[n // 2:].clone() * step
        opt.step()
        L.adamw(w, g0 * step * 4, m, v, w16, group, n, [2e-3, 1e-3, 0, 0], [0.01, 0.0, 0, 0], 0.9, 0.999, 1e-8, step, grad_scale=0.25)
    assert relmax(w, torch.cat([pa.data, pb.data])) < 1e-5
    assert torch.equal(w16, w.to(torch.bfloat16))
    # first-token gather / scatter
    t = bf(torch.randn(4 * 7, 192, device=DEV))
    h0 = torch.empty(4, 192, device=DEV)
    L.gather_first(t, 7 * 192, h0)
    assert torch.equal(h0, t.view(4, 7, 192)[:, 0].float())
    dst = torch.zeros_like(t)
    L.scatter_first(h0, dst, 7 * 192)
    assert torch.equal(dst.view(4, 7, 192)[:, 0], t.view(4, 7, 192)[:, 0]) and float(dst.view(4, 7, 192)[:, 1:].abs().sum()) == 0
    pt, pv = torch.relu(torch.randn(80, 1024, device=DEV)), torch.relu(torch.randn(80, 1024, device=DEV))
    pooled = torch.empty_like(pt)
    L.pool_mul_fwd(pt, pv, pooled, 0.1, 9)
    kept = (pooled != 0) | (pt * pv == 0)
    assert abs(float(kept.float().mean()) - 0.9) < 0.35
    dpo = torch.randn_like(pt); dut, duv = torch.empty_like(pt), torch.empty_like(pt)
    L.pool_mul_bwd(dpo, pt, pv, dut, duv, 0.0, 0)
    assert relmax(dut, dpo * pv * (pt > 0)) < 1e-6 and relmax(duv, dpo * pt * (pv > 0)) < 1e-6


def test_adamw_vector_form_equals_scalar_form_and_shard_offsets():
    """The 128-bit AdamW kernel (4 consecutive elements per thread) gives the bits of the scalar kernel; a range that starts inside a
    64-element block (`group_offset`, a data-parallel rank's shard) picks up the right (lr, weight decay) group in both forms."""
    torch.manual_seed(5)
    n = 64 * 40
    w0, g0 = torch.randn(n + 8, device=DEV), torch.randn(n + 8, device=DEV)
    m0, v0 = torch.rand(n + 8, device=DEV) * 0.1, torch.rand(n + 8, device=DEV) * 0.01
    group = (torch.arange(n // 64 + 1, device=DEV) % 4).to(torch.uint8)
    lr4, wd4 = [2e-3, 1e-3, 5e-4, 3e-3], [0.01, 0.0, 0.02, 0.0]

    def run(lo, hi, shift):
        """update elements [lo, hi) of a copy whose storage is shifted by `shift` elements (shift = 1: misaligned -> scalar kernel)"""
        w, g, m, v = (torch.empty(n + 8, device=DEV) for _ in range(4))
        w16 = torch.zeros(n + 8, device=DEV, dtype=torch.bfloat16)
        for dst, src in ((w, w0), (g, g0), (m, m0), (v, v0)):
            dst[shift:shift + n].copy_(src[:n])
        s = slice(shift + lo, shift + hi)
        L.adamw(w[s], g[s], m[s], v[s], w16[s] if shift == 0 else None, group[lo // 64:], hi - lo, lr4, wd4, 0.9, 0.999, 1e-8, 3, group_offset=lo % 64)
        return w[shift:shift + n].clone(), m[shift:shift + n].clone(), v[shift:shift + n].clone(), w16[:n].clone()

    for lo, hi in ((0, n), (64 * 3 + 8, 64 * 17 + 24), (64 * 5 + 60, 64 * 5 + 64)):
        wa, ma, va, w16a = run(lo, hi, 0)          # aligned: vector kernel
        wb, mb, vb, _ = run(lo, hi, 1)             # storage shifted by one float: scalar kernel
        assert torch.equal(wa, wb) and torch.equal(ma, mb) and torch.equal(va, vb)
        assert torch.equal(wa[:lo], w0[:lo]) and torch.equal(wa[hi:], w0[hi:n])       # nothing outside the range is touched
        assert not torch.equal(wa[lo:hi], w0[lo:hi])
        assert torch.equal(w16a[lo:hi], wa[lo:hi].to(torch.bfloat16))
        # reference arithmetic (torch.optim.AdamW semantics) with the per-block groups
        gi = group[(torch.arange(lo, hi, device=DEV) // 64)].long() % 4
        lr, wd = torch.tensor(lr4, device=DEV)[gi], torch.tensor(wd4, device=DEV)[gi]
        g, m, v, w = g0[lo:hi], m0[lo:hi], v0[lo:hi], w0[lo:hi]
        m1, v1 = 0.9 * m + 0.1 * g, 0.999 * v + 0.001 * g * g
        ref = w * (1 - lr * wd) - lr / (1 - 0.9 ** 3) * m1 / (v1.sqrt() / (1 - 0.999 ** 3) ** 0.5 + 1e-8)
        assert relmax(wa[lo:hi], ref) < 1e-5
