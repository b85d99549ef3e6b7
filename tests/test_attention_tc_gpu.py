"""tcgen05 / TMEM / TMA attention (csrc/attention_tc.cu) against the mma.sync kernels (csrc/attention.cu) on identical inputs:
same masks, same dropout counters, same log-sum-exp layout.  CRCT_ATTN_LEGACY_NOW=1 (read per call) selects the mma.sync path."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from cqa_crct_b200 import _lib as L   # noqa: E402

DEV = 'cuda'


def bf(x):
    return x.to(torch.bfloat16)


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))


os.environ['CRCT_ATTN_TC_POLICY'] = '15'          # this module: tcgen05 wherever eligible (the default policy is per-shape, by measurement)


class legacy:
    def __enter__(self):
        os.environ['CRCT_ATTN_LEGACY_NOW'] = '1'

    def __exit__(self, *a):
        os.environ.pop('CRCT_ATTN_LEGACY_NOW', None)


SHAPES = [(3, 16, 48, 124, 124), (2, 16, 64, 44, 44), (3, 32, 32, 124, 44), (3, 32, 32, 44, 124), (2, 4, 48, 17, 33), (1, 2, 64, 128, 128),
          (2, 2, 32, 64, 64), (2, 3, 48, 65, 3), (80, 16, 48, 124, 124)]


@pytest.mark.parametrize('B,nh,dh,Lq,Lk', SHAPES)
@pytest.mark.parametrize('p', [0.0, 0.1])
@pytest.mark.parametrize('packed', [False, True])
def test_tc_attention_matches_mma_sync_attention(B, nh, dh, Lq, Lk, p, packed):
    torch.manual_seed(B + dh + Lq)
    H = nh * dh
    ld = 3 * H                                                    # q / k / v as column slices of packed projections, as the engine calls it
    g = torch.Generator().manual_seed(5)
    qkv_q = bf(torch.randn(B * Lq, ld, device=DEV))
    qkv_k = qkv_q if Lq == Lk else bf(torch.randn(B * Lk, ld, device=DEV))
    dout = bf(torch.randn(B * Lq, H, device=DEV))
    lq = torch.randint(1, Lq + 1, (B,), generator=g)
    lk = lq if Lq == Lk else torch.randint(1, Lk + 1, (B,), generator=g)
    mask = torch.zeros(B, Lk, device=DEV)
    cu_q = cu_k = None
    if packed:
        cu_q = torch.zeros(B + 1, dtype=torch.int32); cu_q[1:] = lq.cumsum(0); cu_q = cu_q.to(DEV)
        cu_k = torch.zeros(B + 1, dtype=torch.int32); cu_k[1:] = lk.cumsum(0); cu_k = cu_k.to(DEV)
        qkv_q[int(cu_q[-1]):] = float('nan')                      # rows past the packed total: never-written memory
        if qkv_k is not qkv_q:
            qkv_k[int(cu_k[-1]):] = float('nan')
        dout[int(cu_q[-1]):] = float('nan')
        mask = None
    else:
        for b in range(B):
            mask[b, int(lk[b]):] = -10000.0
    res = {}
    for impl in ('tc', 'mma'):
        out = torch.zeros(B * Lq, H, device=DEV, dtype=torch.bfloat16)
        lse = torch.zeros(B, nh, Lq, device=DEV)
        dq, dk = torch.zeros(B * Lq, ld, device=DEV, dtype=torch.bfloat16), torch.zeros(B * Lk, ld, device=DEV, dtype=torch.bfloat16)
        kw = dict(B=B, nh=nh, dh=dh, Lq=Lq, Lk=Lk, ldq=ld, ldk=ld, ldv=ld, ldo=H, dropout_p=p, seed=77, cu_q=cu_q, cu_k=cu_k)

        def run():
            L.attn_fwd(qkv_q, qkv_k[:, H:], qkv_k[:, 2 * H:], mask, out, lse, **kw)
            L.attn_bwd(qkv_q, qkv_k[:, H:], qkv_k[:, 2 * H:], mask, out, dout, lse, dq, dk[:, H:], dk[:, 2 * H:], lddo=H, lddq=ld, lddk=ld, lddv=ld, **kw)
        if impl == 'mma':
            with legacy():
                run()
        else:
            run()
        torch.cuda.synchronize()
        res[impl] = (out, lse, dq[:, :H], dk[:, H:2 * H], dk[:, 2 * H:])
    names = ['out', 'lse', 'dq', 'dk', 'dv']
    for i, name in enumerate(names):
        a, c = res['tc'][i].float(), res['mma'][i].float()
        if packed:                                                # compare the valid rows
            rows_q = torch.cat([torch.arange(int(cu_q[b]), int(cu_q[b + 1])) for b in range(B)]).to(DEV)
            rows_k = torch.cat([torch.arange(int(cu_k[b]), int(cu_k[b + 1])) for b in range(B)]).to(DEV)
            if name == 'lse':
                sel = torch.zeros(B, nh, Lq, dtype=torch.bool, device=DEV)
                for b in range(B):
                    sel[b, :, :int(lq[b])] = True
                a, c = a[sel], c[sel]
            else:
                rows = rows_q if name in ('out', 'dq') else rows_k
                a, c = a[rows], c[rows]
        elif name != 'lse':                                       # padded: rows of padded queries / keys are defined too
            pass
        assert bool(torch.isfinite(a).all()), name
        tol = 1e-5 if name == 'lse' else (8e-3 if name == 'out' else 1.6e-2)
        assert relmax(a, c) < tol, (name, relmax(a, c))


def test_tc_attention_is_the_default_for_single_tile_shapes():
    """The dispatch itself: a 124 x 124 problem gives different bits with CRCT_ATTN_LEGACY_NOW (two implementations), a 248 x 248
    problem the same bits (both run mma.sync)."""
    torch.manual_seed(0)
    for Lseq, same in ((124, False), (248, True)):
        B, nh, dh = 2, 4, 48
        H = nh * dh
        q, k, v = (bf(torch.randn(B * Lseq, H, device=DEV)) for _ in range(3))
        mask = torch.zeros(B, Lseq, device=DEV)
        outs = []
        for use_legacy in (False, True):
            out = torch.zeros(B * Lseq, H, device=DEV, dtype=torch.bfloat16)
            if use_legacy:
                with legacy():
                    L.attn_fwd(q, k, v, mask, out, None, B=B, nh=nh, dh=dh, Lq=Lseq, Lk=Lseq, ldq=H, ldk=H, ldv=H, ldo=H)
            else:
                L.attn_fwd(q, k, v, mask, out, None, B=B, nh=nh, dh=dh, Lq=Lseq, Lk=Lseq, ldq=H, ldk=H, ldv=H, ldo=H)
            outs.append(out)
        assert torch.equal(outs[0], outs[1]) == same
        assert relmax(outs[0].float(), outs[1].float()) < 8e-3
