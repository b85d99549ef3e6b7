"""GPU bring-up for the tcgen05 GEMM: runs each variant in a child process under a timeout so a
deadlocked kernel cannot take the whole call down.  Usage: python tools/bringup_gemm.py [stage]"""
import itertools
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_and_run(M, N, K, a_major, b_major, epi, block_n=0, dbg=None, split_k=0, max_ctas=0, seed=0):
    import torch
    from cqa_crct_b200 import _lib as L
    torch.manual_seed(seed)
    dev = 'cuda'
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    ref = A.float() @ B.float().t()
    A_st = A.t().contiguous() if a_major else A          # a_major=1: stored [K, M]
    B_st = B.t().contiguous() if b_major else B          # b_major=1: stored [K, N]
    bias = torch.randn(N, device=dev)
    aux = torch.randn(M, N, device=dev).bfloat16()
    kw = dict(M=M, N=N, K=K, a_major=a_major, b_major=b_major, epilogue=epi, block_n=block_n, dbg=dbg,
              split_k=split_k, max_ctas=max_ctas)
    if epi == L.EPI_F32:
        D = torch.zeros(M, N, device=dev)
        L.gemm(A_st, B_st, D, accumulate=1, **kw)
        out, want = D, ref
    elif epi == L.EPI_BIAS:
        D = torch.full((M, N), 7.0, device=dev).bfloat16()
        L.gemm(A_st, B_st, D, bias=bias, **kw)
        out, want = D.float(), ref + bias
    elif epi == L.EPI_BIAS_GELU:
        D = torch.zeros(M, N, device=dev).bfloat16(); D2 = torch.zeros_like(D)
        L.gemm(A_st, B_st, D, bias=bias, D2=D2, **kw)
        u = ref + bias
        x = u.clone().requires_grad_(True)
        torch.nn.functional.gelu(x).sum().backward()
        out, want = torch.cat([D.float(), D2.float()]), torch.cat([torch.nn.functional.gelu(u), x.grad])
    elif epi == L.EPI_BIAS_RES:
        D = torch.zeros(M, N, device=dev).bfloat16()
        L.gemm(A_st, B_st, D, bias=bias, aux=aux, **kw)
        out, want = D.float(), ref + bias + aux.float()
    elif epi == L.EPI_MUL:
        D = torch.zeros(M, N, device=dev).bfloat16()
        L.gemm(A_st, B_st, D, aux=aux, **kw)
        out, want = D.float(), ref * aux.float()
    torch.cuda.synchronize()
    err = (out - want).abs().max().item()
    scale = want.abs().max().item()
    return err / scale


def child(spec):
    import torch
    from cqa_crct_b200 import _lib as L
    L.check(L.lib().crct_device_check())
    M, N, K, am, bm, epi, bn, split_k, max_ctas = spec[:9]
    dbg = spec[9:] if len(spec) > 9 else None
    rel = ref_and_run(M, N, K, am, bm, epi, bn, dbg if dbg else None, split_k, max_ctas)
    print(f'RESULT rel_err={rel:.3e}')


def run_child(spec, timeout=90):
    cmd = [sys.executable, os.path.abspath(__file__), 'child'] + [str(x) for x in spec]
    t0 = time.time()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
        out = (r.stdout + r.stderr).strip().splitlines()
        res = [l for l in out if l.startswith('RESULT')]
        msg = res[0] if res else 'FAIL rc=%d %s' % (r.returncode, ' | '.join(out[-3:]))
    except subprocess.TimeoutExpired:
        msg = 'TIMEOUT'
    print(f'{spec} -> {msg}  ({time.time() - t0:.1f}s)', flush=True)
    return msg


def main():
    E_BIAS, E_GELU, E_RES, E_MUL, E_F32 = 0, 1, 2, 3, 4
    ok = lambda m: m.startswith('RESULT') and float(m.split('=')[1]) < 2e-2
    # 1. K-major / K-major, smallest then persistent / multi-tile / ragged
    base = [(128, 128, 64, 0, 0, E_BIAS, 128, 0, 0), (128, 256, 64, 0, 0, E_BIAS, 256, 0, 0),
            (128, 128, 256, 0, 0, E_BIAS, 128, 0, 0), (256, 512, 768, 0, 0, E_BIAS, 256, 0, 0),
            (1000, 192, 192, 0, 0, E_BIAS, 0, 0, 0), (9920, 2304, 768, 0, 0, E_BIAS, 0, 0, 0),
            (1024, 1024, 512, 0, 0, E_BIAS, 128, 0, 3), (1024, 1024, 512, 0, 0, E_BIAS, 256, 0, 3),
            (3520, 1024, 1024, 0, 0, E_BIAS_GELU if False else E_GELU, 0, 0, 0), (992, 768, 3072, 0, 0, E_RES, 0, 0, 0)]
    kk_ok = all([ok(run_child(s)) for s in base])
    print('K-major/K-major:', 'PASS' if kk_ok else 'FAIL', flush=True)
    # 2. dgrad form (B MN-major) and wgrad form (both MN-major)
    dg = [(128, 128, 64, 0, 1, E_BIAS, 128, 0, 0), (128, 256, 128, 0, 1, E_BIAS, 256, 0, 0),
          (1000, 768, 3072, 0, 1, E_MUL, 0, 0, 0), (9920, 768, 2304, 0, 1, E_RES, 0, 0, 0)]
    dg_ok = all([ok(run_child(s)) for s in dg])
    print('dgrad (B MN-major):', 'PASS' if dg_ok else 'FAIL', flush=True)
    wg = [(128, 128, 64, 1, 1, E_F32, 128, 1, 0), (256, 256, 512, 1, 1, E_F32, 256, 1, 0),
          (768, 768, 9920, 1, 1, E_F32, 0, 0, 0), (3072, 768, 992, 1, 1, E_F32, 0, 0, 0), (192, 576, 1000, 1, 1, E_F32, 0, 0, 0)]
    wg_ok = all([ok(run_child(s)) for s in wg])
    print('wgrad (A,B MN-major):', 'PASS' if wg_ok else 'FAIL', flush=True)
    # 3. if an MN-major form failed, sweep descriptor fields on the smallest case
    if not dg_ok:
        print('--- sweeping B MN-major descriptor (lbo, sbo, kstep) ---', flush=True)
        for lbo, sbo, kstep in itertools.product((8192, 1024, 2048, 16384), (1024, 8192, 2048), (2048, 32, 4096)):
            run_child((128, 128, 64, 0, 1, E_BIAS, 128, 0, 0, 1, 16, 1024, 32, lbo, sbo, kstep), timeout=60)
    if dg_ok and not wg_ok:
        print('--- sweeping A MN-major descriptor (lbo, sbo, kstep) ---', flush=True)
        for lbo, sbo, kstep in itertools.product((8192, 1024, 2048, 16384), (1024, 8192, 2048), (2048, 32, 4096)):
            run_child((128, 128, 64, 1, 1, E_F32, 128, 1, 0, 1, lbo, sbo, kstep, 8192, 1024, 2048), timeout=60)
    if not kk_ok:
        print('--- sweeping K-major descriptor (lbo, sbo, kstep) ---', flush=True)
        for lbo, sbo, kstep in itertools.product((16, 0, 1024), (1024, 128, 2048), (32, 64)):
            run_child((128, 128, 64, 0, 0, E_BIAS, 128, 0, 0, 1, lbo, sbo, kstep, lbo, sbo, kstep), timeout=60)


def perf():
    import torch
    from cqa_crct_b200 import _lib as L
    dev = 'cuda'
    shapes = [(9920, 2304, 768), (9920, 768, 768), (9920, 3072, 768), (9920, 768, 3072), (3520, 1024, 1024), (3520, 3072, 1024), (9920, 3072, 768)]
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
        D = torch.empty(M, N, device=dev).bfloat16(); bias = torch.zeros(N, device=dev)
        for bn, cg in ((128, 1), (256, 1), (128, 2), (256, 2)):
            for _ in range(3):
                L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, block_n=bn, cta_group=cg)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, block_n=bn, cta_group=cg)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 20
            print(f'perf M={M} N={N} K={K} BN={bn} cta_group={cg}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s', flush=True)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3): torch.matmul(A, B.t())
        s.record()
        for _ in range(20): torch.matmul(A, B.t())
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 20
        print(f'  cublas ref: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'child':
        child([int(x) for x in sys.argv[2:]])
    elif len(sys.argv) > 1 and sys.argv[1] == 'perf':
        perf()
    elif len(sys.argv) > 1 and sys.argv[1] == 'perf_wgrad':
        pass
    else:
        main()


def perf_wgrad():
    import torch
    from cqa_crct_b200 import _lib as L
    dev = 'cuda'
    for (rows, No, Ki) in [(9920, 3072, 768), (9920, 768, 3072), (9920, 768, 768), (9920, 2304, 768), (3520, 1024, 1024), (3520, 3072, 1024), (9920, 1024, 768)]:
        dy = torch.randn(rows, No, device=dev).bfloat16(); x = torch.randn(rows, Ki, device=dev).bfloat16()
        dW = torch.zeros(No, Ki, device=dev)
        for bn, cg, sk in ((128, 1, 0), (256, 1, 0), (128, 2, 0), (256, 2, 0), (256, 2, 1), (256, 2, 2), (256, 2, 4)):
            kw = dict(M=No, N=Ki, K=rows, a_major=1, b_major=1, epilogue=L.EPI_F32, accumulate=1, block_n=bn, cta_group=cg, split_k=sk)
            for _ in range(3):
                L.gemm(dy, x, dW, **kw)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(20):
                L.gemm(dy, x, dW, **kw)
            e.record(); torch.cuda.synchronize()
            ms = s.elapsed_time(e) / 20
            print(f'wgrad rows={rows} out={No} in={Ki} BN={bn} cta_group={cg} split_k={sk}: {ms*1e3:.1f} us  {2*rows*No*Ki/ms/1e9:.1f} TFLOP/s', flush=True)


if __name__ == '__main__' and len(sys.argv) > 1 and sys.argv[1] == 'perf_wgrad':
    perf_wgrad()
