"""Cost of the fused epilogues: one FFN-shaped GEMM per variant, cold L2, CUDA-event timed (us, TFLOP/s)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L

dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(name, M, N, K, reps=12, **kw):
    b_major = kw.get('b_major', 0)
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(K, N, device=dev) * 0.5).bfloat16() if b_major else (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    D = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ts = []
    for r in range(reps + 2):
        flush.fill_(r)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.gemm(A, B, D, M=M, N=N, K=K, **kw)
        e1.record()
        torch.cuda.synchronize()
        if r >= 2:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    us = ts[len(ts) // 2]
    print(f'{name:34s} M={M} N={N} K={K}  {us:7.1f} us  {2.0 * M * N * K / us / 1e6:7.0f} TFLOP/s', flush=True)


for (M, I, H) in ((9920, 3072, 768), (3520, 1024, 1024)):
    bias = torch.zeros(I, device=dev)
    biasH = torch.zeros(H, device=dev)
    auxI = torch.randn(M, I, device=dev).bfloat16()
    auxH = torch.randn(M, H, device=dev).bfloat16()
    D2 = torch.empty(M, I, device=dev, dtype=torch.bfloat16)
    for cg in (1, 2):
        t = f'cg{cg} '
        run(t + 'up   bias', M, I, H, bias=bias, cta_group=cg)
        run(t + 'up   gelu (no D2)', M, I, H, bias=bias, epilogue=L.EPI_BIAS_GELU, cta_group=cg)
        run(t + 'up   gelu + gelu\'', M, I, H, bias=bias, epilogue=L.EPI_BIAS_GELU, D2=D2, cta_group=cg)
        run(t + 'dgrad plain', M, I, H, b_major=1, cta_group=cg)
        run(t + 'dgrad mul', M, I, H, b_major=1, epilogue=L.EPI_MUL, aux=auxI, cta_group=cg)
        run(t + 'down bias', M, H, I, bias=biasH, cta_group=cg)
        run(t + 'down res', M, H, I, bias=biasH, epilogue=L.EPI_BIAS_RES, aux=auxH, cta_group=cg)
        run(t + 'down res+dropout', M, H, I, bias=biasH, epilogue=L.EPI_BIAS_RES, aux=auxH, dropout_p=0.1, seed=5, cta_group=cg)
