"""Runs a handful of GEMM launches (for `ncu --set full` captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L
M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (9920, 2304, 768)
A = torch.randn(M, K, device='cuda').bfloat16(); B = torch.randn(N, K, device='cuda').bfloat16()
D = torch.empty(M, N, device='cuda', dtype=torch.bfloat16); bias = torch.zeros(N, device='cuda')
for cg in (1, 2):
    for _ in range(3):
        L.gemm(A, B, D, M=M, N=N, K=K, bias=bias, block_n=256, cta_group=cg)
torch.cuda.synchronize()
