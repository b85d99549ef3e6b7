"""A few FFN-up launches per epilogue variant (for `ncu --set full` captures): bias, gelu+gelu', dgrad mul; 2-CTA form."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L
M, I, H = 9920, 3072, 768
dev = 'cuda'
A = (torch.randn(M, H, device=dev) * 0.5).bfloat16(); W = (torch.randn(I, H, device=dev) * 0.5).bfloat16()
D = torch.empty(M, I, device=dev, dtype=torch.bfloat16); D2 = torch.empty_like(D); aux = torch.randn(M, I, device=dev).bfloat16()
bias = torch.zeros(I, device=dev)
for _ in range(2):
    L.gemm(A, W, D, M=M, N=I, K=H, bias=bias, cta_group=2)
    L.gemm(A, W, D, M=M, N=I, K=H, bias=bias, epilogue=L.EPI_BIAS_GELU, D2=D2, cta_group=2)
    L.gemm(A, W.t().contiguous().view(H, I), D, M=M, N=I, K=H, b_major=1, epilogue=L.EPI_MUL, aux=aux, cta_group=2)
torch.cuda.synchronize()
