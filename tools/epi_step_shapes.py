"""The epilogue-heavy GEMMs of the train step at their in-step shapes (packed rows, dropout, fp32 residual), two launches each,
for `ncu --set full --import-source on -k regex:gemm_tcgen05` captures: RES_F32 K=768, RES_F32 K=3072, GELU+GELU', bias only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cqa_crct_b200 import _lib as L
dev = 'cuda'
M, rows, H, I = 9920, 6900, 768, 3072
n = torch.tensor([rows], dtype=torch.int32, device=dev); n.hint = rows
src = torch.arange(M, dtype=torch.int32, device=dev)
x = (torch.randn(M, H, device=dev) * 0.5).bfloat16(); h = (torch.randn(M, I, device=dev) * 0.5).bfloat16()
Wo = (torch.randn(H, H, device=dev) * 0.05).bfloat16(); W1 = (torch.randn(I, H, device=dev) * 0.05).bfloat16(); W2 = (torch.randn(H, I, device=dev) * 0.05).bfloat16()
bH, bI = torch.zeros(H, device=dev), torch.zeros(I, device=dev)
aux32 = torch.randn(M, H, device=dev); z = torch.empty(M, H, device=dev)
u = torch.empty(M, I, device=dev, dtype=torch.bfloat16); dg = torch.empty_like(u)
which = sys.argv[1:] or ['res768', 'res3072', 'gelu', 'bias']
for _ in range(2):
    if 'res768' in which:
        L.gemm(x, Wo, z, M=M, N=H, K=H, bias=bH, epilogue=L.EPI_BIAS_RES_F32, aux=aux32, dropout_p=0.1, seed=5, rows_dev=n, drop_rows=src)
    if 'res3072' in which:
        L.gemm(h, W2, z, M=M, N=H, K=I, bias=bH, epilogue=L.EPI_BIAS_RES_F32, aux=aux32, dropout_p=0.1, seed=5, rows_dev=n, drop_rows=src)
    if 'gelu' in which:
        L.gemm(x, W1, u, M=M, N=I, K=H, bias=bI, epilogue=L.EPI_BIAS_GELU, D2=dg, rows_dev=n)
    if 'bias' in which:
        L.gemm(x, W1, u, M=M, N=I, K=H, bias=bI, rows_dev=n)
torch.cuda.synchronize()
