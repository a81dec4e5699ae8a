"""One tcgen05 GEMM variant a few times (ncu target).  usage: gemm_one.py <variant: plain|gg|ge_aux|res|mulaux>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
from bifold_b200 import _cabi as K
k = Kernels()
var = sys.argv[1] if len(sys.argv) > 1 else "gg"
M, N, Kd = 12544, 3072, 768
A = torch.randn(M, Kd, device="cuda").bfloat16(); B = torch.randn(N, Kd, device="cuda").bfloat16(); Bt = torch.randn(Kd, N, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda"); D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); aux = torch.randn(M, N, device="cuda").bfloat16()
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for _ in range(4):
    flush.zero_()
    if var == "plain": k.gemm(A, B, D, M, N, Kd, Kd, Kd, N, bias=bias)
    elif var == "gg": k.gemm(A, B, D, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_TANH_GRAD, aux=aux, ld_aux=N)
    elif var == "ge_aux": k.gemm(A, B, D, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_TANH, aux=aux, ld_aux=N)
    elif var == "mulaux": k.gemm(A, Bt, D, M, N, Kd, Kd, N, N, trans_b=True, act=K.ACT_MUL_AUX, aux=aux, ld_aux=N)
    elif var == "res":
        if "R" not in globals():
            B2 = torch.randn(768, Kd, device="cuda").bfloat16(); R = torch.randn(M, 768, device="cuda"); D32 = torch.empty(M, 768, device="cuda")
        k.gemm(A, B2, D32, M, 768, Kd, Kd, Kd, 768, bias=bias[:768].contiguous(), residual=R, ldr=768)
torch.cuda.synchronize()
print("done", var)
