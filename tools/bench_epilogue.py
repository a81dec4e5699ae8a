"""Micro-benchmark of the tcgen05 GEMM epilogue variants on one shape (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
from bifold_b200 import _cabi as K
k = Kernels()
M, N, Kd = 12544, 3072, 768
A = torch.randn(M, Kd, device="cuda").bfloat16()
B = torch.randn(N, Kd, device="cuda").bfloat16()       # trans_b = 0
Bt = torch.randn(Kd, N, device="cuda").bfloat16()      # trans_b = 1
bias = torch.randn(N, device="cuda")
flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
def timeit(name, fn, reps=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:40s} {1e3*ts[len(ts)//2]:8.1f} us   {2*M*N*Kd/ts[len(ts)//2]/1e9:7.1f} TF/s", flush=True)
D16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
pa = torch.randn(M, N, device="cuda").bfloat16()
timeit("plain bf16 out", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N))
timeit("bias", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, bias=bias))
timeit("bias+gelu_tanh+aux", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_TANH, aux=aux, ld_aux=N))
timeit("bias+gelu_erf+aux", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_ERF, aux=aux, ld_aux=N))
timeit("bias+gelu_tanh (no aux)", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_TANH))
timeit("trans_b plain", lambda: k.gemm(A, Bt, D16, M, N, Kd, Kd, N, N, trans_b=True))
timeit("trans_b dgelu_tanh", lambda: k.gemm(A, Bt, D16, M, N, Kd, Kd, N, N, trans_b=True, act=K.ACT_DGELU_TANH, aux=pa, ld_aux=N))
timeit("trans_b dgelu_erf", lambda: k.gemm(A, Bt, D16, M, N, Kd, Kd, N, N, trans_b=True, act=K.ACT_DGELU_ERF, aux=pa, ld_aux=N))
aux2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
timeit("bias+gelu_tanh_grad (value+derivative)", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_TANH_GRAD, aux=aux2, ld_aux=N))
timeit("bias+gelu_erf_grad (value+derivative)", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, bias=bias, act=K.ACT_GELU_ERF_GRAD, aux=aux2, ld_aux=N))
timeit("trans_b mul_aux", lambda: k.gemm(A, Bt, D16, M, N, Kd, Kd, N, N, trans_b=True, act=K.ACT_MUL_AUX, aux=pa, ld_aux=N))
timeit("plain dgelu_tanh (trans_b=0)", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N, act=K.ACT_DGELU_TANH, aux=pa, ld_aux=N))
# residual fp32
M2, N2 = 12544, 768
B2 = torch.randn(N2, Kd, device="cuda").bfloat16()
res = torch.randn(M2, N2, device="cuda"); D32 = torch.empty(M2, N2, device="cuda")
D16b = torch.empty(M2, N2, device="cuda", dtype=torch.bfloat16)
def t2(name, fn):
    global M, N
    M0, N0 = M, N
    M, N = M2, N2
    timeit(name, fn)
    M, N = M0, N0
t2("N=768 plain bf16", lambda: k.gemm(A, B2, D16b, M2, N2, Kd, Kd, Kd, N2))
t2("N=768 bias+residual fp32", lambda: k.gemm(A, B2, D32, M2, N2, Kd, Kd, Kd, N2, bias=bias[:N2].contiguous(), residual=res, ldr=N2))
