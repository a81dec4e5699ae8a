"""GEMM diagnosis on the policy's main shapes: full kernel vs main-loop-only vs epilogue-only (BFK_GEMM_DBG) vs cuBLAS."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
from bifold_b200 import _cabi as K
k = Kernels()
flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
def timeit(name, fn, flops, reps=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:44s} {1e3*ts[len(ts)//2]:8.1f} us   {flops/ts[len(ts)//2]/1e9:7.1f} TF/s", flush=True)
dbg = os.environ.get("BFK_GEMM_DBG", "0")
for (M, N, Kd) in [(12544, 3072, 768), (12544, 768, 3072), (12544, 2304, 784), (12544, 768, 768), (16768, 3072, 768)]:
    A = torch.randn(M, Kd, device="cuda").bfloat16()
    B = torch.randn(N, Kd, device="cuda").bfloat16()
    D16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    fl = 2 * M * N * Kd
    timeit(f"bfk dbg={dbg} {M}x{N}x{Kd} plain bf16", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N), fl)
    if dbg == "0":
        Bt = B.t()
        timeit(f"cuBLAS {M}x{N}x{Kd} (torch.matmul)", lambda: torch.matmul(A, Bt, out=D16), fl)

# QKV dgrad with the LoRA columns: M x 784 = dqkv[M x 2304] . Wext[2304 x 784] (B stored [K, N]); wgrad split-K
M, N, Kd = 12544, 784, 2304
A = torch.randn(M, Kd, device="cuda").bfloat16(); Bt = torch.randn(Kd, N, device="cuda").bfloat16()
D16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
timeit(f"bfk dgrad N=784 {M}x{N}x{Kd} trans_b", lambda: k.gemm(A, Bt, D16, M, N, Kd, Kd, N, N, trans_b=True), 2 * M * N * Kd)
M, N, Kd = 3072, 768, 16768
At = torch.randn(Kd, M, device="cuda").bfloat16(); Bt = torch.randn(Kd, N, device="cuda").bfloat16()
D32 = torch.empty(M, N, device="cuda")
timeit(f"bfk wgrad {M}x{N}x{Kd} split-K", lambda: k.gemm(At, Bt, D32, M, N, Kd, M, N, N, trans_a=True, trans_b=True), 2 * M * N * Kd)
