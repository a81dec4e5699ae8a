"""Epilogue diagnosis (single-CTA kernel, BFK_GEMM_2CTA=0): BFK_GEMM_DBG bits 2 = no main loop, 4 = no stores, 8 = no TMEM read."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
def timeit(name, fn, reps=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:50s} {1e3*ts[len(ts)//2]:8.1f} us", flush=True)
dbg = os.environ.get("BFK_GEMM_DBG", "0")
M, N, Kd = 12544, 3072, 768
A = torch.randn(M, Kd, device="cuda").bfloat16(); B = torch.randn(N, Kd, device="cuda").bfloat16()
D16 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
D32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
timeit(f"dbg={dbg} {M}x{N}x{Kd} bf16 out", lambda: k.gemm(A, B, D16, M, N, Kd, Kd, Kd, N))
timeit(f"dbg={dbg} {M}x{N}x{Kd} fp32 out", lambda: k.gemm(A, B, D32, M, N, Kd, Kd, Kd, N))
