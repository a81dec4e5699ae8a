"""LayerNorm forward variants on the vision-tower shape (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
k.seed_dev = torch.zeros(1, device="cuda", dtype=torch.int64)
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
    print(f"{name:44s} {1e3*ts[len(ts)//2]:8.1f} us", flush=True)
M, D = 12544, 768
x = torch.randn(M, D, device="cuda"); g = torch.randn(D, device="cuda"); b = torch.randn(D, device="cuda")
y = torch.empty(M, D + 16, device="cuda", dtype=torch.bfloat16); y2 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
A = torch.randn(16, D, device="cuda").bfloat16()
mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
timeit("ln fwd plain", lambda: k.layernorm_fwd(x, M, D, g, b, 1e-6, y, D + 16, mean, rstd))
timeit("ln fwd + dropped copy", lambda: k.layernorm_fwd(x, M, D, g, b, 1e-6, y, D + 16, mean, rstd, y2=y2, ldy2=D, y2_p=0.1, y2_seed=3))
timeit("ln fwd + dropped copy + lora down", lambda: k.layernorm_fwd(x, M, D, g, b, 1e-6, y, D + 16, mean, rstd, y2=y2, ldy2=D, y2_p=0.1, y2_seed=3, lora_a=A, lora_u=y[:, D:], ld_u=D + 16))
timeit("lora down as N=16 GEMM", lambda: k.gemm(y2, A, y[:, D:], M, 16, D, D, D, D + 16))
