"""Attention micro-benchmark on the policy's three shapes (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
def timeit(name, fn, reps=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:36s} {1e3*ts[len(ts)//2]:8.1f} us", flush=True)
shapes = [("vision", 64, 196, 12, 64), ("text", 64, 64, 12, 64), ("fusion", 64, 262, 16, 48)]
only = os.environ.get("ATTN_SHAPE")
for name, B, N, H, hd in shapes:
    if only and only != name: continue
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H * N, device="cuda")
    dout = torch.randn(B * N, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    mask = torch.ones(B, N, device="cuda") if name == "fusion" else None
    sc = hd ** -0.5
    timeit(f"{name} fwd  B{B} N{N} H{H} hd{hd}", lambda: k.attention_fwd(qkv, out, lse, B, N, H, hd, sc, row_mask=mask))
    timeit(f"{name} bwd", lambda: k.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, sc, row_mask=mask))
