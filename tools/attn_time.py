"""Timing only (vision / fusion / text shapes), used with BFK_ATTN_DBG ablations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
def timeit(fn, reps=6):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return 1e3 * ts[len(ts) // 2]
res = []
for name, B, N, H, hd in [("vision", 64, 196, 12, 64), ("fusion", 64, 262, 16, 48), ("text", 64, 64, 12, 64)]:
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.zeros(B * H * N, device="cuda")
    dout = torch.randn(B * N, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    sc = hd ** -0.5
    f = timeit(lambda: k.attention_fwd(qkv, out, lse, B, N, H, hd, sc))
    b = timeit(lambda: k.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, sc)) if os.environ.get("ATTN_BWD", "1") == "1" else 0.0
    res.append(f"{name} fwd {f:6.1f} bwd {b:6.1f}")
print(f"DBG={os.environ.get('BFK_ATTN_DBG', '0'):>3s} | " + " | ".join(res), flush=True)
