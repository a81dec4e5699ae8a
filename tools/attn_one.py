"""One forward + one backward attention launch per shape (profiling target for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
for name, B, N, H, hd in [("vision", 64, 196, 12, 64), ("fusion", 64, 262, 16, 48)]:
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H * N, device="cuda")
    dout = torch.randn(B * N, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    sc = hd ** -0.5
    for _ in range(2):
        k.attention_fwd(qkv, out, lse, B, N, H, hd, sc)
        k.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, sc)
    torch.cuda.synchronize()
print("done")
