"""Dump the in-kernel pipeline trace of CTA 0 (BFK_ATTN_DBG=64) for the vision forward shape."""
import os, sys
os.environ["BFK_ATTN_DBG"] = os.environ.get("BFK_ATTN_DBG", "64")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
B, N, H, hd = 64, 196, 12, 64
D = H * hd
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).bfloat16()
out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
lse = torch.zeros(B * H * N, device="cuda")
for _ in range(3):
    lse.zero_()
    k.attention_fwd(qkv, out, lse, B, N, H, hd, hd ** -0.5)
torch.cuda.synchronize()
raw = lse.view(torch.int32).cpu().numpy().astype("uint32")
def dump(name, off, cap):
    ev = raw[off:off + cap].reshape(-1, 2)
    ev = [(int(t), int(c)) for t, c in ev if t != 0]
    if not ev: return
    t0 = ev[0][1]
    print(name, " ".join(f"{t}:{(c - t0) & 0xffffffff}" for t, c in ev[:160]))
dump("ROW", 0, 4096); dump("MMA", 4096, 4096); dump("LDR", 8192, 2048)
