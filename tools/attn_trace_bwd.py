import os, sys, re
os.environ["BFK_ATTN_DBG"] = "64"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
B, N, H, hd = 64, 196, 12, 64
D = H * hd
qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).bfloat16()
out = torch.randn(B * N, D, device="cuda").bfloat16()
lse = torch.zeros(B * H * N, device="cuda")
dout = torch.randn(B * N, D, device="cuda").bfloat16()
dqkv = torch.empty_like(qkv)
for _ in range(3):
    lse.zero_()
    k.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, hd ** -0.5)
torch.cuda.synchronize()
raw = lse.view(torch.int32).cpu().numpy().astype("uint32")
ev = raw[0:4096].reshape(-1, 2)
ev = [(int(t), int(c)) for t, c in ev if 0 < t < 1000]
prev = ev[0][1]; out_s = []
for t, c in ev[:90]:
    out_s.append(f"{t}(+{(c - prev) & 0xffffffff})"); prev = c
print("ROW", " ".join(out_s))
