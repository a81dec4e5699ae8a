"""Column-sum (bias gradient) kernel on the fusion shapes, L2 flushed (run on the GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for T, M, N in ((torch.float32, 16768, 768), (torch.bfloat16, 16768, 3072)):
    X = torch.randn(M, N, device="cuda").to(T); out = torch.empty(N, device="cuda")
    ts = []
    for it in range(8):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); k.colsum(X, M, N, N, out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    ok = torch.allclose(out, X.float().sum(0), atol=5e-2, rtol=1e-3)
    print(f"colsum {T} {M}x{N}: {ms*1e3:.1f} us  {X.numel()*X.element_size()/ms/1e6:.0f} GB/s  ok={ok}", flush=True)
