"""A few isolated launches of the three big tail kernels at the bench shape (B=64, 4 heads, 224x224), L2 flushed between them:
the target of `ncu --set full -k regex:heatmap_fwd_kernel|heatmap_bwd_stage1|head_project_fwd` (tools/gpu_tail_ncu.sh)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200 import _cabi as K
lib = K.load()
B, heads, g = int(os.environ.get("TAIL_B", 64)), 4, 14
H = 16 * g
dev = "cuda"
st = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 * 1024 * 1024, device=dev, dtype=torch.uint8)
grid = torch.randn(B, heads, g * g, device=dev)
ts_h = torch.empty(H, dtype=torch.int32); tw_h = torch.empty(H, 3)
K.check(lib.bfk_upsample16_taps(g, ts_h.data_ptr(), tw_h.data_ptr()))
ts, tw = ts_h.to(dev), tw_h.to(dev)
heat = torch.empty(heads, B, H, H, device=dev)
targets = [torch.rand(B, H, H, device=dev) for _ in range(heads)]
tptr = torch.tensor([t.data_ptr() for t in targets], dtype=torch.int64).to(dev)
loss = torch.empty(heads + 1, device=dev); key = torch.empty(heads * B, device=dev, dtype=torch.int64)
part = torch.empty(lib.bfk_heatmap_fwd_workspace_floats(B, heads, g), device=dev)
counters = torch.zeros(1 + heads * B, device=dev, dtype=torch.int32)
fa = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), tptr.data_ptr(), None, 0,
                      loss.data_ptr(), key.data_ptr(), part.data_ptr(), counters.data_ptr())
gscale = torch.ones(heads + 1, device=dev); dgrid = torch.empty(B, heads, g * g, device=dev)
ws = torch.empty(lib.bfk_heatmap_bwd_workspace_floats(B, heads, g), device=dev)
ba = K.HeatmapBwdArgs(B, heads, g, 0, heat.data_ptr(), tptr.data_ptr(), None, 0, gscale.data_ptr(), 1.0 / (B * H * H), 1, None,
                      ts.data_ptr(), tw.data_ptr(), dgrid.data_ptr(), ws.data_ptr(), counters.data_ptr())
D, Np, Nf = 768, g * g, 262
X = torch.randn(B * Nf, D, device=dev); w_eff = torch.randn(heads, D, device=dev); b_eff = torch.randn(heads, device=dev)
grid2 = torch.empty_like(grid)
pa = K.HeadProjectArgs(B, heads, Np, D, X.data_ptr(), Nf * D, Nf - Np, w_eff.data_ptr(), b_eff.data_ptr(), grid2.data_ptr(),
                       None, Nf, None, None, 0, None, None, None)
for it in range(3):
    flush.sum(); K.check(lib.bfk_head_project_fwd(C.byref(pa), st))
    flush.sum(); K.check(lib.bfk_heatmap_fwd(C.byref(fa), st))
    flush.sum(); K.check(lib.bfk_heatmap_bwd(C.byref(ba), st))
torch.cuda.synchronize()
print("ok", loss.tolist())
