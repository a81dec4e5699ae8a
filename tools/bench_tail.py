"""HBM roofline of the heatmap tail kernels (forward: upsample+sigmoid+BCE+argmax; backward; argmax-only;
AdamW) with CUDA events, L2 flushed between launches.  Run on the GPU box."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200 import _cabi as K
lib = K.load()
peak = 6533.8
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
st = lambda: torch.cuda.current_stream().cuda_stream

def timeit(fn, reps=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.sum()                        # read-only L2 flush (a write flush leaves dirty lines to evict)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

out = {}
for B in (64, 512):
    heads, g = 4, 14
    H = 16 * g
    grid = torch.randn(B, heads, g * g, device="cuda")
    ts_h = torch.empty(H, dtype=torch.int32); tw_h = torch.empty(H, 3)
    K.check(lib.bfk_upsample16_taps(g, ts_h.data_ptr(), tw_h.data_ptr()))
    ts, tw = ts_h.cuda(), tw_h.cuda()
    heat = torch.empty(heads, B, H, H, device="cuda")
    targets = [torch.rand(B, H, H, device="cuda") for _ in range(heads)]
    tptr = torch.tensor([t.data_ptr() for t in targets], dtype=torch.int64).cuda()
    loss = torch.empty(heads + 1, device="cuda"); key = torch.empty(heads * B, device="cuda", dtype=torch.int64)
    part = torch.empty(lib.bfk_heatmap_fwd_workspace_floats(B, heads, g), device="cuda")
    counters = torch.zeros(1 + heads * B, device="cuda", dtype=torch.int32)
    fa = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), tptr.data_ptr(), None, 0,
                          loss.data_ptr(), key.data_ptr(), part.data_ptr(), counters.data_ptr())
    ms = timeit(lambda: K.check(lib.bfk_heatmap_fwd(C.byref(fa), st())))
    by = heads * B * H * H * 4 * 2 + grid.numel() * 4
    out[f"heatmap_fwd_loss_argmax_B{B}"] = dict(us=ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / peak, bytes=by)
    fa2 = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), None, None, 0, None, key.data_ptr(), part.data_ptr(), counters.data_ptr())
    ms = timeit(lambda: K.check(lib.bfk_heatmap_fwd(C.byref(fa2), st())))
    by = heads * B * H * H * 4 + grid.numel() * 4
    out[f"heatmap_fwd_infer_argmax_B{B}"] = dict(us=ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / peak, bytes=by)
    gscale = torch.ones(heads + 1, device="cuda"); dgrid = torch.empty(B, heads, g * g, device="cuda")
    ws = torch.empty(lib.bfk_heatmap_bwd_workspace_floats(B, heads, g), device="cuda")
    ba = K.HeatmapBwdArgs(B, heads, g, 0, heat.data_ptr(), tptr.data_ptr(), None, 0, gscale.data_ptr(), 1.0 / (B * H * H), 1, None,
                          ts.data_ptr(), tw.data_ptr(), dgrid.data_ptr(), ws.data_ptr(), counters.data_ptr())
    ms = timeit(lambda: K.check(lib.bfk_heatmap_bwd(C.byref(ba), st())))
    by = heads * B * H * H * 4 * 2
    out[f"heatmap_bwd_bce_B{B}"] = dict(us=ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / peak, bytes=by)
    ms = timeit(lambda: K.check(lib.bfk_argmax2d(heat.data_ptr(), heads * B, H * H, key.data_ptr(), st())))
    by = heads * B * H * H * 4
    out[f"argmax2d_B{B}"] = dict(us=ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / peak, bytes=by)
    # head projection (reads fp32 tokens)
    D, Np, Nf = 768, g * g, 262
    X = torch.randn(B * Nf, D, device="cuda"); w_eff = torch.randn(heads, D, device="cuda"); b_eff = torch.randn(heads, device="cuda")
    pa = K.HeadProjectArgs(B, heads, Np, D, X.data_ptr(), Nf * D, Nf - Np, w_eff.data_ptr(), b_eff.data_ptr(), grid.data_ptr(),
                           None, Nf, None, None, 0, None, None, None)
    ms = timeit(lambda: K.check(lib.bfk_head_project_fwd(C.byref(pa), st())))
    by = B * Np * D * 4
    out[f"head_project_fwd_B{B}"] = dict(us=ms * 1e3, GBps=by / ms / 1e6, frac=by / ms / 1e6 / peak, bytes=by)
n = 60_000_000
p, gr, m, v = (torch.randn(n, device="cuda") for _ in range(4)); v.abs_()
aa = K.AdamWArgs(n, p.data_ptr(), gr.data_ptr(), m.data_ptr(), v.data_ptr(), None, 1e-4, 0.9, 0.999, 1e-8, 0.0, 0, 1, 1.0)
ms = timeit(lambda: K.check(lib.bfk_adamw(C.byref(aa), st())))
out["adamw_60M"] = dict(us=ms * 1e3, GBps=28 * n / ms / 1e6, frac=28 * n / ms / 1e6 / peak, bytes=28 * n)
for k_, v_ in out.items():
    print(f"{k_:36s} {v_['us']:9.1f} us  {v_['GBps']:8.1f} GB/s  {100 * v_['frac']:5.1f}% of {peak:.0f} GB/s (measured)  [{v_['bytes'] / 1e6:.1f} MB algorithmic]")
print(json.dumps(out))
