"""Where the fixed cost of an end-to-end run goes: host timestamps around the phases of bench.py's e2e loop (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from bifold_b200.config import Cfg
from bifold_b200.data import DevicePrefetcher, InputStage, ScalarReader
from bifold_b200.graphs import GraphedTrainStep
from bifold_b200.losses import Losses
from bifold_b200.models import Models
from bifold_b200.optim import Optimizers
w = bench.W(mode="train", batch=64, context=0, image_size=224, tower="base", precision="bf16", fp32_inputs=False)
dev = torch.device("cuda", 0)
model = Models.get_by_name(bench.model_cfg(w), device=dev).to(dev); model.train()
loss_fn = Losses.get_by_name(cfg=Cfg(name="bce_gaussmap", is_bimanual=True, mask_pick_heatmap=False))
optim = Optimizers.get_by_name(cfg=Cfg(name="adam", lr=1e-4, betas=[0.9, 0.999], eps=1e-8, weight_decay=0), params=filter(lambda p: p.requires_grad, model.parameters()))
stage = InputStage(bench.HEADS, 224, sigma=5.0)
step = GraphedTrainStep(model, loss_fn, optim, warmup=2, input_stage=stage)
host = bench.synth_raw(w, 64, seed=1000)
devb = {k: v.to(dev) for k, v in host.items()}
for _ in range(8): step(devb)
torch.cuda.synchronize()
def run(n, tag):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pf = DevicePrefetcher((host for _ in range(n)), dev); t1 = time.perf_counter()
    rd = ScalarReader(dev); t2 = time.perf_counter()
    ts = []
    for s in pf:
        rd.push(step(s)); ts.append(time.perf_counter())
    t3 = time.perf_counter(); rd.drain(); t4 = time.perf_counter()
    e1.record(); torch.cuda.synchronize()
    print(f"{tag}: n={n} gpu {e0.elapsed_time(e1):.2f} ms | prefetcher init {1e3*(t1-t0):.2f} reader {1e3*(t2-t1):.2f} loop {1e3*(t3-t2):.2f} drain {1e3*(t4-t3):.2f} | first 3 iters {[round(1e3*(b-a),2) for a,b in zip([t2]+ts[:2], ts[:3])]}", flush=True)
run(2, "warm"); run(10, "e2e"); run(10, "e2e"); run(20, "e2e")
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): step(devb)
e1.record(); torch.cuda.synchronize(); print(f"resident: 10 steps gpu {e0.elapsed_time(e1):.2f} ms")
