"""2-rank NCCL check: the whole DDP step (incl. the bucketed all-reduces) captured in a CUDA graph == eager DDP step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from bifold_b200.config import Cfg
from bifold_b200.losses import Losses
from bifold_b200.optim import Optimizers
from oracle.synth import tiny_spec, make_sample
from tests.util import build_model, to_device

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
spec = tiny_spec(name="siglip", is_bimanual=True, lora=True)
GB = 8
per = GB // world
samples = []
for i in range(3):
    full = make_sample(spec, GB, seed=30 + i)
    samples.append(to_device({k: v[rank * per:(rank + 1) * per] for k, v in full.items()}))
loss_fn = Losses.get_by_name(cfg=Cfg(name="bce_gaussmap", is_bimanual=True, mask_pick_heatmap=False))


def run(graph):
    from bifold_b200.ddp import BatchShardedDDP
    from bifold_b200.graphs import GraphedTrainStep
    model, _ = build_model(spec, "bf16", device="cuda")
    model.train()
    optim = Optimizers.get_by_name(cfg=Cfg(name="adam", lr=1e-3, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                   params=filter(lambda p: p.requires_grad, model.parameters()))
    BatchShardedDDP(model, optim)
    optim.enable_device_schedule()
    step = GraphedTrainStep(model, loss_fn, optim, warmup=2, allow_distributed=True) if graph else None
    losses = []
    for i in range(6):
        s = samples[i % 3]
        if step is not None:
            losses.append(step(s).item())
        else:
            loss, _ = loss_fn(model(s), s)
            optim.zero_grad(); loss.backward(); optim.step()
            losses.append(loss.item())
    torch.cuda.synchronize()
    if step is not None:
        print(f"rank {rank}: graph captured={step.graph is not None} failed={step.failed} launches/replay={step.launches_per_replay}", flush=True)
    return losses, {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}

le, pe = run(False)
lg, pg = run(True)
worst = max((pe[n] - pg[n]).abs().max().item() for n in pe)
print(f"rank {rank}: losses eager {['%.6f' % x for x in le]} graph {['%.6f' % x for x in lg]}; max |param diff| {worst:.3e}", flush=True)
assert all(abs(a - b) <= 1e-5 * abs(a) for a, b in zip(le, lg)) and worst < 1e-6
dist.barrier()
if rank == 0:
    print("DDP GRAPH CHECK OK")
dist.destroy_process_group()
