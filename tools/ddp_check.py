"""2-rank NCCL check: DDP gradients of the sharded batch == single-process gradients of the whole batch."""
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
spec = tiny_spec(name="siglip_sequential", is_bimanual=True, lora=True, context_length=2)
GB = 8
full = make_sample(spec, GB, seed=7)
loss_fn = Losses.get_by_name(cfg=Cfg(name="bce_gaussmap", is_bimanual=True, mask_pick_heatmap=False))


def run(sample, ddp, clip=False):
    model, _ = build_model(spec, "fp32", device="cuda")
    model.train()
    optim = Optimizers.get_by_name(cfg=Cfg(name="adam", lr=1e-3, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                   params=filter(lambda p: p.requires_grad, model.parameters()))
    wrap = None
    if ddp:
        from bifold_b200.ddp import BatchShardedDDP
        wrap = BatchShardedDDP(model, optim)
    s = to_device(sample)
    loss, _ = loss_fn(model(s), s)
    optim.zero_grad(); loss.backward()
    if clip:      # the reference Trainer's optional clip_grad_norm_ between backward and step (bifold/__main__.py:148-149)
        if wrap is not None:
            wrap.finish_reduce()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad], 0.05)
    optim.step()
    torch.cuda.synchronize()
    return {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}, optim

per = GB // world
shard = {k: v[rank * per:(rank + 1) * per] for k, v in full.items()}
p_ddp, opt = run(shard, True)
p_ref, _ = run(full, False)
worst = max((p_ddp[n] - p_ref[n]).abs().max().item() for n in p_ref)
print(f"rank {rank}: max |param_ddp - param_single| after one Adam step = {worst:.3e}; reduced {opt.pre_step_hook.__self__.bytes_reduced} bytes", flush=True)
assert worst < 5e-5
p_ddp, _ = run(shard, True, clip=True)
p_ref, _ = run(full, False, clip=True)
worst_c = max((p_ddp[n] - p_ref[n]).abs().max().item() for n in p_ref)
print(f"rank {rank}: with finish_reduce() + clip_grad_norm_: {worst_c:.3e}", flush=True)
assert worst_c < 5e-5
dist.barrier()
if rank == 0:
    print("DDP CHECK OK")
dist.destroy_process_group()
