#!/usr/bin/env python
"""bench.py — BiFold policy hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus 1 --steps 10 --warmup 3            # our arm, 1 GPU
    torchrun ... bench.py --gpus N --steps K --warmup W       # our arm, N GPUs (batch-sharded DDP, weak scaling)
    python bench.py --impl reference --steps 2 --warmup 1     # CPU arm: the oracle restatement of the reference

Headline workload (BASELINE.json configs[1], "C2"): BiFold bimanual single-frame TRAINING step, per-GPU batch 64, 224x224 RGB +
1 instruction, SigLIP-base towers with LoRA (r=8) on q/v, 8-layer fusion transformer, 4 ConvDecoder heads, BCE loss, Adam.  bf16
operands on tcgen05, fp32 accumulate / residual / loss.  One step = forward + loss + backward + optimizer step over one
synthetic batch; metric: train samples/s (whole job).  The same JSON line carries, under "extra", short runs of the other
BASELINE configs (C3 context T=3 training, C4 batched inference, 384x384 reference-default resolution, C5 so400m-shaped towers),
so every config has a driver-visible number.  `--mode infer` makes inference the headline instead.

Inputs: the host sends the RAW form of a batch (uint8 frames, token ids, integer pick/place centres); the GPU input stage
(bifold_b200.data.InputStage) rebuilds the reference's sample dict (normalised fp32 frames, Gaussian target maps) inside the
captured step.  `value` has the raw batch resident in HBM; `e2e` copies it from pinned host memory every step.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mode", default="train", choices=["train", "infer"])
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--context", type=int, default=0, help="context frames T (0 = single-frame, config 2; 3 = config 3)")
    ap.add_argument("--image-size", type=int, default=224)
    ap.add_argument("--tower", default="base", choices=["base", "so400m"], help="so400m = config 5 (D=1152, 27 layers, MLP 4304, patch 16)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="bounded sample for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short runs of the other BASELINE configs")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: only the device-resident timed loop")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the captured CUDA graph")
    ap.add_argument("--fp32-inputs", action="store_true", help="feed the reference's fp32 sample dict (103 MB/step at B=64) instead of the raw batch")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# workload description
# ---------------------------------------------------------------------------------------------------
class W(dict):
    """One workload: mode, per-GPU batch, context frames, image size, tower family, precision."""
    __getattr__ = dict.__getitem__


def tower_dims(w):
    if w.tower == "so400m":     # so400m shape with patch 16 so that the x16 decoder emits image_size^2
        return dict(D=1152, L=27, heads=16, mlp=4304, name=f"google/siglip-so400m-patch16-{w.image_size}")
    return dict(D=768, L=12, heads=12, mlp=3072, name=f"google/siglip-base-patch16-{w.image_size}")


def model_cfg(w):
    from bifold_b200.config import Cfg
    t = tower_dims(w)
    cfg = Cfg(name="siglip_sequential" if w.context else "siglip", image_size=w.image_size, is_bimanual=True,
              requires_graph=False, patch_size=16, automodel_name=t["name"], dim=t["D"],
              emb_dropout=0.0, lora=True, r=8, lora_alpha=32, lora_dropout=0.01, target_modules=["q_proj", "v_proj"],
              threshold=0.5, text_encoder=None, pick_place_model="pick_place_convdecoder", fusion_model="concat_transformer",
              depth=8, heads=16, mlp_ratio=4, dropout=0.0, precision=w.precision,
              pretrained=False)     # random-init towers of that architecture (no network for checkpoints): stated in `data`
    if w.context:
        cfg["context_length"] = w.context
    return cfg


HEADS = ("left_pick", "left_place", "right_pick", "right_place")


def synth_raw(w, batch, seed, pinned=True):
    """RAW synthetic batch on the HOST: uint8 frames at model resolution, token ids, integer target centres
    (every 4th sample has no action for the right arm: all-zero maps, bifold/data/processor.py:296), context mask with
    i mod (T+1) leading ones (exercises the query-row mask quirk), a filled-ellipse cloth mask for get_action."""
    import torch
    from bifold_b200.ops import NO_ACTION
    g = torch.Generator().manual_seed(seed)
    H, T = w.image_size, w.context
    s = {"rgb_u8": torch.randint(0, 256, (batch, H, H, 3), generator=g, dtype=torch.uint8),
         "instruction": torch.randint(0, 32000, (batch, 64), generator=g)}
    if T:
        s["rgb_context_u8"] = torch.randint(0, 256, (batch, T, H, H, 3), generator=g, dtype=torch.uint8)
        s["context_attention_mask"] = (torch.arange(T)[None] < (torch.arange(batch) % (T + 1))[:, None]).long()
    if w.mode == "train":
        c = torch.randint(0, H, (4, batch, 2), generator=g, dtype=torch.int32)
        c[2:, 3::4, :] = NO_ACTION
        s["centers"] = c
    else:
        ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="ij")
        m = (((xs - H / 2) / (0.35 * H)) ** 2 + ((ys - H / 2) / (0.3 * H)) ** 2 <= 1).to(torch.uint8)
        s["mask_u8"] = m[None].repeat(batch, 1, 1)
    if pinned:
        s = {k: v.contiguous().pin_memory() for k, v in s.items()}
    return s


def flops_per_sample(w):
    """Algorithmic GEMM+attention FLOPs per sample (SURVEY.md §8d formulas): (train, forward).  Training = forward x 2 for
    the LoRA towers (fwd + dgrad; the base weights are frozen) and x 3 for the fully trainable fusion transformer."""
    t = tower_dims(w)
    D, M, L, Np, Lt = t["D"], t["mlp"], t["L"], (w.image_size // 16) ** 2, 64
    layer = lambda N, mlp: 8 * N * D * D + 4 * N * D * mlp + 4 * N * N * D
    T = w.context
    vision = (L * layer(Np, M) + 2 * Np * D * 768) * (T + 1)
    text = L * layer(Lt, M)
    Nf = 65 + (T + 1) * (Np + 1)
    fusion = 8 * layer(Nf, 4 * D)
    return 2 * (vision + text) + 3 * fusion, vision + text + fusion


def workload_config(w, world):
    t = tower_dims(w)
    return {"workload": f"BiFold bimanual {'context T=%d ' % w.context if w.context else 'single-frame '}"
                        f"{'training step (fwd+BCE+bwd+Adam)' if w.mode == 'train' else 'inference (heatmaps + per-arm argmax action)'}, "
                        f"SigLIP-{w.tower}-patch16-{w.image_size} towers (D={t['D']}, {t['L']} layers) + LoRA r=8, 8-layer fusion, 4 ConvDecoder heads",
            "per_gpu_batch": w.batch, "global_batch": w.batch * world, "image": w.image_size, "context_frames": w.context,
            "tower": w.tower, "parallelism": f"dp{world}",
            "inputs": "fp32 sample dict" if w.get("fp32_inputs") else "raw batch (uint8 frames, ids, int32 target centres) -> GPU input stage",
            "l2_policy": "working set per step (activations + weights, several GB) is far larger than the 126 MB L2; no explicit flush"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------
def oracle_spec(w):
    from oracle.synth import PolicySpec
    t = tower_dims(w)
    return PolicySpec(name="siglip_sequential" if w.context else "siglip", image_size=w.image_size, is_bimanual=True, lora=True,
                      lora_dropout=0.0, context_length=w.context, dim=t["D"], tower_layers=t["L"], tower_heads=t["heads"], tower_mlp=t["mlp"])


def time_oracle(w, batch, warmup, steps):
    """The CPU oracle restatement of the reference (oracle/policy_oracle.py), all host threads, bounded batch."""
    import torch
    from oracle import policy_oracle as po
    from oracle.synth import make_sample, make_state_dict, trainable_keys
    nthr = os.cpu_count() or 1
    torch.set_num_threads(nthr)
    spec = oracle_spec(w)
    sd = make_state_dict(spec)
    keys = trainable_keys(spec, sd)
    sample = make_sample(spec, batch, seed=0)
    state, times = {}, []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        if w.mode == "train":
            po.train_step(sd, sample, spec, keys, state, lr=1e-4)
        else:
            with torch.no_grad():
                po.get_action(po.policy_forward(sd, sample, spec), sample, spec)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    return batch / (ms / 1e3), ms, nthr


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    B = args.cpu_batch
    value, ms, nthr = time_oracle(w, B, args.warmup, args.steps)
    metric = "train_samples_per_s" if w.mode == "train" else "infer_actions_per_s"
    unit = "samples/s" if w.mode == "train" else "actions/s"
    wc = W(w); wc["batch"] = B; wc["fp32_inputs"] = True
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(wc, 1),
            "cpu_baseline": {"value": value, "unit": unit, "cores": nthr, "kind": "port",
                             "sample": f"{args.steps} steps of batch {B} (same model/shapes as the GPU workload; CPU oracle port of the "
                                       "reference, fp32 torch; it skips the SigLIP pooling head / logits the reference computes and discards)"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm: one workload
# ---------------------------------------------------------------------------------------------------
KERNEL_CLASSES = (
    ("attention", ("attn_",)),
    ("gemm", ("gemm_tc", "gemm_simt", "splitk_reduce", "lora_grads")),
    ("layernorm", ("ln_fwd", "ln_bwd", "lora_down_mma")),   # lora_down_mma: second launch of bfk_layernorm_fwd (LoRA down-projection over its output)
    ("tail", ("heatmap_", "head_project", "collapse_", "argmax2d", "action_kernel", "loss_finalize", "bce_prob", "gauss_targets")),
)


def classify(name):
    for cls, pats in KERNEL_CLASSES:
        if any(p in name for p in pats):
            return cls
    return "other"


def profile_kernels(fn, n):
    """Per-kernel device durations of `n` calls of fn (graph replays) through CUPTI (torch.profiler): {name: [launches, us]}."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type != torch.autograd.DeviceType.CUDA:
            continue
        us = float(e.time_range.elapsed_us())
        a = agg.setdefault(e.name, [0, 0.0])
        a[0] += 1
        a[1] += us
    return agg


def run_workload(w, dev, rank, world, steps, warmup, flags, main):
    """Builds the model of workload `w`, times `steps` steps (device-resident inputs and end-to-end from pinned host memory) and,
    for the main workload, the per-class roofline of the replayed graph.  Returns a dict of results."""
    import torch
    import torch.distributed as dist
    from bifold_b200.config import Cfg
    from bifold_b200.data import DevicePrefetcher, InputStage, ScalarReader
    from bifold_b200.graphs import GraphedAction, GraphedTrainStep
    from bifold_b200.losses import Losses
    from bifold_b200.models import Models
    from bifold_b200.optim import Optimizers

    torch.manual_seed(3407)
    model = Models.get_by_name(model_cfg(w), device=dev).to(dev)
    with torch.no_grad():     # LoRA B is zero at init (peft); give it values so the LoRA path carries signal in the timed step
        for n, p in model.named_parameters():
            if ".lora_B." in n:
                p.normal_(0, 0.02)
    train = w.mode == "train"
    model.train(train)
    eng = model._ensure_engine()
    B = w.batch
    stage = InputStage(HEADS, w.image_size, sigma=5.0)
    loss_fn = optim = ddp = None
    if train:
        loss_fn = Losses.get_by_name(cfg=Cfg(name="bce_gaussmap", is_bimanual=True, mask_pick_heatmap=False))
        optim = Optimizers.get_by_name(cfg=Cfg(name="adam", lr=1e-4, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                       params=filter(lambda p: p.requires_grad, model.parameters()))
        if world > 1:
            from bifold_b200.ddp import BatchShardedDDP
            ddp = BatchShardedDDP(model, optim)
    host = synth_raw(w, B, seed=1000 + rank)
    if flags.fp32_inputs:           # the reference's own sample dict built on the host side of the link (round-1 behaviour)
        with torch.no_grad():
            host = {k: v.cpu().contiguous().pin_memory() for k, v in stage({k: v.to(dev) for k, v in host.items()}).items()}
        stage = None
    dev_batch = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    use_graph = not flags.no_graph and w.precision == "bf16"        # fp32 = SIMT parity mode, eager only
    try:    # an instrumented eager step after a capture keeps old AccumulateGrad nodes alive: a benign warning
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    except Exception:
        pass

    gobj = None
    if train:
        if use_graph:
            gobj = GraphedTrainStep(model, loss_fn, optim, warmup=2, allow_distributed=world > 1, input_stage=stage)
            step = gobj
        else:
            def step(batch):
                s = stage(batch) if stage is not None else batch
                loss, _ = loss_fn(model(s), s)
                optim.zero_grad(); loss.backward(); optim.step()
                return loss
    else:
        if use_graph:
            gobj = GraphedAction(model, warmup=2, input_stage=stage)
            step = gobj.enqueue
        else:
            def step(batch):
                s = stage(batch) if stage is not None else batch
                return model.action_on_device(s)[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    Wm = max(warmup, 3) + (3 if use_graph else 0)      # the first 2 graph-mode calls are eager, the 3rd captures
    for _ in range(Wm):
        step(dev_batch)
    # ---- device-resident inputs: `value` ----
    w0, l0, o0 = dict(eng.k.work), eng.k.launches, (optim.launches if optim else 0)
    sampler = ClockSampler(dev.index or 0)
    if main:
        sampler.start()
    torch.cuda.profiler.start()      # no-op unless run under ncu --profile-from-start off
    ms_dev = timed(lambda: step(dev_batch), steps)
    torch.cuda.profiler.stop()
    clocks = sampler.stop() if main else None
    graphed = gobj is not None and gobj.graph is not None
    if graphed:
        launches = gobj.launches_per_replay * steps
        work = dict(gobj.work_per_replay)
    else:
        launches = (eng.k.launches - l0) + ((optim.launches - o0) if optim else 0)
        work = {k: (eng.k.work[k] - w0[k]) / steps for k in w0}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(dev_batch)                 # one step: the launch queue never fills, so this is host time only
    host_enqueue_ms = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()

    # ---- end to end through the public API with HOST (pinned) buffers: `e2e` ----
    # Every step copies its raw inputs from pinned host memory (bifold_b200.data.DevicePrefetcher: batch i+1 moves on a copy
    # stream while batch i is computed) and reads its result back (loss through a pinned ring one step late; for inference the
    # (heads,B,2) pixel table).  All copies and reads happen inside the timed region (the first copy is fully exposed).
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    heads = len(HEADS)
    d2h = 4 if train else heads * B * 2 * 4

    infer_ring = None if train else torch.empty(2, heads, B, 2, dtype=torch.int32).pin_memory()     # allocated once, like a serving loop would

    def e2e_run(n):
        pf = DevicePrefetcher((host for _ in range(n)), dev)
        if train:
            rd = ScalarReader(dev)
            for s in pf:
                rd.push(step(s))
            vals = rd.drain()
            assert len(vals) == n
        else:
            ring = infer_ring
            evs = [None, None]
            for i, s in enumerate(pf):
                xy = step(s)
                if evs[i % 2] is not None:
                    evs[i % 2].synchronize()
                ring[i % 2].copy_(xy, non_blocking=True)
                evs[i % 2] = torch.cuda.Event(); evs[i % 2].record()
            for e in evs:
                if e is not None:
                    e.synchronize()
        assert pf.bytes_copied == n * h2d

    if flags.skip_e2e:
        ms_e2e = ms_dev
    else:
        e2e_run(2)
        ms_e2e = timed(lambda: e2e_run(steps), 1)

    per_step = ms_dev / steps
    unit = "samples/s" if train else "actions/s"
    tr_flops, fw_flops = flops_per_sample(w)
    mf = tr_flops if train else fw_flops
    peaks = load_peaks()
    res = {"metric": "train_samples_per_s" if train else "infer_actions_per_s", "value": B * world / (per_step / 1e3), "unit": unit,
           "ms_per_step": per_step,
           "e2e": {"value": B * world / (ms_e2e / steps / 1e3), "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": ms_e2e / steps},
           "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
           "launch_mode": ("CUDA graph replay (whole step incl. input stage)" if graphed else
                           ("eager" if gobj is None or gobj.failed is None else "eager (graph capture failed: %s)" % gobj.failed[:200])),
           "model_flops_per_sample": mf, "model_tflops_per_gpu": mf * B / (per_step / 1e3) / 1e12,
           "model_frac_of_bf16_peak": mf * B / (per_step / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
           "clocks": clocks}

    # ---- roofline by kernel class: CUPTI durations of the replayed step (no launch gaps, no host in the windows) ----
    if main and not flags.no_roofline:
        try:
            agg = profile_kernels(lambda: step(dev_batch), 3)
            res.update(roofline_from_profile(agg, 3, work, per_step, peaks))
        except Exception as e:      # CUPTI unavailable on the box: say so instead of inventing numbers
            res["roofline"] = None
            res["roofline_error"] = f"{type(e).__name__}: {e}"[:300]
    if main and world > 1 and train:
        # replicas must stay bit-identical: max |p - p_rank0| over all trainable parameters after the timed steps
        drift = torch.zeros(1, device=dev)
        for st in optim._flat.values():
            ref = st["p"].clone()
            dist.broadcast(ref, src=0)
            drift = torch.maximum(drift, (st["p"] - ref).abs().max().reshape(1))
        dist.all_reduce(drift, op=dist.ReduceOp.MAX)
        res["ddp_param_drift"] = drift.item()
    # ---- tear-down: a live CUDA graph must not outlive the NCCL communicator / the next workload's allocations ----
    if gobj is not None:
        gobj.graph = None
        gobj.static_in = None
    del gobj, step, model, eng, optim, ddp, dev_batch
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return res


def roofline_from_profile(agg, n, work, per_step_ms, peaks):
    """by_class = {class: ms per step, achieved rate, fraction of its bound}; roofline = the dominant kernel class (tcgen05 GEMM)."""
    cls_ms, top = {}, []
    for name, (cnt, us) in agg.items():
        c = classify(name)
        cls_ms[c] = cls_ms.get(c, 0.0) + us / n / 1e3
        top.append((us / n / 1e3, cnt // n, name))
    top.sort(reverse=True)
    tf, hbm = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    by = {}
    for c, ms in sorted(cls_ms.items(), key=lambda kv: -kv[1]):
        e = {"ms": ms, "share_of_kernel_time": ms / max(sum(cls_ms.values()), 1e-9)}
        if c == "gemm" and work.get("gemm_flops"):
            e.update(bound="tensor", achieved=work["gemm_flops"] / (ms / 1e3) / 1e12, unit="TFLOP/s")
            e["frac"] = e["achieved"] / tf
        elif c == "attention" and work.get("attn_flops"):
            e.update(bound="tensor", achieved=work["attn_flops"] / (ms / 1e3) / 1e12, unit="TFLOP/s")
            e["frac"] = e["achieved"] / tf
        elif c == "layernorm" and work.get("ln_bytes"):
            e.update(bound="hbm", achieved=work["ln_bytes"] / (ms / 1e3) / 1e9, unit="GB/s")
            e["frac"] = e["achieved"] / hbm
        elif c == "tail" and work.get("tail_bytes"):
            e.update(bound="hbm", achieved=work["tail_bytes"] / (ms / 1e3) / 1e9, unit="GB/s")
            e["frac"] = e["achieved"] / hbm
        by[c] = e
    total = sum(cls_ms.values())
    out = {"roofline_by_class": {"classes": by, "kernel_ms_sum": total, "ms_per_step": per_step_ms,
                                 "sum_over_step": total / per_step_ms,
                                 "note": "CUPTI kernel durations of 3 replays of the captured step; the graph runs up to three branches "
                                         "concurrently (text tower, off-chain weight gradients), so the sum may exceed the step time",
                                 "top_kernels": [{"ms": round(ms, 4), "launches": c, "name": nm[:90]} for ms, c, nm in top[:14]]}}
    g = by.get("gemm")
    if g and "achieved" in g:
        out["roofline"] = {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel (tcgen05.mma kind::f16, TMA-fed, TMEM accumulators)",
                           "achieved": g["achieved"], "peak": tf, "unit": "TFLOP/s", "frac": g["frac"], "traffic": gemm_traffic(),
                           "gemm_ms_per_step": g["ms"], "gemm_flops_per_step": work["gemm_flops"],
                           "how": "sum of 2*M*N*K over the GEMM launches of one step / sum of their CUPTI kernel durations in the replayed graph",
                           "peak_source": peaks["source"] + ", sustained figure (kernels timed inside a long step)"}
    else:
        out["roofline"] = None
    return out


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        p["source"] = "measured (MEASURED_PEAKS.json)"
        return p
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def _profile_json(*names):
    for n in names:
        try:
            with open(os.path.join(ROOT, "profiles", n)) as f:
                return json.load(f), n
        except Exception:
            continue
    return None, None


def gemm_traffic():
    """DRAM bytes (read + write) per launch of the dominant GEMM kernel, from the committed `ncu --set full` capture
    (tools/ncu_traffic.py); None if absent."""
    t, n = _profile_json("r02_gemm_traffic.json", "r01_gemm_traffic.json")
    if t is None:
        return None
    return {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": t["algorithmic_bytes_per_launch"],
            "kernel": t["kernel"], "shape": t["shape"], "source": f"profiles/{n} (ncu --set full)"}


def tail_roofline(B, g, dev):
    """HBM roofline of the heatmap tail kernels at the bench batch (and at batch 512, BASELINE config 4): algorithmic bytes =
    heads*B*H*W*4 written + as much target read (+ the tiny grid) for the fused forward (x16 upsample + sigmoid + BCE + argmax).
    Durations are KERNEL durations from CUPTI (torch.profiler) of isolated launches with the L2 flushed before each one; the
    CUDA-event bracket of the same launches (which adds ~5 us of launch / event overhead to a 20-30 us kernel) is reported beside."""
    import ctypes as C
    import torch
    from bifold_b200 import _cabi as K
    lib = K.load()
    peaks = load_peaks()
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(192 * 1024 * 1024, device=dev, dtype=torch.uint8)

    def measure(B):
        heads, H = 4, 16 * g
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
        D, Np, Nf = 768, g * g, 65 + g * g + 1
        X = torch.randn(B * Nf, D, device=dev); w_eff = torch.randn(heads, D, device=dev); b_eff = torch.randn(heads, device=dev)
        pa = K.HeadProjectArgs(B, heads, Np, D, X.data_ptr(), Nf * D, Nf - Np, w_eff.data_ptr(), b_eff.data_ptr(), grid.data_ptr(),
                               None, Nf, None, None, 0, None, None, None)
        out = {}
        for nm, pat, call, nbytes in (
                ("heatmap_fwd", "heatmap_fwd_kernel", lambda: K.check(lib.bfk_heatmap_fwd(C.byref(fa), st)), heads * B * H * H * 4 * 2 + grid.numel() * 4),
                ("heatmap_bwd (BCE gradient -> dgrid)", "heatmap_bwd_stage1", lambda: K.check(lib.bfk_heatmap_bwd(C.byref(ba), st)), heads * B * H * H * 4 * 2),
                ("head_project_fwd (fp32 tokens -> 4 x g x g grids)", "head_project_fwd", lambda: K.check(lib.bfk_head_project_fwd(C.byref(pa), st)), B * Np * D * 4)):
            ev = []
            for it in range(8):
                flush.sum()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); call(); e1.record(); torch.cuda.synchronize()
                if it >= 3:
                    ev.append(e0.elapsed_time(e1))
            ev_ms = sorted(ev)[len(ev) // 2]
            k_ms = None
            try:
                agg = profile_kernels(lambda: (flush.sum(), call()), 5)
                hit = [(cnt, us) for name, (cnt, us) in agg.items() if pat in name]
                if hit:
                    k_ms = sum(us for _, us in hit) / max(sum(c for c, _ in hit), 1) / 1e3
            except Exception:
                k_ms = None
            t_ms = k_ms if k_ms else ev_ms
            out[nm] = {"us_per_launch": t_ms * 1e3, "us_per_launch_events": ev_ms * 1e3, "timer": "cupti" if k_ms else "events",
                       "bytes_per_launch": nbytes, "achieved": nbytes / t_ms / 1e6, "frac": nbytes / t_ms / 1e6 / peaks["hbm_gbs"]}
        return out

    main_b = measure(B)
    fwd = main_b.pop("heatmap_fwd")
    big = None
    if B != 512:
        try:
            big = {k: {kk: v[kk] for kk in ("us_per_launch", "achieved", "frac", "bytes_per_launch")} for k, v in measure(512).items()}
        except Exception as e:      # e.g. not enough free memory next to a large workload
            big = {"error": f"{type(e).__name__}: {e}"[:200]}
    t, n = _profile_json("r02_tail_traffic.json")
    traffic = None
    if t is not None:
        traffic = {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": t.get("algorithmic_bytes_per_launch"),
                   "kernel": t.get("kernel"), "shape": t.get("shape"), "source": f"profiles/{n} (ncu --set full)"}
    return {"bound": "hbm", "kernel": "heatmap_fwd_kernel (x16 upsample + sigmoid + BCE + argmax)", "achieved": fwd["achieved"],
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": fwd["frac"], "traffic": traffic, "bytes_per_launch": fwd["bytes_per_launch"],
            "us_per_launch": fwd["us_per_launch"], "us_per_launch_events": fwd["us_per_launch_events"], "timer": fwd["timer"],
            "batch": B, "peak_source": peaks["source"], "other_tail_kernels": main_b, "batch_512": big,
            "how": "isolated launches, L2 flushed before each; kernel durations from CUPTI (median-free mean of 5), event-bracket times beside"}


# ---------------------------------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    w = W(mode=args.mode, batch=args.batch, context=args.context, image_size=args.image_size, tower=args.tower, precision=args.precision,
          fp32_inputs=args.fp32_inputs)
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist
    from bifold_b200.build import build
    if local_rank == 0:
        build()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()

    res = run_workload(w, dev, rank, world, args.steps, max(args.warmup, 3), args, main=True)

    # ---- multi-GPU: what the gap to N x the 1-GPU step is made of.  Every rank times the SAME step locally (no collective, no
    #      barrier inside) while all the other GPUs of the box do the same, i.e. under the same power / thermal conditions as the
    #      data-parallel run: the lock-step job can never beat the slowest GPU (straggler share); what is left is the exposed part
    #      of the gradient all-reduces (+ their SM / HBM contention with the backward kernels). ----
    if world > 1 and args.mode == "train":
        import copy
        la = copy.copy(args); la.skip_e2e = True; la.no_roofline = True
        dist.barrier()
        lr = run_workload(w, dev, rank, 1, args.steps, max(args.warmup, 3), la, main=False)
        t = torch.tensor([lr["ms_per_step"]], device=dev)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        loc = [x.item() for x in allt]
        ddp_ms = res["ms_per_step"]
        res["scaling_diag"] = {"local_ms_per_rank": [round(x, 3) for x in loc], "local_mean_ms": sum(loc) / world, "local_max_ms": max(loc),
                               "ddp_ms": ddp_ms, "straggler_share": (max(loc) - sum(loc) / world) / ddp_ms,
                               "exposed_comm_share": (ddp_ms - max(loc)) / ddp_ms,
                               "how": "each rank replays the same captured step WITHOUT the all-reduces, all GPUs busy at the same time; "
                                      "straggler = (slowest - mean local step) / DDP step, exposed comm = (DDP step - slowest local step) / DDP step"}

    # ---- the other BASELINE configs, short runs in the same process (driver-visible numbers for every config) ----
    extra = {}
    if not args.no_extra:
        plans = [("C3_context_T3_train", W(w, mode="train", batch=16, context=3, image_size=224, tower="base")),
                 ("C4_infer_b64", W(w, mode="infer", batch=64, context=0, image_size=224, tower="base")),
                 ("C2_384_train", W(w, mode="train", batch=32, context=0, image_size=384, tower="base")),
                 ("C5_so400m_384_train", W(w, mode="train", batch=16, context=0, image_size=384, tower="so400m"))]
        if world == 1:
            plans.insert(2, ("C4_infer_b512", W(w, mode="infer", batch=512, context=0, image_size=224, tower="base")))
        for name, pw in plans:
            if all(pw[k] == w[k] for k in ("mode", "batch", "context", "image_size", "tower")):
                continue
            t0 = time.perf_counter()
            try:
                r = run_workload(pw, dev, rank, world, 3, 3, args, main=False)
                extra[name] = {"metric": r["metric"], "value": r["value"], "unit": r["unit"], "ms_per_step": r["ms_per_step"],
                               "e2e_value": r["e2e"]["value"], "h2d_bytes_per_step": r["e2e"]["h2d_bytes_per_step"],
                               "model_frac_of_bf16_peak": r["model_frac_of_bf16_peak"], "launch_mode": r["launch_mode"],
                               "host_enqueue_ms_per_step": r["host_enqueue_ms_per_step"], "steps": 3, "n_gpus": world,
                               "config": workload_config(pw, world), "wall_s": round(time.perf_counter() - t0, 1)}
            except Exception as e:
                extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
                gc.collect()
                torch.cuda.empty_cache()

    tail = None
    if rank == 0 and not args.no_roofline:
        try:
            tail = tail_roofline(w.batch, w.image_size // 16, dev)
        except Exception as e:
            tail = {"error": f"{type(e).__name__}: {e}"[:200]}

    if world > 1:
        # all ranks leave NCCL before rank 0 starts the CPU baseline (no rank spins in a barrier meanwhile); a watchdog ends the
        # process if the tear-down itself stalls
        sys.stdout.flush()
        threading.Thread(target=lambda: (time.sleep(600 if rank == 0 else 30), os._exit(0)), daemon=True).start()
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
    if rank != 0:
        return

    cpu_baseline = None
    if not args.no_cpu_baseline:
        v, ms, nthr = time_oracle(w, args.cpu_batch, 1, 2)
        cpu_baseline = {"value": v, "unit": res["unit"], "cores": nthr, "kind": "port",
                        "sample": f"2 timed steps (1 warm-up) of batch {args.cpu_batch}, same model and shapes, fp32 torch CPU oracle port of the "
                                  "reference (oracle/policy_oracle.py)"}
    line = {"metric": res["metric"], "value": res["value"], "unit": res["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision, "data": "synthetic (random uint8 frames / ids / target centres; random-init weights of the named architecture)",
            "config": workload_config(w, world)}
    for k in ("e2e", "gpu_launches", "host_enqueue_ms_per_step", "launch_mode", "clocks", "model_flops_per_sample", "model_tflops_per_gpu",
              "model_frac_of_bf16_peak", "roofline", "roofline_by_class", "roofline_error", "ddp_param_drift", "scaling_diag"):
        if k in res:
            line[k] = res[k]
    line.setdefault("roofline", None)
    line["roofline_tail"] = tail
    line["cpu_baseline"] = cpu_baseline
    line["extra"] = extra
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
