#!/usr/bin/env python
"""bench.py — BiFold policy hot path on B200 (driver contract: see the task statement).

    python bench.py --gpus 1 --steps 10 --warmup 3            # our arm, 1 GPU
    torchrun ... bench.py --gpus N --steps K --warmup W       # our arm, N GPUs (batch-sharded DDP, weak scaling)
    python bench.py --impl reference --steps 2 --warmup 1     # CPU arm: the oracle restatement of the reference

Workload (BASELINE.json configs[1]): BiFold bimanual single-frame TRAINING step, per-GPU batch 64,
224x224 RGB + 1 instruction, SigLIP-base towers with LoRA (r=8) on q/v, 8-layer fusion transformer,
4 ConvDecoder heads, BCE loss, Adam.  bf16 operands on tcgen05, fp32 accumulate / residual / loss.
One step = forward + loss + backward + optimizer step over one synthetic batch.
Metric: train samples/s (whole job).  `--mode infer` reports inference actions/s instead.
"""
from __future__ import annotations

import argparse
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
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-batch", type=int, default=8, help="bounded sample for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true", help="profiling runs: only the device-resident timed loop")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying the captured whole-step CUDA graph")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
def model_cfg(args, device):
    from bifold_b200.config import Cfg
    cfg = Cfg(name="siglip_sequential" if args.context else "siglip", image_size=args.image_size, is_bimanual=True,
              requires_graph=False, patch_size=16, automodel_name=f"google/siglip-base-patch16-{args.image_size}", dim=768,
              emb_dropout=0.0, lora=True, r=8, lora_alpha=32, lora_dropout=0.01, target_modules=["q_proj", "v_proj"],
              threshold=0.5, text_encoder=None, pick_place_model="pick_place_convdecoder", fusion_model="concat_transformer",
              depth=8, heads=16, mlp_ratio=4, dropout=0.0, precision=args.precision)
    if args.context:
        cfg["context_length"] = args.context
    return cfg


def synth_batch(args, batch, seed, pinned=True):
    """Synthetic sample dict on the HOST (SURVEY.md §8d shapes; values do not need to match the oracle's)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    H = args.image_size
    s = {"rgb": torch.randn(batch, 3, H, H, generator=g),
         "instruction": torch.randint(0, 32000, (batch, 64), generator=g)}
    if args.context:
        s["rgb_context"] = torch.randn(batch, args.context, 3, H, H, generator=g)
        s["context_attention_mask"] = (torch.arange(args.context)[None] < (torch.arange(batch) % (args.context + 1))[:, None]).long()
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(H, dtype=torch.float32), indexing="ij")
    for h in ("left_pick", "left_place", "right_pick", "right_place"):
        c = torch.randint(0, H, (batch, 2), generator=g).float()
        s[f"{h}_heatmap"] = torch.exp(-((xs[None] - c[:, 0, None, None]) ** 2 + (ys[None] - c[:, 1, None, None]) ** 2) / 50.0)
    s["mask"] = (((xs - H / 2) / (0.35 * H)) ** 2 + ((ys - H / 2) / (0.3 * H)) ** 2 <= 1).float()[None, None].repeat(batch, 1, 1, 1)
    if pinned:
        s = {k: v.contiguous().pin_memory() for k, v in s.items()}
    return s


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


def train_flops_per_sample(args):
    """Algorithmic GEMM+attention FLOPs of one training step per sample (SURVEY.md §8d formulas):
    forward x (2 for LoRA towers: fwd + dgrad; 3 for the fully trainable fusion)."""
    D, M, Np, L = 768, 3072, (args.image_size // 16) ** 2, 64
    layer = lambda N: 8 * N * D * D + 4 * N * D * M + 4 * N * N * D
    T = args.context
    vision = (12 * layer(Np) + 2 * Np * D * 768) * (T + 1)
    text = 12 * layer(L)
    Nf = 65 + (T + 1) * (Np + 1)
    fusion = 8 * layer(Nf)
    return 2 * (vision + text) + 3 * fusion, vision + text + fusion


# ---------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the oracle restatement of the reference (oracle/policy_oracle.py), all host threads."""
    if rank != 0:
        return
    import torch
    from oracle import policy_oracle as po
    from oracle.synth import PolicySpec, make_sample, make_state_dict, trainable_keys
    torch.set_num_threads(os.cpu_count() or 1)
    spec = PolicySpec(name="siglip_sequential" if args.context else "siglip", image_size=args.image_size, is_bimanual=True,
                      lora=True, lora_dropout=0.0, context_length=args.context)
    sd = make_state_dict(spec)
    keys = trainable_keys(spec, sd)
    B = args.cpu_batch
    sample = make_sample(spec, B, seed=0)
    state = {}
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        if args.mode == "train":
            po.train_step(sd, sample, spec, keys, state, lr=1e-4)
        else:
            with torch.no_grad():
                out = po.policy_forward(sd, sample, spec)
                po.get_action(out, sample, spec)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = B / (ms / 1e3)
    metric = "train_samples_per_s" if args.mode == "train" else "infer_actions_per_s"
    unit = "samples/s" if args.mode == "train" else "actions/s"
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B, 1),
            "cpu_baseline": {"value": value, "unit": unit, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{args.steps} steps of batch {B} (same model/shapes as the GPU workload; CPU oracle port of the reference, fp32)"},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, per_gpu_batch, world):
    return {"workload": f"BiFold bimanual {'context T=%d ' % args.context if args.context else 'single-frame '}"
                        f"{'training step (fwd+BCE+bwd+Adam)' if args.mode == 'train' else 'inference (heatmaps + per-arm argmax action)'}, "
                        f"SigLIP-base-patch16-{args.image_size} towers + LoRA r=8, 8-layer fusion, 4 ConvDecoder heads",
            "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * world, "image": args.image_size,
            "context_frames": args.context, "parallelism": f"dp{world}",
            "l2_policy": "working set per step (activations + weights, several GB) is far larger than the 126 MB L2; no explicit flush"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
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
    from bifold_b200.losses import Losses
    from bifold_b200.models import Models
    from bifold_b200.optim import Optimizers
    from bifold_b200.config import Cfg

    torch.manual_seed(3407)
    model = Models.get_by_name(model_cfg(args, dev), device=dev).to(dev)
    # LoRA B is zero at init (peft); give it values so the LoRA path carries signal in the timed step
    with torch.no_grad():
        for n, p in model.named_parameters():
            if ".lora_B." in n:
                p.normal_(0, 0.02)
    loss_fn = Losses.get_by_name(cfg=Cfg(name="bce_gaussmap", is_bimanual=True, mask_pick_heatmap=False))
    optim = Optimizers.get_by_name(cfg=Cfg(name="adam", lr=1e-4, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                   params=filter(lambda p: p.requires_grad, model.parameters()))
    if world > 1:
        from bifold_b200.ddp import BatchShardedDDP
        BatchShardedDDP(model, optim)
    B = args.batch
    host = synth_batch(args, B, seed=1000 + rank)
    dev_sample = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    eng = model._ensure_engine()
    train = args.mode == "train"
    model.train(train)

    def eager_step(sample, read_back):
        if train:
            eng.step_seed += 1
            out = model(sample)
            loss, _ = loss_fn(out, sample)
            optim.zero_grad()
            loss.backward()
            optim.step()
            return loss.item() if read_back else loss
        action = model.get_action(sample)      # D2H of the (heads,B,2) pixel table happens inside
        return action

    # Training replays ONE captured CUDA graph per step (bifold_b200.graphs): the ~750 launches of the eager step cost
    # 14-19 ms of host time against ~22 ms of GPU work, so any host contention makes it launch-bound.  With N > 1 the
    # bucketed NCCL all-reduces of BatchShardedDDP are captured in the same graph (every rank captures and replays in
    # lock-step; tools/ddp_graph_check.py: bit-identical to the eager DDP step).  Inference uses the eager path.
    try:    # the extra instrumented eager step after the capture keeps old AccumulateGrad nodes alive: a benign warning
        torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
    except Exception:
        pass
    use_graph = train and not args.no_graph and args.precision == "bf16"     # fp32 = SIMT parity mode, eager only
    gstep = None
    if use_graph:
        from bifold_b200.graphs import GraphedTrainStep
        gstep = GraphedTrainStep(model, loss_fn, optim, warmup=2, allow_distributed=world > 1)

    def step(sample, read_back):
        if gstep is not None:
            loss = gstep(sample)
            return loss.item() if read_back else loss
        return eager_step(sample, read_back)

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

    W = max(args.warmup, 3)
    for _ in range(W):
        step(dev_sample, False)
    # ---- device-resident inputs: `value` ----
    l0, o0 = eng.k.launches, optim.launches
    sampler = ClockSampler(local_rank)
    sampler.start()
    torch.cuda.profiler.start()      # no-op unless run under ncu --profile-from-start off
    ms_dev = timed(lambda: step(dev_sample, False), args.steps)
    torch.cuda.profiler.stop()
    clocks = sampler.stop()
    launches = (eng.k.launches - l0) + (optim.launches - o0)
    if gstep is not None and gstep.graph is not None:
        launches = gstep.launches_per_replay * args.steps      # kernels inside the replayed graph (counted at capture)
    # host time to ENQUEUE one step (no synchronisation inside): how far the launch queue runs ahead of the GPU
    import time as _time
    torch.cuda.synchronize()
    _t0 = _time.perf_counter()
    step(dev_sample, False)                 # one step: the launch queue (~1k entries) never fills, so this is host time only
    host_enqueue_ms = (_time.perf_counter() - _t0) * 1e3
    torch.cuda.synchronize()
    # ---- end-to-end through the public API with HOST (pinned) buffers: `e2e` ----
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    # Every step copies its inputs from pinned host memory and reads its result (loss / action table) back.  The
    # copies go through the package's input pipeline (bifold_b200.data): batch i+1 moves on a copy stream while
    # batch i is computed, the loss comes back through pinned memory one step late.  All `steps` copies and reads
    # happen inside the timed region (the first copy is fully exposed).
    from bifold_b200.data import DevicePrefetcher, ScalarReader

    def e2e_run(n):
        pf = DevicePrefetcher((host for _ in range(n)), dev)
        rd = ScalarReader(dev)
        for s in pf:
            r = step(s, False)
            if train:
                rd.push(r)
        vals = rd.drain()
        assert pf.bytes_copied == n * h2d and (not train or len(vals) == n)
        return vals

    if args.skip_e2e:
        ms_e2e = ms_dev
    else:
        e2e_run(2)
        ms_e2e = timed(lambda: e2e_run(args.steps), 1)
    d2h = 4 if train else 4 * B * 2 * 4
    per_step = ms_dev / args.steps
    value = B * world / (per_step / 1e3)
    e2e_value = B * world / (ms_e2e / args.steps / 1e3)
    metric = "train_samples_per_s" if train else "infer_actions_per_s"
    unit = "samples/s" if train else "actions/s"

    # ---- roofline of the dominant kernel (tcgen05 GEMM), one extra instrumented step ----
    roofline = None
    if not args.no_roofline:
        if rank == 0:
            roofline = gemm_roofline(eng, lambda: eager_step(dev_sample, False))
        else:
            eager_step(dev_sample, False)      # the instrumented step contains the DDP collectives: every rank must run it
            torch.cuda.synchronize()
    tail = None
    if rank == 0 and not args.no_roofline:
        tail = tail_roofline(B, args.image_size // 16, dev)
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu_baseline = cpu_baseline_leg(args)
    if rank == 0:
        tr_flops, fw_flops = train_flops_per_sample(args)
        peaks = load_peaks()
        line = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": W,
                "ms_per_step": per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic", "config": workload_config(args, B, world),
                "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "host_enqueue_ms_per_step": host_enqueue_ms,
                "launch_mode": "whole-step CUDA graph replay" if (gstep is not None and gstep.graph is not None) else
                ("eager" if gstep is None or gstep.failed is None else "eager (graph capture failed: %s)" % gstep.failed[:200]), "clocks": clocks,
                "model_flops_per_sample": tr_flops if train else fw_flops,
                "model_tflops_per_gpu": (tr_flops if train else fw_flops) * B / (per_step / 1e3) / 1e12,
                "model_frac_of_bf16_peak": (tr_flops if train else fw_flops) * B / (per_step / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
                "roofline": roofline, "roofline_tail": tail, "cpu_baseline": cpu_baseline}
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tear-down order matters: NCCL must not destroy a communicator that a live CUDA graph still references.
        # A watchdog ends the process if the tear-down itself stalls (the result line is already printed).
        import gc
        import threading
        sys.stdout.flush()
        threading.Thread(target=lambda: (time.sleep(30), os._exit(0)), daemon=True).start()
        if gstep is not None:
            gstep.graph = None
            gstep.static_loss = None
            gstep.static_in = None
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        p["source"] = "measured (MEASURED_PEAKS.json)"
        return p
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def gemm_traffic():
    """DRAM bytes (read + write) per launch of the dominant GEMM kernel, from the committed `ncu --set full` capture of
    this bench command (profiles/r01_gemm_traffic.json, written by tools/ncu_traffic.py); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")) as f:
            t = json.load(f)
        return {"dram_bytes_per_launch": t["dram_bytes_per_launch"], "algorithmic_bytes_per_launch": t["algorithmic_bytes_per_launch"],
                "kernel": t["kernel"], "shape": t["shape"], "source": "profiles/r01_gemm_traffic.json (ncu --set full)"}
    except Exception:
        return None


def gemm_roofline(eng, step_fn):
    """CUDA-event timing of every tcgen05 GEMM launch of one extra (untimed-for-the-headline) step, on the
    launching stream: achieved = sum(2*M*N*K) / sum(duration)."""
    import torch
    k = eng.k
    recs = []
    orig = k.gemm

    def timed_gemm(A, B, D, M, N, Kd, *a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(A, B, D, M, N, Kd, *a, **kw)
        e1.record()
        recs.append((2.0 * M * N * Kd, e0, e1, A.dtype == torch.bfloat16))

    k.gemm = timed_gemm
    overlap = eng.overlap_towers
    eng.overlap_towers = False      # one stream: no concurrent text-tower kernels inside the per-launch event windows
    wgrad = eng.overlap_wgrad
    eng.overlap_wgrad = False
    try:
        torch.cuda.synchronize()
        torch.cuda._sleep(int(0.03 * 1.9e9))   # ~30 ms head start for the host: the event windows then hold kernels only
        step_fn()
        torch.cuda.synchronize()
    finally:
        k.gemm = orig
        eng.overlap_towers = overlap
        eng.overlap_wgrad = wgrad
    fl = sum(r[0] for r in recs if r[3])
    ms = sum(r[1].elapsed_time(r[2]) for r in recs if r[3])
    peaks = load_peaks()
    if not recs or ms <= 0:
        return None
    ach = fl / (ms / 1e3) / 1e12
    peak = peaks["bf16_tflops_sustained"]
    return {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05.mma kind::f16, TMA-fed, TMEM accumulators)", "achieved": ach,
            "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": gemm_traffic(), "launches": len(recs),
            "gemm_ms_per_step": ms, "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)"}


def tail_roofline(B, g, dev):
    """HBM roofline of the fused heatmap tail kernel (x16 upsample + sigmoid + BCE + argmax) at the bench batch:
    algorithmic bytes = heads*B*H*W*4 written + as much target read (+ the tiny grid), CUDA events, L2 flushed."""
    import ctypes as C
    import torch
    from bifold_b200 import _cabi as K
    lib = K.load()
    heads, H = 4, 16 * g
    grid = torch.randn(B, heads, g * g, device=dev)
    ts_h = torch.empty(H, dtype=torch.int32); tw_h = torch.empty(H, 3)
    K.check(lib.bfk_upsample16_taps(g, ts_h.data_ptr(), tw_h.data_ptr()))
    ts, tw = ts_h.to(dev), tw_h.to(dev)
    heat = torch.empty(heads, B, H, H, device=dev)
    targets = [torch.rand(B, H, H, device=dev) for _ in range(heads)]
    tptr = torch.tensor([t.data_ptr() for t in targets], dtype=torch.int64).to(dev)
    loss = torch.empty(heads + 1, device=dev); key = torch.empty(heads * B, device=dev, dtype=torch.int64)
    part = torch.empty(heads * B * lib.bfk_heatmap_bands(g), device=dev)
    fa = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), tptr.data_ptr(), None, 0,
                          loss.data_ptr(), key.data_ptr(), part.data_ptr())
    flush = torch.empty(192 * 1024 * 1024, device=dev, dtype=torch.uint8)
    st = torch.cuda.current_stream().cuda_stream
    ts_ms = []
    for it in range(8):
        flush.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); K.check(lib.bfk_heatmap_fwd(C.byref(fa), st)); e1.record(); torch.cuda.synchronize()
        if it >= 3:
            ts_ms.append(e0.elapsed_time(e1))
    ms = sorted(ts_ms)[len(ts_ms) // 2]
    by = heads * B * H * H * 4 * 2 + grid.numel() * 4
    peaks = load_peaks()
    ach = by / ms / 1e6
    return {"bound": "hbm", "kernel": "heatmap_fwd_kernel (x16 upsample + sigmoid + BCE + argmax)", "achieved": ach,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "traffic": None, "bytes_per_launch": by,
            "us_per_launch": ms * 1e3, "peak_source": peaks["source"], "note": "see profiles/r01_tail_roofline.txt for B=512 and the other tail kernels"}


def cpu_baseline_leg(args):
    """The oracle port of the reference timed on this box's host cores (bounded sample)."""
    import torch
    from oracle import policy_oracle as po
    from oracle.synth import PolicySpec, make_sample, make_state_dict, trainable_keys
    nthr = os.cpu_count() or 1
    torch.set_num_threads(nthr)
    spec = PolicySpec(name="siglip_sequential" if args.context else "siglip", image_size=args.image_size, is_bimanual=True,
                      lora=True, lora_dropout=0.0, context_length=args.context)
    sd = make_state_dict(spec)
    keys = trainable_keys(spec, sd)
    B = args.cpu_batch
    sample = make_sample(spec, B, seed=0)
    state, times = {}, []
    for it in range(3):
        t0 = time.perf_counter()
        if args.mode == "train":
            po.train_step(sd, sample, spec, keys, state, lr=1e-4)
        else:
            with torch.no_grad():
                po.get_action(po.policy_forward(sd, sample, spec), sample, spec)
        if it >= 1:
            times.append(time.perf_counter() - t0)
    v = B / (sum(times) / len(times))
    return {"value": v, "unit": "samples/s" if args.mode == "train" else "actions/s", "cores": nthr, "kind": "port",
            "sample": f"2 timed steps (1 warm-up) of batch {B}, same model and shapes, fp32 torch CPU oracle of the reference"}


if __name__ == "__main__":
    main()
