"""Whole-step CUDA graph for training: forward + loss + backward + optimizer as ONE graph launch.

The eager step issues ~750 kernel launches from Python (≈14 ms of host time per step on an idle host, against ≈23 ms
of GPU work): any host contention makes the step launch-bound.  A captured graph replays the same kernels with one
launch.  What changes between replays lives in device memory:

* inputs       - copied into static tensors before the replay;
* dropout mask - every kernel that draws one adds a device counter to its seed (``Kernels.seed_dev``), advanced by a
                 captured increment;
* Adam         - step count, bias corrections and learning rate come from device scalars (``FusedAdam`` device
                 schedule, ``bfk_adamw_advance``); the host pushes learning-rate changes before a replay.

The fork/join side streams (text tower, off-chain weight gradients) become parallel branches of the graph; with
``allow_distributed=True`` the bucketed NCCL all-reduces of ``BatchShardedDDP`` are captured on their own branch too (every rank
captures and replays in lock-step).  An optional ``input_stage`` (``bifold_b200.data.InputStage``) is recorded in front of the
forward, so the static inputs are the raw uint8 frames and integer target centres.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import os

import torch


class GraphedTrainStep:
    """``step(sample) -> loss`` (a device scalar, valid until the next call).

    Drop references to losses of earlier EAGER steps before the capturing call: a live autograd graph keeps its AccumulateGrad
    nodes bound to the stream they were created on, and autograd then synchronises the capture stream with it (capture fails;
    the step object falls back to the eager path and records the reason in ``failed``).

    The first ``warmup`` calls run eagerly (they also build every cache the capture must not touch: tensor maps,
    pointer tables, split-K workspaces, optimizer flat buffers); the next call captures; later calls replay.
    ``loss_fn(out, sample)`` must return ``(loss, aux)`` like ``bifold.losses`` does.
    """

    def __init__(self, model, loss_fn: Callable, optimizer, warmup: int = 3, allow_distributed: bool = False,
                 input_stage: Optional[Callable] = None):
        eng = model._ensure_engine()
        # a model wired to BatchShardedDDP (its engine announces gradient groups to the reducer) captures NCCL all-reduces: the
        # caller must say so; a purely local model inside a multi-process job (bench.py's scaling diagnostic) is not affected
        if getattr(eng, "grad_ready_hook", None) is not None and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size() > 1 and not allow_distributed:
            raise RuntimeError("GraphedTrainStep: pass allow_distributed=True to capture the NCCL all-reduces of "
                               "BatchShardedDDP inside the graph (every rank must capture and replay in lock-step)")
        if eng.precision != "bf16" and eng.cfg.lora and eng.cfg.lora_dropout > 0:
            # the fp32 parity kernels (bfk_lora_down & co) take their dropout seed by value: a replay would reuse the mask
            raise RuntimeError("GraphedTrainStep needs precision='bf16' when LoRA dropout is active")
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.input_stage = input_stage      # e.g. bifold_b200.data.InputStage: raw uint8 frames / target centres -> sample, in-graph
        self.warmup = max(2, int(warmup))
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_in: Optional[Dict[str, torch.Tensor]] = None
        self.static_loss: Optional[torch.Tensor] = None
        self.launches_per_replay = 0
        self.work_per_replay: Dict[str, float] = {}     # algorithmic work per replay by roofline class (Kernels.work deltas at capture)
        self.failed: Optional[str] = None
        optimizer.enable_device_schedule()

    def _eager(self, sample):
        if self.input_stage is not None:
            sample = self.input_stage(sample)
        out = self.model(sample)
        loss, _ = self.loss_fn(out, sample)
        self.optimizer.zero_grad()
        loss.backward()
        self.optimizer.step()
        return loss

    def _stage(self, sample):
        if self.static_in is None:
            self.static_in = {k: torch.empty_like(v) if torch.is_tensor(v) else v for k, v in sample.items()}
        for k, v in sample.items():
            if torch.is_tensor(v):
                if v.data_ptr() != self.static_in[k].data_ptr():
                    self.static_in[k].copy_(v, non_blocking=True)
            else:
                self.static_in[k] = v
        return self.static_in

    def __call__(self, sample) -> torch.Tensor:
        self.calls += 1
        eng = self.model._ensure_engine()
        if self.calls <= self.warmup:
            s = self._stage(sample)              # warm up on the static tensors: their addresses go into cached tables
            loss = self._eager(s)
            eng.k.seed_dev.add_(1)
            return loss
        if self.failed is not None:                # capture was refused once: stay on the eager path (still correct)
            loss = self._eager(self._stage(sample))
            eng.k.seed_dev.add_(1)
            return loss
        if self.graph is None:
            s = self._stage(sample)
            self.optimizer.sync_lr()             # a pending learning-rate change must not be recorded as a graph node
            torch.cuda.synchronize()
            l0, o0, w0 = eng.k.launches, self.optimizer.launches, dict(eng.k.work)
            graph = torch.cuda.CUDAGraph()
            steps_before = {gi: st["step"] for gi, st in self.optimizer._flat.items()}
            try:
                # BFK_GRAPH_PRIORITY=1 (A/B knob): capture the dependency chain on a high-priority stream, so its kernel nodes win
                # the block scheduler against the off-chain branches (weight gradients, text tower) whenever both are ready
                cap_stream = torch.cuda.Stream(priority=-1) if os.environ.get("BFK_GRAPH_PRIORITY") == "1" else None
                with torch.cuda.graph(graph, stream=cap_stream):
                    self.static_loss = self._eager(s)
                    eng.k.seed_dev.add_(1)       # the next replay draws new dropout masks
            except Exception as e:               # e.g. an allocator / driver that refuses the capture
                self.failed = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()
                for gi, st in self.optimizer._flat.items():      # the recorded step never ran: restore the host counters
                    back = st["step"] - steps_before.get(gi, st["step"])
                    st["step"] -= back
                    for p in st["ps"]:
                        self.optimizer.state[p]["step"] -= back
                loss = self._eager(s)            # stay on the eager path from here on (still correct)
                eng.k.seed_dev.add_(1)
                return loss
            self.graph = graph
            self.launches_per_replay = (eng.k.launches - l0) + (self.optimizer.launches - o0)
            self.work_per_replay = {k: eng.k.work[k] - w0[k] for k in w0}
            # capture does not execute: run the step that was just recorded
            self.graph.replay()
            # (the capture pass already advanced the optimizer's host-side counters for this step)
            return self.static_loss
        self._stage(sample)
        self.optimizer.sync_lr()
        self.graph.replay()
        self.optimizer.note_replayed_step()
        return self.static_loss


class GraphedAction:
    """``get_action`` for a fixed batch shape as ONE CUDA-graph launch (config 4: batched inference).

    The eager ``model.get_action`` issues ~330 launches (≈8.7 ms of host time at B = 64 against ≈8.5 ms of GPU work, so the
    host paces it).  Here forward + per-head argmax + mask snap + arm gate are captured once; a call copies the inputs into
    static tensors, replays, and reads the (heads, B, 2) pixel table back — the only D2H traffic, as in
    ``BaseModel.get_action`` (bifold/models/__init__.py:79-173 returns numpy pixel arrays).
    """

    def __init__(self, model, warmup: int = 2, input_stage: Optional[Callable] = None):
        self.model, self.input_stage = model, input_stage
        self.warmup = max(1, int(warmup))
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.static_in: Optional[Dict[str, torch.Tensor]] = None
        self.static_xy = self.static_conf = None
        self.launches_per_replay = 0
        self.work_per_replay: Dict[str, float] = {}
        self.failed: Optional[str] = None

    _stage = GraphedTrainStep._stage

    def _eager(self, sample):
        if self.input_stage is not None:
            sample = self.input_stage(sample)
        xy, conf, _ = self.model.action_on_device(sample)
        return xy, conf

    def enqueue(self, sample):
        """Stage the inputs and launch; returns the device pixel table (valid until the next call)."""
        self.calls += 1
        if self.model.training:
            raise RuntimeError("GraphedAction captures the inference path: call model.eval() first")
        s = self._stage(sample)
        if self.calls <= self.warmup or self.failed is not None:
            self.static_xy, self.static_conf = self._eager(s)
            return self.static_xy
        if self.graph is None:
            eng = self.model._ensure_engine()
            torch.cuda.synchronize()
            l0, w0 = eng.k.launches, dict(eng.k.work)
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph):
                    self.static_xy, self.static_conf = self._eager(s)
            except Exception as e:
                self.failed = f"{type(e).__name__}: {e}"
                torch.cuda.synchronize()
                self.static_xy, self.static_conf = self._eager(s)
                return self.static_xy
            self.graph = graph
            self.launches_per_replay = eng.k.launches - l0
            self.work_per_replay = {k: eng.k.work[k] - w0[k] for k in w0}
        self.graph.replay()
        return self.static_xy

    def __call__(self, sample):
        return self.model.action_from_table(self.enqueue(sample))
