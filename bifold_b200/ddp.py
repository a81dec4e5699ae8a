"""Batch-sharded data parallelism for the policy (one process per GPU, torch.distributed / NCCL).

The reference is single-process; this is the one parallel strategy the hot path
admits: samples are independent in the forward, the loss is a mean over B*H*W, so with
equal per-rank batches the mean of the rank gradients is the global-batch gradient.

The engine hands out all parameter gradients as views of ONE flat fp32 arena and announces, during
its explicit backward, when a group of them is final ("decoders", "fusion.7" ... "fusion.0", "tokens",
"text", "vision").  Each announcement all-reduces that group's arena ranges on a side stream while
the backward keeps computing; the optimizer waits for the outstanding work and folds 1/world_size
into its fused Adam kernel (``grad_scale``), so there is no separate averaging pass.
Inference is batch-sharded with no collective.
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import torch
import torch.distributed as dist


def plan_groups(names: List[str], sizes: List[int]) -> Dict[str, List[Tuple[int, int]]]:
    """Arena ranges (offset, length in floats; parameters padded to 8 floats) of every gradient group."""
    groups: Dict[str, List[Tuple[int, int]]] = {}
    off = 0
    for n, s in zip(names, sizes):
        padded = (s + 7) // 8 * 8
        if "_decoder." in n or ".mask_head." in n:
            g = "decoders"
        elif ".transformer_encoder.layers." in n:
            g = "fusion." + n.split(".transformer_encoder.layers.")[1].split(".")[0]
        elif ".text_model." in n:
            g = "text"
        elif ".vision_model." in n:
            g = "vision"
        else:
            g = "tokens"
        r = groups.setdefault(g, [])
        if r and r[-1][0] + r[-1][1] == off:
            r[-1] = (r[-1][0], r[-1][1] + padded)
        else:
            r.append((off, padded))
        off += padded
    return groups


class GradReducer:
    """All-reduce (sum) of arena ranges as soon as the engine reports them final."""

    def __init__(self, names: List[str], sizes: List[int], process_group=None, comm_stream=None):
        self.groups = plan_groups(names, sizes)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.comm_stream = comm_stream
        self.pending = []
        self.arena = None
        self.bytes_reduced = 0
        self.n_collectives = 0
        self.deferred: List[Tuple[int, int]] = []
        # bucket_bytes: 0 = one all-reduce per announced group range; > 0 = bucketed (BatchShardedDDP picks the default,
        # BFK_DDP_BUCKET_MB overrides)
        self.bucket_bytes = int(float(os.environ.get("BFK_DDP_BUCKET_MB", "0")) * (1 << 20))

    def bind(self, arena: torch.Tensor):
        self.arena = arena

    def on_ready(self, group: str):
        """Called by the engine's backward when every gradient of `group` is final.  With ``bucket_bytes`` > 0 the ranges are
        collected until that many bytes are pending (or the fusion part ends with "tokens": the tower backward that follows hides
        the transfer) and adjacent arena ranges merge into one all-reduce: fewer, larger collectives."""
        if self.world == 1 or self.arena is None:
            return
        if self.bucket_bytes > 0:
            self.deferred.extend(self.groups.get(group, []))
            if 4 * sum(n for _, n in self.deferred) >= self.bucket_bytes or group == "tokens":
                self.flush()
            return
        self._reduce(self.groups.get(group, []))

    def flush(self):
        if self.deferred:
            merged = []
            for off, n in sorted(self.deferred):
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + n)
                else:
                    merged.append((off, n))
            self.deferred = []
            self._reduce(merged)

    def _reduce(self, ranges):
        for off, n in ranges:
            view = self.arena[off:off + n]
            if view.is_cuda:
                ev = torch.cuda.Event()
                ev.record()
                self.comm_stream.wait_event(ev)
                with torch.cuda.stream(self.comm_stream):
                    w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            else:
                w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
            self.pending.append(w)
            self.bytes_reduced += 4 * n
            self.n_collectives += 1

    def wait(self):
        self.flush()
        for w in self.pending:
            w.wait()
        if self.pending and self.comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self.pending = []


class BatchShardedDDP:
    """Wires a model (bifold_b200.models.SigLip*) and its FusedAdam for data-parallel training."""

    def __init__(self, model, optimizer, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.model, self.optimizer = model, optimizer
        eng = model._ensure_engine()
        names = eng.trainable_names()
        sizes = [eng.params[n].numel() for n in names]
        self.reducer = GradReducer(names, sizes, process_group, torch.cuda.Stream())
        if "BFK_DDP_BUCKET_MB" not in os.environ:
            # Measured on 2 x B200 (profiles/r02_ddp_buckets.txt): one all-reduce per fusion layer (8 x 28 MB) costs 3.5 % of
            # the step, ONE all-reduce of everything announced up to "tokens" 1.7 %: with LoRA towers ~5 ms of tower backward
            # follow and hide it.  Frozen towers have nothing behind the fusion backward: 64 MB buckets overlap with it instead.
            self.reducer.bucket_bytes = (1 << 40) if eng.cfg.lora else (64 << 20)
        self.world = self.reducer.world
        # identical weights on every rank
        for p in model.parameters():
            dist.broadcast(p.data, src=0, group=process_group)
        self.engine = eng

        def hook(group):
            if group == "decoders":                         # first announcement of every backward
                optimizer.grad_scale = 1.0 / self.world     # (finish_reduce() of the previous step may have reset it)
            self.reducer.bind(eng.last_arena)               # the arena is allocated at the start of Engine.backward
            self.reducer.on_ready(group)

        eng.grad_ready_hook = hook
        optimizer.grad_scale = 1.0 / self.world
        optimizer.pre_step_hook = self.reducer.wait

    def finish_reduce(self):
        """Wait for the outstanding all-reduces and turn the summed gradients into the global-batch MEAN in place.

        Between ``loss.backward()`` and ``optimizer.step()`` the ``.grad`` tensors are otherwise in flight on the
        communication stream and hold un-averaged sums (the 1/world factor normally rides inside the Adam kernel).
        Call this before anything that reads gradients there — the reference Trainer's optional
        ``clip_grad_norm_`` (bifold/__main__.py:148-149), gradient logging, a debugger.  ``optimizer.step()`` afterwards
        applies the gradients as they are (scale 1)."""
        self.reducer.wait()
        arena = getattr(self.engine, "last_arena", None)
        if arena is not None and self.world > 1 and self.optimizer.grad_scale != 1.0:
            arena.mul_(1.0 / self.world)
            self.optimizer.grad_scale = 1.0

    def shard(self, global_batch: int, rank: int = None):
        rank = dist.get_rank() if rank is None else rank
        per = global_batch // self.world
        return rank * per, per
