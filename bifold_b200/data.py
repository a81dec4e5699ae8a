"""Host -> device input pipeline for the policy step.

The reference feeds its model from a torch DataLoader and moves every batch with ``.to(device)`` on the
compute stream (bifold/__main__.py:328-331, the per-batch ``sample[k] = val.to(self.device)`` loop).  On a B200 that copy
(about 100 MB per 64-sample bimanual batch) costs ~2 ms of an otherwise GPU-bound 25 ms step, so this
module copies batch i+1 on a dedicated copy stream while batch i is being computed, and reads scalar
results (the loss) back through pinned memory one step late so the host never stalls the launch queue.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional

import torch


class DevicePrefetcher:
    """Iterates device-resident batches; the copy of the next host batch overlaps the current step.

    ``batches`` yields dicts of (ideally pinned) host tensors with fixed shapes.  Exactly one host->device copy
    is issued per yielded batch, one batch ahead of the consumer, into one of two persistent device buffer sets
    (no allocation in steady state).  A yielded batch is valid until the next ``__next__`` call returns.
    """

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device: torch.device):
        self.it: Iterator[Dict[str, torch.Tensor]] = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bytes_copied = 0
        self._bufs: List[Optional[Dict[str, torch.Tensor]]] = [None, None]
        self._free: List[Optional[torch.cuda.Event]] = [None, None]    # consumer finished enqueuing work on the slot
        self._ready: List[Optional[torch.cuda.Event]] = [None, None]   # copy into the slot finished
        self._i = 0            # index of the next batch to hand out
        self._pending = self._issue(0)

    def _issue(self, slot: int) -> bool:
        try:
            host = next(self.it)
        except StopIteration:
            return False
        bufs = self._bufs[slot]
        if bufs is None or any(torch.is_tensor(v) and (k not in bufs or bufs[k].shape != v.shape or bufs[k].dtype != v.dtype)
                               for k, v in host.items()):
            bufs = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host.items() if torch.is_tensor(v)}
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))   # allocation happened on the consumer's stream
        with torch.cuda.stream(self.copy_stream):
            if self._free[slot] is not None:
                self.copy_stream.wait_event(self._free[slot])
            out = {}
            for k, v in host.items():
                if torch.is_tensor(v):
                    bufs[k].copy_(v, non_blocking=True)
                    self.bytes_copied += v.numel() * v.element_size()
                    out[k] = bufs[k]
                else:
                    out[k] = v
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._bufs[slot] = bufs
        self._out = getattr(self, "_out", [None, None])
        self._out[slot] = out
        self._ready[slot] = ev
        return True

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        if not self._pending:
            raise StopIteration
        slot = self._i % 2
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self._ready[slot])
        batch = self._out[slot]
        # everything the consumer enqueued for batch i-1 (the other slot) precedes this point on `cur`
        other = 1 - slot
        ev = torch.cuda.Event()
        ev.record(cur)
        self._free[other] = ev
        self._i += 1
        self._pending = self._issue(other)     # batch i+1 starts moving while batch i is computed
        return batch


class ScalarReader:
    """Device scalars -> host through a small pinned ring, read one step late (no launch-queue stall)."""

    _pinned = {}     # depth -> pinned ring, allocated once per process (cudaHostAlloc costs up to a millisecond and may synchronise)

    def __init__(self, device: torch.device, depth: int = 2):
        if depth not in ScalarReader._pinned:
            ScalarReader._pinned[depth] = torch.empty(depth, dtype=torch.float32).pin_memory()
        self.buf = ScalarReader._pinned[depth]
        self.events: List[Optional[torch.cuda.Event]] = [None] * depth
        self.depth, self.n = depth, 0
        self.values: List[float] = []
        self.bytes_read = 0

    def push(self, scalar: torch.Tensor):
        slot = self.n % self.depth
        if self.events[slot] is not None:      # the value written `depth` steps ago
            self.events[slot].synchronize()
            self.values.append(float(self.buf[slot]))
        self.buf[slot:slot + 1].copy_(scalar.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[slot] = ev
        self.bytes_read += 4
        self.n += 1

    def drain(self) -> List[float]:
        for i in range(max(0, self.n - self.depth), self.n):
            slot = i % self.depth
            if self.events[slot] is not None:
                self.events[slot].synchronize()
                self.values.append(float(self.buf[slot]))
                self.events[slot] = None
        return self.values


class InputStage:
    """GPU-side input stage: raw batch -> the sample dict the reference's model and loss read.

    The reference builds every training sample on the host (``bifold/data/processor.py``): frames go through the SigLIP image
    processor (``_process_rgb``, :143-148) and every pick/place pixel becomes an H x W fp32 Gaussian map (``_make_gaussmap``,
    :405-439).  Shipped over PCIe that is ~103 MB per 64-sample bimanual batch, half of it Gaussian maps that 8 integers per
    sample describe.  Here the host sends the raw form and two kernels rebuild the reference's tensors on the device:

        raw["rgb_u8"]            uint8 [B, H, W, 3]      frame at model resolution          -> sample["rgb"]  fp32 [B, 3, H, W]
        raw["rgb_context_u8"]    uint8 [B, T, H, W, 3]   (sequential models)                -> sample["rgb_context"]
        raw["centers"]           int32 [heads, B, 2]     rounded [x, y] per head, x = NO_ACTION for a missing arm
                                 or    [heads, B, P, 2]  several candidate pixels ("gmm")   -> sample["<head>_heatmap"] fp32 [B, H, W]
        raw["mask_u8"]           uint8 [B, H, W]         cloth mask (get_action only)       -> sample["mask"] fp32 [B, 1, H, W]
        everything else (instruction ids, context_attention_mask, ...) passes through.

    Capturable: no allocation-dependent host logic, so ``GraphedTrainStep(..., input_stage=stage)`` records it in the step graph.
    """

    def __init__(self, heads_out, image_size: int, sigma: float = 5.0, mean: float = 0.5, std: float = 0.5):
        from . import ops  # noqa: F401  (registers torch.ops.bfk.*)
        self.heads_out, self.image_size = list(heads_out), int(image_size)
        self.sigma, self.mean, self.std = float(sigma), float(mean), float(std)

    def __call__(self, raw: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        H = self.image_size
        s = {k: v for k, v in raw.items() if k not in ("rgb_u8", "rgb_context_u8", "centers", "mask_u8")}
        s["rgb"] = torch.ops.bfk.normalize_u8(raw["rgb_u8"], self.mean, self.std)
        if "rgb_context_u8" in raw:
            s["rgb_context"] = torch.ops.bfk.normalize_u8(raw["rgb_context_u8"], self.mean, self.std)
        if "centers" in raw:
            c = raw["centers"]
            maps = torch.ops.bfk.gauss_targets(c, H, H, self.sigma, c.dim() == 4)
            for i, h in enumerate(self.heads_out):
                s[f"{h}_heatmap"] = maps[i]
        if "mask_u8" in raw:
            s["mask"] = raw["mask_u8"].unsqueeze(1).to(torch.float32)
        return s
