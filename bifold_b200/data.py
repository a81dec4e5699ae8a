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

    ``batches`` yields dicts of (ideally pinned) host tensors.  Exactly one host->device copy is issued
    per yielded batch, one batch ahead of the consumer.
    """

    def __init__(self, batches: Iterable[Dict[str, torch.Tensor]], device: torch.device):
        self.it: Iterator[Dict[str, torch.Tensor]] = iter(batches)
        self.device = device
        self.copy_stream = torch.cuda.Stream(device=device)
        self.bytes_copied = 0
        self._next: Optional[Dict[str, torch.Tensor]] = None
        self._ready: Optional[torch.cuda.Event] = None
        self._issue()

    def _issue(self):
        try:
            host = next(self.it)
        except StopIteration:
            self._next, self._ready = None, None
            return
        with torch.cuda.stream(self.copy_stream):
            dev = {}
            for k, v in host.items():
                if torch.is_tensor(v):
                    dev[k] = v.to(self.device, non_blocking=True)
                    self.bytes_copied += v.numel() * v.element_size()
                else:
                    dev[k] = v
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self._next, self._ready = dev, ev

    def __iter__(self):
        return self

    def __next__(self) -> Dict[str, torch.Tensor]:
        if self._next is None:
            raise StopIteration
        batch, ev = self._next, self._ready
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for v in batch.values():
            if torch.is_tensor(v):
                v.record_stream(cur)      # allocated on the copy stream, consumed on the compute stream
        self._issue()                     # batch i+1 starts moving while batch i is computed
        return batch


class ScalarReader:
    """Device scalars -> host through a small pinned ring, read one step late (no launch-queue stall)."""

    def __init__(self, device: torch.device, depth: int = 2):
        self.buf = torch.empty(depth, dtype=torch.float32).pin_memory()
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
