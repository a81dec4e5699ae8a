"""Host-side mirror of ``bifold.optim`` (bifold/optim/__init__.py:4-92): Adam / AdamW with torch
semantics executed by one fused CUDA kernel over flat buffers (bfk_adamw).

The reference's learning-rate scheduler (``bifold.optim.Schedulers`` / ``LinearWarmup``, bifold/optim/__init__.py:17-92) is host
Python that only rewrites ``optimizer.param_groups[i]["lr"]``; it is NOT re-implemented here — the reference's own class drives
``FusedAdam`` unchanged, eagerly and under CUDA-graph replay (``sync_lr`` pushes the new value to the device)."""
from __future__ import annotations

import ctypes as C

import weakref

import torch

from . import _cabi as K


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (L2 weight decay) / AdamW (decoupled) on the bfk_adamw kernel.

    Parameters are re-homed into one flat fp32 buffer per param group (the nn.Parameter objects keep
    their identity; ``p.data`` becomes a view), so a step is ONE launch when the gradients are views of
    one flat arena laid out the same way (that is how bifold_b200.engine hands them out) and one launch
    per tensor otherwise.  ``self.state`` is torch-compatible (per-parameter ``step``, ``exp_avg``,
    ``exp_avg_sq``), so ``state_dict()`` / ``load_state_dict()`` interoperate with torch.optim.Adam.
    """

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, **unused):
        defaults = dict(lr=float(lr), betas=tuple(float(b) for b in betas), eps=float(eps), weight_decay=float(weight_decay),
                        decoupled=bool(decoupled))
        super().__init__(params, defaults)
        self.grad_scale = 1.0          # DDP folds 1/world_size in here
        self.pre_step_hook = None      # DDP: wait for the gradient all-reduce
        self._flat = {}
        self.launches = 0
        self.device_schedule = False   # step count / lr / bias corrections live on the device (CUDA-graph replay)

    def _flatten(self, gi, group):
        ps = [p for p in group["params"] if p.requires_grad]
        if not ps:
            return None
        dev = ps[0].device
        if dev.type != "cuda":
            raise K.BfkError("bifold_b200 optimizers run on CUDA only (no CPU fallback)")
        offs, tot = [], 0
        for p in ps:
            offs.append(tot)
            tot += (p.numel() + 7) // 8 * 8
        flat = torch.zeros(tot, device=dev, dtype=torch.float32)
        m, v = torch.zeros_like(flat), torch.zeros_like(flat)
        step = 0
        for p, o in zip(ps, offs):
            flat[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = flat[o:o + p.numel()].view(p.shape)
            s = self.state[p]
            if "exp_avg" in s:      # state loaded from a checkpoint: keep it
                m[o:o + p.numel()].copy_(s["exp_avg"].reshape(-1))
                v[o:o + p.numel()].copy_(s["exp_avg_sq"].reshape(-1))
                step = max(step, int(s["step"]))
            s["step"] = torch.tensor(float(step))
            s["exp_avg"] = m[o:o + p.numel()].view(p.shape)
            s["exp_avg_sq"] = v[o:o + p.numel()].view(p.shape)
        st = dict(ps=ps, offs=offs, tot=tot, p=flat, m=m, v=v, step=step, shadow=None)
        self._flat[gi] = st
        for p, o in zip(ps, offs):          # lets bifold_b200.engine find the group's bf16 shadow (ensure_shadow)
            p._bfk_opt = (weakref.ref(self), gi, o)
            p._bfk_shadow_ver = None
        return st

    def ensure_shadow(self, gi):
        """bf16 copy of the flat parameter buffer of group ``gi``, same element offsets, kept current by the Adam kernel itself
        (2 extra bytes written per parameter instead of a separate cast pass over every trainable GEMM weight).  Created on the
        first request (the engine asks for it in bf16 mode); ``p._bfk_shadow_ver == p._version`` marks a parameter's slice as
        current."""
        st = self._flat[gi]
        if st["shadow"] is None:
            st["shadow"] = torch.empty(st["tot"], device=st["p"].device, dtype=torch.bfloat16)
            st["shadow"].copy_(st["p"])
            for p in st["ps"]:
                p._bfk_shadow_ver = p._version
        return st["shadow"]

    def _mark_shadow(self, st):
        if st.get("shadow") is not None:
            for p in st["ps"]:
                p._bfk_shadow_ver = p._version

    def _launch(self, n, p, g, m, v, group, step, dyn=None, shadow=None):
        b1, b2 = group["betas"]
        a = K.AdamWArgs(n, p, g, m, v, shadow, group["lr"], b1, b2, group["eps"], group["weight_decay"],
                        int(group["decoupled"]), step, float(self.grad_scale), K.ptr(dyn))
        K.check(K.load().bfk_adamw(C.byref(a), torch.cuda.current_stream().cuda_stream), "bfk_adamw")
        self.launches += 1

    # ---- device-resident schedule: what a captured CUDA graph needs (the kernel arguments of a graph are frozen) ----
    def enable_device_schedule(self):
        """Keep the step counter, learning rate and bias corrections of every flat group in device memory:
        ``bfk_adamw_advance`` (one thread) increments the counter and refreshes them right before ``bfk_adamw``,
        which reads them instead of its by-value arguments.  Works eagerly and inside a captured graph."""
        self.device_schedule = True

    def _dev_state(self, st, group):
        if "step_dev" not in st:
            dev = st["p"].device
            # created inside step(), after the host counter was incremented: the device counter holds the steps DONE
            st["step_dev"] = torch.full((1,), int(st["step"]) - 1, device=dev, dtype=torch.int32)
            st["lr_dev"] = torch.full((1,), float(group["lr"]), device=dev, dtype=torch.float32)
            st["lr_host"] = float(group["lr"])
            st["dyn"] = torch.zeros(3, device=dev, dtype=torch.float32)
        return st

    def sync_lr(self):
        """Push host-side learning-rate changes (schedulers mutate param_groups) to the device copies; call before a
        graph replay.  A no-op when nothing changed."""
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi)
            if st is not None and "lr_dev" in st and st["lr_host"] != float(group["lr"]):
                st["lr_dev"].fill_(float(group["lr"]))
                st["lr_host"] = float(group["lr"])

    def note_replayed_step(self):
        """Host bookkeeping for a step that ran as a graph replay (the device counter advanced by itself)."""
        for st in self._flat.values():
            st["step"] += 1
            for p in st["ps"]:
                self.state[p]["step"] += 1
            # the replayed kernels wrote the parameters through raw pointers: an eager forward after a replay must
            # re-derive its operand copies (Engine.prepare compares version counters)
            torch._C._increment_version(st["ps"])
            self._mark_shadow(st)           # the replayed Adam kernel wrote the bf16 shadow too

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        if self.pre_step_hook is not None:
            self.pre_step_hook()
        for gi, group in enumerate(self.param_groups):
            st = self._flat.get(gi)
            if st is None or any(p.data_ptr() != st["p"].data_ptr() + 4 * o for p, o in zip(st["ps"], st["offs"])) or \
                    any(self.state[p].get("exp_avg") is None or
                        self.state[p]["exp_avg"].data_ptr() != st["m"].data_ptr() + 4 * o for p, o in zip(st["ps"], st["offs"])):
                st = self._flatten(gi, group)
            if st is None:
                continue
            st["step"] += 1
            for p in st["ps"]:
                self.state[p]["step"] += 1
            ps, offs = st["ps"], st["offs"]
            g0 = ps[0].grad
            flat_ok = g0 is not None and g0.dtype == torch.float32 and all(
                p.grad is not None and p.grad.is_contiguous() and p.grad.data_ptr() == g0.data_ptr() + 4 * o
                for p, o in zip(ps, offs))
            if flat_ok and self.device_schedule:
                self._dev_state(st, group)
                self.sync_lr()
                b1, b2 = group["betas"]
                K.check(K.load().bfk_adamw_advance(st["step_dev"].data_ptr(), st["lr_dev"].data_ptr(), b1, b2, st["dyn"].data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream), "bfk_adamw_advance")
                self.launches += 1
                self._launch(st["tot"], st["p"].data_ptr(), g0.data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(), group, st["step"],
                             dyn=st["dyn"], shadow=K.ptr(st["shadow"]))
            elif flat_ok:
                self._launch(st["tot"], st["p"].data_ptr(), g0.data_ptr(), st["m"].data_ptr(), st["v"].data_ptr(), group, st["step"],
                             shadow=K.ptr(st["shadow"]))
            else:
                sh = st["shadow"].data_ptr() if st["shadow"] is not None else None
                for p, o in zip(ps, offs):
                    if p.grad is None:
                        continue
                    g = p.grad.contiguous().float()
                    if g.data_ptr() % 16:
                        g = g.clone()
                    self._launch(p.numel(), st["p"].data_ptr() + 4 * o, g.data_ptr(), st["m"].data_ptr() + 4 * o,
                                 st["v"].data_ptr() + 4 * o, group, st["step"], shadow=(sh + 2 * o) if sh is not None else None)
            # the kernels wrote through raw pointers: tell torch (and the engine's operand-copy cache) the values changed
            torch._C._increment_version(ps)
            self._mark_shadow(st)
        return loss


class Optimizers:
    @staticmethod
    def get_by_name(cfg, params):
        if cfg.name == "adam":
            del cfg.name
            return FusedAdam(params, decoupled=False, **cfg)
        if cfg.name == "adamw":
            del cfg.name
            return FusedAdam(params, decoupled=True, **cfg)
        raise ValueError(f"Optimizer {cfg.name} not recognized")
