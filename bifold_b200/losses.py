"""Host-side mirror of ``bifold.losses`` for the hot path (bifold/losses/__init__.py:5-155).

``BCEGaussMap`` computes nn.BCELoss per head (mean over B*H*W, log clamped at -100) summed over
heads, on our CUDA kernels:

* fused path  – when the model already computed the per-head BCE inside its heatmap kernel against
  the very same target tensors (``PolicyOutput.fused_loss``), the loss is just read back and the
  gradient goes into the tail backward kernel as (p - t)/(B*H*W);
* general path – ``torch.ops.bfk.bce_prob`` (C ABI ``bfk_bce_prob`` / ``bfk_bce_prob_bwd``) on the probabilities (any caller-supplied
  heatmaps, or ``mask_pick_heatmap=True``).
"""
from __future__ import annotations

import torch

from . import _cabi as K
from . import ops  # noqa: F401  (registers torch.ops.bfk.*)


class Losses:
    @staticmethod
    def get_by_name(cfg=None, name=None, *args, **kwargs):
        if cfg is not None:
            name = cfg.name
        if name == "bce_gaussmap":
            Loss = BCEGaussMap
        elif name == "composed":
            Loss = ComposedLoss
        elif name == "bce_mask":
            Loss = BCEMask
        elif name == "dice":
            Loss = DiceLoss
        elif name == "focal":
            Loss = FocalLoss
        else:
            raise ValueError(f"Loss {name} not recognized")
        if cfg is not None:
            del cfg.name
            return Loss(*args, **cfg, **kwargs)
        return Loss(*args, **kwargs)


class ComposedLoss:   # bifold/losses/__init__.py:30-51
    def __init__(self, loss_names, weights, **kwargs):
        assert len(loss_names) == len(weights)
        self.loss_dict = {n: Losses.get_by_name(name=n, **kwargs) for n in loss_names}
        self.weights = dict(zip(loss_names, weights))

    def __call__(self, output, sample):
        loss, inter = None, {}
        for n, fn in self.loss_dict.items():
            cur, cur_inter = fn(output, sample)
            loss = cur * self.weights[n] if loss is None else loss + cur * self.weights[n]
            inter[n] = cur
            for k, v in cur_inter.items():
                inter[n + " " + k] = v
        return loss, inter


class BCEGaussMap:    # bifold/losses/__init__.py:54-109
    def __init__(self, is_bimanual, mask_pick_heatmap, **kwargs):
        self.is_bimanual = is_bimanual
        self.mask_pick_heatmap = mask_pick_heatmap

    def _heads(self):
        if self.is_bimanual:
            return [f"{arm}_{action}" for arm in ("left", "right") for action in ("pick", "place")]
        return ["pick", "place"]

    def __call__(self, output, sample):
        heads = self._heads()
        fused = getattr(output, "fused_loss", None)
        if fused is not None and not self.mask_pick_heatmap and len(heads) + 1 == fused.numel():
            ptrs = tuple(sample[f"{h}_heatmap"].data_ptr() for h in heads)
            if ptrs == output.fused_targets:
                # fused[:heads] = per-head BCE, fused[heads] = their sum (computed in the finalize kernel)
                return fused[len(heads)], {h: fused[i] for i, h in enumerate(heads)}
        inter, loss = {}, None
        for h in heads:
            p = output[f"{h}_heatmap"]
            if not p.is_cuda:
                raise K.BfkError("bifold_b200 losses run on CUDA only (no CPU fallback)")
            t = sample[f"{h}_heatmap"].to(p.device)
            tm = sample["mask"].squeeze(1).to(p.device) if (self.mask_pick_heatmap and h.endswith("pick")) else None
            cur = torch.ops.bfk.bce_prob(p, t, tm)
            inter[h] = cur
            loss = cur if loss is None else loss + cur
        return loss, inter


def _mask_inputs(output, sample):
    p = output["mask_heatmap"]            # needs a compute_mask=True model, like the reference (KeyError otherwise)
    if not p.is_cuda:
        raise K.BfkError("bifold_b200 losses run on CUDA only (no CPU fallback)")
    return p, sample["mask"].to(p.device)


class BCEMask:        # bifold/losses/__init__.py:112-117
    def __init__(self, **kwargs):
        pass

    def __call__(self, output, sample):
        p, m = _mask_inputs(output, sample)
        return torch.ops.bfk.bce_prob(p, m.squeeze(1), None), {}


class DiceLoss:       # bifold/losses/__init__.py:126-136
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, output, sample):
        p, m = _mask_inputs(output, sample)
        return torch.ops.bfk.dice_loss(p.flatten(1), m.flatten(1))[0], {}


class FocalLoss:      # bifold/losses/__init__.py:139-155
    def __init__(self, alpha=0.25, gamma=2, **kwargs):
        self.alpha = alpha
        self.gamma = gamma

    def __call__(self, output, sample):
        p, m = _mask_inputs(output, sample)
        return torch.ops.bfk.focal_loss(p, m.squeeze(1), float(self.alpha), float(self.gamma)), {}
