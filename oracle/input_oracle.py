"""CPU restatement (numpy) of the two Processor steps the GPU input stage replaces.

TEST INFRASTRUCTURE ONLY: imported by tests/ (and nothing in bifold_b200/).  Pinned by tests/golden/input_stage.npz, which
oracle/make_golden_inputs.py generates by calling the REAL reference's ``Processor._make_gaussmap``.

* ``make_gaussmap``  <- bifold/data/processor.py:405-425 (strategies "first" and "gmm"; "fit" needs scipy and no shipped
                        config uses it with a single point, see processor.py:411)
* ``process_rgb_u8`` <- bifold/data/processor.py:143-148 through the SigLIP image processor for frames already at model
                        resolution: rescale by 1/255, normalise with mean = std = 0.5, HWC -> CHW
"""
from __future__ import annotations

import numpy as np


def make_gaussmap(points, image_size: int, sigma: float, strategy: str = "gmm") -> np.ndarray:
    """points: None (no action -> zeros) or array [P, 2] of [x, y]; returns float32 [H, W] (processor.py:405-439)."""
    xy = np.arange(image_size)
    x, y = np.meshgrid(xy, xy)                                        # :406-407
    if points is None:
        g = np.zeros_like(x, dtype=np.float64)                        # :409
    else:
        pts = np.asarray(points)
        strat = strategy if len(pts) > 1 else "first"                 # :412
        if strat == "first":
            cx, cy = round(pts[0, 0]), round(pts[0, 1])               # :414-415 (Python round: half to even)
            g = np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sigma * sigma))   # :416-417
        elif strat == "gmm":
            g = np.zeros_like(x, dtype=np.float64)
            for cx, cy in np.round(pts):                              # :420
                g += np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2 * sigma ** 2)) / ((2 * np.pi) * sigma ** 2)   # :421-424
            assert g.max() != 0
            g /= g.max()                                              # :426
        else:
            raise ValueError(strat)
    return g.astype(np.float32)                                       # torch.FloatTensor(gauss_map), :439


def process_rgb_u8(frames: np.ndarray, mean: float = 0.5, std: float = 0.5) -> np.ndarray:
    """uint8 [..., H, W, 3] -> float32 [..., 3, H, W]."""
    v = frames.astype(np.float32) * np.float32(1.0 / 255.0)
    v = (v - np.float32(mean)) / np.float32(std)
    return np.moveaxis(v, -1, -3).copy()
