"""Generate tests/golden/input_stage.npz with the REAL reference's Processor._make_gaussmap (bifold/data/processor.py:405-439).

TEST INFRASTRUCTURE ONLY; runs in the build container only (needs /root/reference).  ``ftfy`` and ``open3d`` (imported by
modules next to the processor, unused by the two functions) are stubbed.

    python -m oracle.make_golden_inputs
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "input_stage.npz")


class _Stub(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(self.__name__ + "." + k)


def cases():
    """(name, image_size, sigma, strategy, points or None) — also imported by the tests."""
    rng = np.random.default_rng(7)
    out = [("none64", 64, 5.0, "gmm", None),
           ("first64", 64, 5.0, "gmm", np.array([[10.4, 50.6]], dtype=np.float32)),
           ("half_even", 64, 5.0, "gmm", np.array([[2.5, 3.5]], dtype=np.float32)),        # round-half-even: (2, 4)
           ("corner224", 224, 5.0, "gmm", np.array([[0.0, 223.0]], dtype=np.float32)),
           ("gmm3_64", 64, 3.0, "gmm", np.array([[10.2, 12.7], [40.5, 41.5], [41.0, 44.0]], dtype=np.float32)),
           ("first_of_many", 64, 5.0, "first", np.array([[30.0, 31.0], [5.0, 6.0]], dtype=np.float32))]
    for i in range(3):
        out.append((f"rand224_{i}", 224, 5.0, "gmm", rng.uniform(0, 223, size=(1, 2)).astype(np.float32)))
    out.append(("gmm4_224", 224, 5.0, "gmm", rng.uniform(0, 223, size=(4, 2)).astype(np.float32)))
    return out


def main():
    ref_shims.import_reference()
    for n in ("ftfy", "open3d"):
        if n not in sys.modules:
            m = _Stub(n)
            m.__path__ = []
            sys.modules[n] = m
    from bifold.data.processor import Processor
    store = {}
    for name, size, sigma, strategy, pts in cases():
        self = types.SimpleNamespace(image_size=size, cfg=types.SimpleNamespace(sigma=sigma, strategy=strategy))
        g = Processor._make_gaussmap(self, None if pts is None else torch.from_numpy(pts))
        assert g.dtype == torch.float32 and tuple(g.shape) == (size, size)
        store[name] = g.numpy()
    np.savez_compressed(OUT, **store)
    print("wrote", OUT, {k: v.shape for k, v in store.items()})


if __name__ == "__main__":
    main()
