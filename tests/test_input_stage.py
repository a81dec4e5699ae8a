"""GPU input stage.  CPU: the numpy oracle against golden maps produced by the REAL reference's
Processor._make_gaussmap.  GPU: bfk::gauss_targets / bfk::normalize_u8 / data.InputStage against the golden maps and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle.input_oracle import make_gaussmap, process_rgb_u8
from oracle.make_golden_inputs import cases

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "input_stage.npz"))


def test_oracle_gaussmap_matches_reference_golden():
    for name, size, sigma, strategy, pts in cases():
        got = make_gaussmap(pts, size, sigma, strategy)
        assert got.dtype == np.float32
        np.testing.assert_array_equal(got, GOLDEN[name], err_msg=name)      # bit-exact: same float64 arithmetic


def test_oracle_rgb_matches_siglip_processor():
    """(x/255 - 0.5)/0.5 is what the HF SigLIP image processor computes for a frame already at model resolution."""
    from transformers import SiglipImageProcessor
    rng = np.random.default_rng(0)
    frame = rng.integers(0, 256, size=(32, 32, 3), dtype=np.uint8)
    proc = SiglipImageProcessor(size={"height": 32, "width": 32})
    ref = proc(images=frame, return_tensors="pt")["pixel_values"][0].numpy()
    got = process_rgb_u8(frame)
    assert got.shape == (3, 32, 32)
    np.testing.assert_allclose(got, ref, atol=1e-6, rtol=0)


def _centers(pts):
    from bifold_b200.ops import NO_ACTION
    if pts is None:
        return [[NO_ACTION, NO_ACTION]]
    return [[int(np.round(p[0])), int(np.round(p[1]))] for p in pts]          # np.round == Python round: half to even


@pytest.mark.gpu
def test_gauss_targets_kernel_matches_reference_golden():
    from bifold_b200 import ops
    for name, size, sigma, strategy, pts in cases():
        c = _centers(pts)
        if strategy == "first":
            c = c[:1]
        ct = torch.tensor(c, dtype=torch.int32, device="cuda")
        if len(c) == 1:
            got = torch.ops.bfk.gauss_targets(ct[0][None], size, size, sigma)[0]
        else:
            got = torch.ops.bfk.gauss_targets(ct[None], size, size, sigma, True)[0]
        ref = torch.from_numpy(GOLDEN[name])
        err = (got.cpu() - ref).abs().max().item()
        assert err <= 1e-6, (name, err)
        # fp64 evaluation rounded once: at most 1 ulp away, and exactly equal nearly everywhere
        assert (got.cpu() != ref).float().mean().item() < 1e-3, name


@pytest.mark.gpu
def test_gauss_targets_batched_padding_and_missing_arm():
    from bifold_b200.ops import NO_ACTION
    H = 224
    c = torch.tensor([[[5, 7], [NO_ACTION, NO_ACTION]], [[100, 50], [30, 31]], [[NO_ACTION, 0], [NO_ACTION, 0]]], dtype=torch.int32, device="cuda")
    maps = torch.ops.bfk.gauss_targets(c, H, H, 5.0, True).cpu().numpy()
    np.testing.assert_allclose(maps[0], make_gaussmap(np.array([[5, 7]]), H, 5.0), atol=1e-7)        # padded second point: "first"
    np.testing.assert_allclose(maps[1], make_gaussmap(np.array([[100, 50], [30, 31]]), H, 5.0), atol=1e-7)
    assert (maps[2] == 0).all()
    assert maps[0].max() == 1.0 and maps[1].max() == 1.0


@pytest.mark.gpu
def test_normalize_u8_and_input_stage():
    from bifold_b200.data import InputStage
    from bifold_b200.ops import NO_ACTION
    rng = np.random.default_rng(1)
    B, T, H = 3, 2, 32
    heads = ["left_pick", "left_place", "right_pick", "right_place"]
    raw = {"rgb_u8": torch.from_numpy(rng.integers(0, 256, size=(B, H, H, 3), dtype=np.uint8)),
           "rgb_context_u8": torch.from_numpy(rng.integers(0, 256, size=(B, T, H, H, 3), dtype=np.uint8)),
           "instruction": torch.randint(0, 100, (B, 64)),
           "context_attention_mask": torch.ones(B, T, dtype=torch.int64),
           "centers": torch.from_numpy(rng.integers(0, H, size=(4, B, 2)).astype(np.int32)),
           "mask_u8": torch.from_numpy(rng.integers(0, 2, size=(B, H, H), dtype=np.uint8))}
    raw["centers"][2, 1, 0] = NO_ACTION
    stage = InputStage(heads, H, sigma=5.0)
    s = stage({k: v.cuda() for k, v in raw.items()})
    np.testing.assert_allclose(s["rgb"].cpu().numpy(), process_rgb_u8(raw["rgb_u8"].numpy()), atol=1e-6, rtol=0)
    np.testing.assert_allclose(s["rgb_context"].cpu().numpy(), process_rgb_u8(raw["rgb_context_u8"].numpy()), atol=1e-6, rtol=0)
    assert s["rgb"].shape == (B, 3, H, H) and s["rgb_context"].shape == (B, T, 3, H, H)
    for i, h in enumerate(heads):
        for b in range(B):
            c = raw["centers"][i, b].numpy()
            ref = make_gaussmap(None if c[0] == NO_ACTION else c[None].astype(np.float32), H, 5.0)
            np.testing.assert_allclose(s[f"{h}_heatmap"][b].cpu().numpy(), ref, atol=1e-7, rtol=0)
    assert s["mask"].shape == (B, 1, H, H) and s["mask"].dtype == torch.float32
    assert torch.equal(s["instruction"].cpu(), raw["instruction"])
