"""CPU: the C-ABI library loads and exports every symbol include/bifold_b200.h declares; the ctypes
prototype table covers the same set.  No compute calls (no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "bifold_b200.h")).read()
    return sorted(set(re.findall(r"BFK_API\s+[\w\s\*]+?\s\**(bfk_\w+)\s*\(", text)))


def test_header_declares_entry_points():
    names = _declared()
    assert len(names) >= 30, names
    for must in ("bfk_gemm", "bfk_attention_fwd", "bfk_attention_bwd", "bfk_layernorm_fwd", "bfk_heatmap_fwd",
                 "bfk_heatmap_bwd", "bfk_adamw", "bfk_action", "bfk_argmax2d", "bfk_assemble_fwd"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from bifold_b200 import _cabi
    from bifold_b200.build import build
    build()
    lib = _cabi.load()
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_cabi._PROTOS) == _declared()
    assert lib.bfk_version() >= 100
    nb = lib.bfk_heatmap_bands(14)
    assert nb in (1, 2, 4, 8, 16)
    assert lib.bfk_heatmap_fwd_workspace_floats(2, 4, 14) == 3 * 2 * 4 * nb


def test_missing_library_fails_loudly(monkeypatch):
    from bifold_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libbifold_b200.so")
    with pytest.raises(_cabi.BfkError):
        _cabi.load()


def test_cpu_tensors_are_rejected():
    """No CPU fallback: the product path raises instead of computing on the host."""
    import torch
    from oracle.synth import make_sample, tiny_spec
    from tests.util import build_model
    from bifold_b200 import _cabi
    spec = tiny_spec(name="siglip", is_bimanual=False, lora=False)
    model, _ = build_model(spec, device="cpu")
    with pytest.raises(_cabi.BfkError):
        model(make_sample(spec, 1))


def test_upsample_taps_match_four_bilinear_stages():
    import torch
    from bifold_b200 import _cabi
    lib = _cabi.load()
    for g in (2, 4, 14, 24, 27):
        H = 16 * g
        ts = torch.empty(H, dtype=torch.int32)
        tw = torch.empty(H, 3)
        assert lib.bfk_upsample16_taps(g, ts.data_ptr(), tw.data_ptr()) == 0
        Cm = torch.zeros(H, g, dtype=torch.float64)
        for y in range(H):
            for k in range(3):
                if ts[y] + k < g:
                    Cm[y, ts[y] + k] = tw[y, k]
        x = torch.randn(1, 1, g, g, dtype=torch.float64)
        y = x
        for _ in range(4):   # bifold/models/modules.py:68-69
            y = torch.nn.functional.interpolate(y, scale_factor=2, mode="bilinear", align_corners=False)
        assert torch.allclose(Cm @ x[0, 0] @ Cm.T, y[0, 0], atol=1e-12)
        assert torch.allclose(Cm.sum(1), torch.ones(H, dtype=torch.float64))
