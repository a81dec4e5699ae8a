"""Checkpoint compatibility: the reference's checkpoint layout (bifold/__main__.py:304-325) with a torch.optim.Adam
optimizer state loads into the drop-in model + FusedAdam and back, and training continues bit-identically."""
import io
import random

import numpy as np
import pytest
import torch

from oracle.synth import make_sample, tiny_spec
from tests.util import Cfg, build_model, to_device

pytestmark = pytest.mark.gpu


def _setup(spec, sd=None, lr=1e-3):
    from bifold_b200.losses import Losses
    from bifold_b200.optim import Optimizers
    model, sd = build_model(spec, "bf16", state_dict=sd)
    model.train()
    loss_fn = Losses.get_by_name(cfg=Cfg(name="bce_gaussmap", is_bimanual=spec.is_bimanual, mask_pick_heatmap=False))
    optim = Optimizers.get_by_name(cfg=Cfg(name="adam", lr=lr, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                   params=filter(lambda p: p.requires_grad, model.parameters()))
    return model, loss_fn, optim


def _step(model, loss_fn, optim, sample):
    loss, _ = loss_fn(model(sample), sample)
    optim.zero_grad()
    loss.backward()
    optim.step()
    return loss.item()


def _save_like_reference(model, optimizer, epoch):
    """bifold/__main__.py:304-325 (save_model)."""
    state = {"epoch": epoch + 1,
             "random_states": (random.getstate(), np.random.get_state(), torch.get_rng_state(), torch.cuda.get_rng_state()),
             "model": model.state_dict(), "optimizer": optimizer.state_dict()}
    buf = io.BytesIO()
    torch.save(state, buf)
    buf.seek(0)
    return buf


def test_reference_checkpoint_layout_round_trip():
    spec = tiny_spec(name="siglip_sequential", is_bimanual=True, lora=True, context_length=2)
    samples = [to_device(make_sample(spec, 3, seed=40 + i)) for i in range(4)]
    a_model, loss_fn, a_opt = _setup(spec)
    for i in range(2):
        _step(a_model, loss_fn, a_opt, samples[i])
    ckpt = torch.load(_save_like_reference(a_model, a_opt, epoch=0), map_location="cuda", weights_only=False)
    assert set(ckpt) >= {"epoch", "random_states", "model", "optimizer"}

    # (1) the optimizer state is a valid torch.optim.Adam state: load it into the stock optimizer over same-shaped parameters ...
    b_model, _, _ = _setup(spec)
    b_model.load_state_dict(ckpt["model"])                                   # strict: identical key surface
    b_params = [p for p in b_model.parameters() if p.requires_grad]
    torch_adam = torch.optim.Adam(b_params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    torch_adam.load_state_dict(ckpt["optimizer"])
    st = torch_adam.state[b_params[0]]
    assert float(st["step"]) == 2.0 and st["exp_avg"].shape == b_params[0].shape
    # ... (2) and what torch.optim.Adam saves loads back into FusedAdam (load_model, bifold/__main__.py:272-302)
    c_model, _, c_opt = _setup(spec)
    c_model.load_state_dict(ckpt["model"])
    c_opt.load_state_dict(torch_adam.state_dict())

    # (3) training continues bit-identically from the restored state
    for i in (2, 3):
        la = _step(a_model, loss_fn, a_opt, samples[i])
        lc = _step(c_model, loss_fn, c_opt, samples[i])
        assert la == lc, (i, la, lc)
    torch.cuda.synchronize()
    pa, pc = dict(a_model.named_parameters()), dict(c_model.named_parameters())
    for n in pa:
        assert torch.equal(pa[n], pc[n]), n
    sa, sc = a_opt.state_dict()["state"], c_opt.state_dict()["state"]
    for k in sa:
        assert float(sa[k]["step"]) == float(sc[k]["step"]) == 4.0
        assert torch.equal(sa[k]["exp_avg_sq"], sc[k]["exp_avg_sq"])


def test_torch_adam_trajectory_matches_fused_adam_from_checkpoint():
    """The stock optimizer and FusedAdam produce the same next step from the same restored state (fp32 tolerance)."""
    spec = tiny_spec(name="siglip", is_bimanual=True, lora=True)
    sample = to_device(make_sample(spec, 3, seed=5))
    model, loss_fn, opt = _setup(spec)
    _step(model, loss_fn, opt, sample)
    ckpt = torch.load(_save_like_reference(model, opt, 0), map_location="cuda", weights_only=False)
    m2, _, _ = _setup(spec)
    m2.load_state_dict(ckpt["model"])
    p2 = [p for p in m2.parameters() if p.requires_grad]
    stock = torch.optim.Adam(p2, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0)
    stock.load_state_dict(ckpt["optimizer"])
    _step(model, loss_fn, opt, sample)
    _step(m2, loss_fn, stock, sample)
    torch.cuda.synchronize()
    a, b = dict(model.named_parameters()), dict(m2.named_parameters())
    for n, p in a.items():
        if p.requires_grad:
            assert torch.allclose(p, b[n], rtol=1e-5, atol=2e-7), n
