"""The CPU restatement (oracle/policy_oracle.py) against fixtures produced by the REAL reference
(oracle/make_golden.py).  CPU only; this is what pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import policy_oracle as po
from oracle.make_golden import CASES, PROBE
from oracle.synth import make_sample, make_state_dict, trainable_keys

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _probe(t, full):
    a = t.detach().cpu().numpy()
    return a if full else a.reshape(-1)[::PROBE][:4096]


def _check_attn_weights(sd, sample, spec, g):
    """CrossAttention's head-averaged attention weights (the module's second return value) against the reference's."""
    b = sample["rgb"].shape[0]
    img = torch.cat((sd["image_token"].expand(b, 1, -1), po.vision_tower(sd, sample["rgb"], spec, True)), dim=1)
    txt = torch.cat((sd["text_token"].expand(b, 1, -1), po.text_tower(sd, sample["instruction"], spec, True)), dim=1)
    if spec.name == "siglip_sequential":
        c = sample["rgb_context"]
        t, n = c.shape[1], img.shape[1]
        ctx = po.vision_tower(sd, c.reshape(b * t, *c.shape[2:]), spec, True).view(b, t, n - 1, -1)
        ctx = torch.cat((sd["image_token"][:, None].expand(b, t, 1, -1), ctx), dim=2).reshape(b, t * n, -1) + sd["context_pos_embedding"]
        masks = torch.cat([torch.ones(b, txt.shape[1]), sample["context_attention_mask"].repeat_interleave(n, dim=1).float(),
                           torch.ones(b, n)], dim=-1)
        _, w = po.cross_attention(sd, (txt, ctx, img), [0, 1, 1], spec, masks, return_weights=True)
    else:
        _, w = po.cross_attention(sd, (txt, img), [0, 1], spec, None, return_weights=True)
    np.testing.assert_allclose(w.detach().numpy(), g["attn_weights"], atol=1e-6, rtol=0)


def _check_case(name):
    spec, batch, full = CASES[name]
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    sd = make_state_dict(spec, seed=3407)
    sample = make_sample(spec, batch, seed=0)
    keys = trainable_keys(spec, sd)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in keys}
    work = dict(sd)
    work.update(leaves)
    logits = po.policy_logits(work, sample, spec, training=True)
    if "attn_weights" in g.files:
        _check_attn_weights(work, sample, spec, g)
    out = {f"{h}_heatmap": z.sigmoid() for h, z in logits.items()}
    loss, inter = po.bce_gaussmap(out, sample, spec)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=2e-5)
    for h in spec.heads_out:
        np.testing.assert_allclose(inter[h].item(), g[f"loss/{h}"], rtol=2e-5)
        scale = float(g[f"logits_absmax/{h}"])
        np.testing.assert_allclose(_probe(logits[h], full), g[f"logits/{h}"], atol=2e-5 * scale, rtol=0)
        np.testing.assert_allclose(_probe(out[f"{h}_heatmap"], full), g[f"heatmap/{h}"], atol=1e-5, rtol=0)
    grads = dict(zip(keys, torch.autograd.grad(loss, [leaves[k] for k in keys])))
    for k in keys:
        gn = float(g[f"gradnorm/{k}"])
        np.testing.assert_allclose(grads[k].double().norm().item(), gn, rtol=1e-3, atol=1e-9, err_msg=k)
        np.testing.assert_allclose(_probe(grads[k], full), g[f"grad/{k}"], atol=2e-4 * gn / max(1.0, np.sqrt(grads[k].numel() / 64)) + 1e-9, rtol=1e-3, err_msg=k)
    # one Adam step, lr 1e-3 as in make_golden.py
    with torch.no_grad():
        params = {k: sd[k] for k in keys}
        po.adam_step(params, grads, {}, lr=1e-3)
    for k in keys:
        a, b = _probe(sd[k], full), g[f"param1/{k}"]
        if k.endswith("cross_attention.in_proj_bias"):
            # the key bias shifts every score of a query row by the same amount: its exact gradient is ZERO, what autograd
            # returns is rounding noise (|g| ~ 1e-10) that Adam normalises to +-lr -> not comparable, in any implementation
            d = a.shape[0] // 3
            a, b = np.concatenate([a[:d], a[2 * d:]]), np.concatenate([b[:d], b[2 * d:]])
        np.testing.assert_allclose(a, b, atol=2e-5, rtol=1e-5, err_msg=k)
    with torch.no_grad():
        out2 = po.policy_forward(sd, sample, spec)
        loss2, _ = po.bce_gaussmap(out2, sample, spec)
        loss2m, _ = po.bce_gaussmap(out2, sample, spec, mask_pick_heatmap=True)
    np.testing.assert_allclose(loss2.item(), g["loss_after_step"], rtol=1e-3)
    np.testing.assert_allclose(loss2m.item(), g["loss_after_step_masked"], rtol=1e-3)
    action = po.get_action(out2, sample, spec)
    for k, v in action.items():
        np.testing.assert_array_equal(np.asarray(v).astype(np.int64), g[f"action/{k}"], err_msg=k)


@pytest.mark.parametrize("name", ["tiny_uni", "tiny_bi_lora", "tiny_bi_seq_lora"])
def test_oracle_matches_reference_tiny(name):
    _check_case(name)


@pytest.mark.parametrize("name", ["tiny_bi_lora_xattn", "tiny_uni_seq_xattn"])
def test_oracle_matches_reference_cross_attention(name):
    """fusion_model=crossattention (bifold/models/fusion.py:67-125): one MultiheadAttention layer, image tokens query the text
    (+ context) tokens; with context frames the float attn_mask is ADDED to the scores with its batch axis tiled (the quirk the
    restatement reproduces)."""
    _check_case(name)


@pytest.mark.parametrize("name", ["base224_bi_lora", "base224_bi_seq_lora"])
def test_oracle_matches_reference_base(name):
    if not os.path.exists(os.path.join(GOLDEN, f"{name}.npz")):
        pytest.skip("fixture not generated")
    _check_case(name)


@pytest.mark.parametrize("name", ["tiny_bi_seq_lora_mask", "tiny_uni_mask_detach"])
def test_oracle_matches_reference_mask_branch(name):
    """compute_mask=True (mask head on the fusion input, gated pick maps, detach_mask) and the composed mask losses
    (bce_mask / focal / dice) against fixtures from the real reference."""
    from oracle.make_golden import MASK_CASES
    spec, batch, loss_names, weights = MASK_CASES[name]
    g = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    sd = make_state_dict(spec, seed=3407)
    sample = make_sample(spec, batch, seed=0)
    keys = trainable_keys(spec, sd)
    assert any(".mask_head." in k for k in keys)
    leaves = {k: sd[k].clone().requires_grad_(True) for k in keys}
    work = dict(sd)
    work.update(leaves)
    out = po.policy_forward(work, sample, spec, training=True)
    loss, inter = po.composed_loss(out, sample, spec, loss_names, weights)
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=3e-5)
    for n in loss_names:
        np.testing.assert_allclose(inter[n].item(), g[f"loss/{n}"], rtol=3e-5)
    for k, v in out.items():
        if v is not None:
            np.testing.assert_allclose(v.detach().numpy(), g[f"heatmap/{k}"], atol=1e-5, rtol=0)
    grads = dict(zip(keys, torch.autograd.grad(loss, [leaves[k] for k in keys])))
    for k in keys:
        gn = float(g[f"gradnorm/{k}"])
        np.testing.assert_allclose(grads[k].double().norm().item(), gn, rtol=2e-3, atol=1e-9, err_msg=k)
        np.testing.assert_allclose(grads[k].numpy(), g[f"grad/{k}"], atol=3e-4 * gn / max(1.0, np.sqrt(grads[k].numel() / 64)) + 1e-9,
                                   rtol=2e-3, err_msg=k)
