"""Generate tests/golden/*.npz by running the REAL reference (/root/reference/bifold).

TEST INFRASTRUCTURE ONLY.  Runs in the build container only (needs /root/reference; the GPU box
never runs this).  For every case: seeded weights (oracle.synth.make_state_dict, loaded with
strict=True so the key surface is checked too) and a seeded sample go through the reference's own
``Models`` / ``Losses`` / ``Optimizers`` factories: forward, BCE loss, backward, one Adam step,
second forward, ``get_action``.  What is stored is small: full tensors for the tiny cases, strided
probes + norms for the base-size cases.

    python -m oracle.make_golden            # regenerates every fixture
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shims  # noqa: E402
from oracle.synth import PolicySpec, base224_spec, make_sample, make_state_dict, tiny_spec, trainable_keys  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (spec, batch, full tensors?)
    "tiny_uni": (tiny_spec(name="siglip", is_bimanual=False, lora=False), 3, True),
    "tiny_bi_lora": (tiny_spec(name="siglip", is_bimanual=True, lora=True), 4, True),
    "tiny_bi_seq_lora": (tiny_spec(name="siglip_sequential", is_bimanual=True, lora=True, context_length=3), 5, True),
    # CrossAttention fusion (bifold/models/fusion.py:67-125): single frame, and with context frames (float attn_mask quirk)
    "tiny_bi_lora_xattn": (tiny_spec(name="siglip", is_bimanual=True, lora=True, fusion_model="crossattention"), 4, True),
    "tiny_uni_seq_xattn": (tiny_spec(name="siglip_sequential", is_bimanual=False, lora=False, context_length=2,
                                     fusion_model="crossattention"), 5, True),
    "base224_bi_lora": (base224_spec(name="siglip", is_bimanual=True, lora=True), 2, False),
    "base224_bi_seq_lora": (base224_spec(name="siglip_sequential", is_bimanual=True, lora=True, context_length=3), 2, False),
}
# compute_mask=True cases (mask head + gated pick maps + the mask losses of bifold/conf/loss/composed*.yaml):
# name: (spec, batch, composed loss names, weights)
MASK_CASES = {
    "tiny_bi_seq_lora_mask": (tiny_spec(name="siglip_sequential", is_bimanual=True, lora=True, context_length=2, compute_mask=True),
                              4, ["bce_gaussmap", "focal", "dice"], [40, 20, 1]),
    "tiny_uni_mask_detach": (tiny_spec(name="siglip", is_bimanual=False, lora=False, compute_mask=True, detach_mask=True),
                             3, ["bce_gaussmap", "bce_mask"], [2, 1]),
}
PROBE = 97  # prime stride for probes of big tensors


def probe(t: torch.Tensor, full: bool) -> np.ndarray:
    a = t.detach().cpu().numpy()
    return a if full else a.reshape(-1)[::PROBE][:4096].copy()


def run_case(name: str, spec: PolicySpec, batch: int, full: bool):
    _, Losses, Optimizers = ref_shims.import_reference()
    sd = make_state_dict(spec, seed=3407)
    sample = make_sample(spec, batch, seed=0)
    model = ref_shims.build_reference_model(spec, sd)
    model.train()
    assert sorted(k for k, p in model.named_parameters() if p.requires_grad) == sorted(trainable_keys(spec, sd))

    logits = {}
    hooks = []
    for h in spec.heads_out:
        dec = getattr(model.pick_place, f"{h}_decoder")
        hooks.append(dec.register_forward_hook(lambda m, i, o, h=h: logits.__setitem__(h, o.squeeze(1).detach().clone())))

    loss_fn = Losses.get_by_name(cfg=ref_shims.Cfg(name="bce_gaussmap", is_bimanual=spec.is_bimanual,
                                                   mask_pick_heatmap=False))
    optim = Optimizers.get_by_name(cfg=ref_shims.Cfg(name="adam", lr=1e-3, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                   params=filter(lambda p: p.requires_grad, model.parameters()))
    out = model(dict(sample))
    loss, inter = loss_fn(out, sample)
    optim.zero_grad()
    loss.backward()
    g = {}
    g["loss"] = loss.detach().numpy()
    for h in spec.heads_out:
        g[f"loss/{h}"] = inter[h].detach().numpy()
        g[f"logits/{h}"] = probe(logits[h], full)
        g[f"heatmap/{h}"] = probe(out[f"{h}_heatmap"], full)
        g[f"logits_absmax/{h}"] = logits[h].abs().max().numpy()
    if out.get("attn_weights") is not None:       # CrossAttention: head-averaged attention weights (bifold/models/fusion.py:117-125)
        g["attn_weights"] = probe(out["attn_weights"], full)
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, k
            g[f"grad/{k}"] = probe(p.grad, full)
            g[f"gradnorm/{k}"] = p.grad.double().norm().numpy()
    optim.step()
    for k, p in model.named_parameters():
        if p.requires_grad:
            g[f"param1/{k}"] = probe(p, full)
    model.eval()
    action, out2 = model.get_action(dict(sample), return_raw_output=True)
    loss2, _ = loss_fn(out2, sample)
    g["loss_after_step"] = loss2.detach().numpy()
    for k, v in vars(action).items():
        g[f"action/{k}"] = np.asarray(v).astype(np.int64)
    # mask_pick_heatmap=True variant of the loss (bifold/losses/__init__.py:76-77)
    loss_fn_m = Losses.get_by_name(cfg=ref_shims.Cfg(name="bce_gaussmap", is_bimanual=spec.is_bimanual,
                                                     mask_pick_heatmap=True))
    g["loss_after_step_masked"] = loss_fn_m(out2, sample)[0].detach().numpy()
    for h in hooks:
        h.remove()
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **g)
    print(f"{name}: loss={float(loss):.6f} loss_after={float(loss2):.6f} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def run_mask_case(name: str, spec: PolicySpec, batch: int, loss_names, weights):
    """compute_mask=True model + ComposedLoss through the reference's own factories (bifold/losses/__init__.py:30-51,112-155)."""
    _, Losses, Optimizers = ref_shims.import_reference()
    sd = make_state_dict(spec, seed=3407)
    sample = make_sample(spec, batch, seed=0)
    model = ref_shims.build_reference_model(spec, sd)
    model.train()
    assert sorted(k for k, p in model.named_parameters() if p.requires_grad) == sorted(trainable_keys(spec, sd))
    loss_fn = Losses.get_by_name(cfg=ref_shims.Cfg(name="composed", loss_names=list(loss_names), weights=list(weights),
                                                   is_bimanual=spec.is_bimanual, mask_pick_heatmap=True))
    optim = Optimizers.get_by_name(cfg=ref_shims.Cfg(name="adam", lr=1e-3, betas=[0.9, 0.999], eps=1e-8, weight_decay=0),
                                   params=filter(lambda p: p.requires_grad, model.parameters()))
    out = model(dict(sample))
    loss, inter = loss_fn(out, sample)
    optim.zero_grad()
    loss.backward()
    g = {"loss": loss.detach().numpy()}
    for k, v in inter.items():
        g[f"loss/{k}"] = v.detach().numpy()
    for k, v in out.items():
        if v is not None:
            g[f"heatmap/{k}"] = v.detach().numpy()
    for k, p in model.named_parameters():
        if p.requires_grad:
            assert p.grad is not None, k
            g[f"grad/{k}"] = p.grad.detach().numpy()
            g[f"gradnorm/{k}"] = p.grad.double().norm().numpy()
    optim.step()
    for k, p in model.named_parameters():
        if p.requires_grad:
            g[f"param1/{k}"] = p.detach().numpy()
    model.eval()
    action, out2 = model.get_action(dict(sample), return_raw_output=True)
    g["loss_after_step"] = loss_fn(out2, sample)[0].detach().numpy()
    for k, v in vars(action).items():
        g[f"action/{k}"] = np.asarray(v).astype(np.int64)
    path = os.path.join(GOLDEN_DIR, f"{name}.npz")
    np.savez_compressed(path, **g)
    print(f"{name}: loss={float(loss):.6f} -> {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def main(argv):
    names = argv or (list(CASES) + list(MASK_CASES))
    torch.set_num_threads(os.cpu_count() or 1)
    for n in names:
        if n in MASK_CASES:
            run_mask_case(n, *MASK_CASES[n])
        else:
            spec, batch, full = CASES[n]
            run_case(n, spec, batch, full)


if __name__ == "__main__":
    main(sys.argv[1:])
