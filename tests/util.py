"""Shared helpers for the GPU parity tests (model construction from an oracle PolicySpec)."""
import numpy as np
import torch

from oracle.synth import PolicySpec, make_sample, make_state_dict


from bifold_b200.config import Cfg  # noqa: E402,F401


def model_cfg(spec: PolicySpec, precision="fp32") -> Cfg:
    cfg = Cfg(name=spec.name, image_size=spec.image_size, is_bimanual=spec.is_bimanual, requires_graph=False,
              patch_size=spec.patch_size, automodel_name="synthetic", dim=spec.dim, emb_dropout=0.0, lora=spec.lora,
              r=spec.r, lora_alpha=spec.lora_alpha, lora_dropout=spec.lora_dropout, target_modules=["q_proj", "v_proj"],
              threshold=0.5, text_encoder=None, pick_place_model="pick_place_convdecoder",
              fusion_model=getattr(spec, "fusion_model", "concat_transformer"), depth=spec.depth, heads=spec.heads, mlp_ratio=spec.mlp_ratio, dropout=0.0,
              precision=precision, pretrained=False,
              tower_config=dict(tower_layers=spec.tower_layers, tower_heads=spec.tower_heads, tower_mlp=spec.tower_mlp,
                                vocab_size=spec.vocab_size, text_len=spec.text_len))
    if spec.name == "siglip_sequential":
        cfg["context_length"] = spec.context_length
    if spec.compute_mask:
        cfg["compute_mask"], cfg["detach_mask"] = True, spec.detach_mask
    return cfg


def build_model(spec: PolicySpec, precision="fp32", device="cuda", seed=3407, state_dict=None):
    from bifold_b200.models import Models
    sd = state_dict if state_dict is not None else make_state_dict(spec, seed=seed)
    model = Models.get_by_name(model_cfg(spec, precision), device=torch.device(device))
    model.load_state_dict(sd, strict=True)
    return model.to(device), sd


def to_device(sample, device="cuda"):
    return {k: v.to(device) for k, v in sample.items()}


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
