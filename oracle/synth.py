"""Seeded synthetic weights and samples shared by the oracle, the golden generator and the tests.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Everything here is numpy-PCG64 seeded so the very same
tensors can be regenerated on the GPU box, where /root/reference does not exist.

* ``PolicySpec``        – the shape of a BiFold policy (reference: bifold/conf/model/siglip*.yaml and
                          the HF SiglipConfig the reference instantiates at bifold/models/siglip.py:26).
* ``make_state_dict``   – a full ``state_dict`` with the reference's key names;
                          ``make_golden.py`` loads it into the real reference with ``strict=True``.
* ``make_sample``       – the sample dict of SURVEY.md §8(d) (bifold/data/processor.py:150-185,405-439).
"""
from __future__ import annotations

import dataclasses
from typing import Dict

import numpy as np
import torch


@dataclasses.dataclass
class PolicySpec:
    name: str = "siglip"            # "siglip" | "siglip_sequential"  (bifold/models/__init__.py:14-17)
    image_size: int = 224
    patch_size: int = 16
    dim: int = 768
    is_bimanual: bool = True
    lora: bool = True
    r: int = 8
    lora_alpha: int = 32
    lora_dropout: float = 0.0
    depth: int = 8                  # fusion transformer (bifold/conf/model/siglip.yaml:19-22)
    heads: int = 16
    mlp_ratio: int = 4
    context_length: int = 0         # T context frames (siglip_sequential only)
    # HF SiglipConfig (vision and text share hidden size: bifold/models/siglip.py:28-29)
    tower_layers: int = 12
    tower_heads: int = 12
    tower_mlp: int = 3072
    vocab_size: int = 32000
    text_len: int = 64
    tower_ln_eps: float = 1e-6
    compute_mask: bool = False      # PickPlaceConvDecoder(compute_mask=..., detach_mask=...) (bifold/models/pickplace.py:18-32)
    detach_mask: bool = False
    fusion_model: str = "concat_transformer"    # or "crossattention" (bifold/models/__init__.py:38-41, fusion.py:67-125)

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def heads_out(self):
        if self.is_bimanual:
            # loss order: bifold/losses/__init__.py:74-75
            return ["left_pick", "left_place", "right_pick", "right_place"]
        return ["pick", "place"]

    @property
    def decoder_names(self):
        return [f"{h}_decoder" for h in self.heads_out] + (["mask_head"] if self.compute_mask else [])

    def fusion_len(self) -> int:
        t = self.context_length if self.name == "siglip_sequential" else 0
        return (self.text_len + 1) + (t + 1) * (self.num_patches + 1)


# Named specs used by tests / golden vectors / bench.
def tiny_spec(**kw) -> PolicySpec:
    base = dict(image_size=64, patch_size=16, dim=64, depth=2, heads=4, tower_layers=2,
                tower_heads=4, tower_mlp=128, vocab_size=128, text_len=16)
    base.update(kw)
    return PolicySpec(**base)


def base224_spec(**kw) -> PolicySpec:
    return PolicySpec(**kw)


def so400m_spec(**kw) -> PolicySpec:
    # SigLIP-so400m shaped with patch 16 so the x16 decoder emits 384x384.
    base = dict(image_size=384, patch_size=16, dim=1152, heads=16, tower_layers=27, tower_heads=16,
                tower_mlp=4304)
    base.update(kw)
    return PolicySpec(**base)


def _t(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a.astype(np.float32)))


def make_state_dict(spec: PolicySpec, seed: int = 3407) -> Dict[str, torch.Tensor]:
    """Reference-keyed state_dict with seeded values (key surface: SURVEY.md appendix B)."""
    rng = np.random.default_rng(seed)
    D = spec.dim
    sd: Dict[str, torch.Tensor] = {}

    def normal(shape, std):
        return _t(rng.standard_normal(shape, dtype=np.float32) * std)

    def linear(prefix, out_f, in_f, bias=True, std=None):
        std = (0.7 / np.sqrt(in_f)) if std is None else std
        sd[prefix + ".weight"] = normal((out_f, in_f), std)
        if bias:
            sd[prefix + ".bias"] = normal((out_f,), 0.02)

    def layernorm(prefix, n):
        sd[prefix + ".weight"] = _t(1.0 + 0.1 * rng.standard_normal(n, dtype=np.float32))
        sd[prefix + ".bias"] = normal((n,), 0.05)

    tw = "siglip_model.model." if spec.lora else "siglip_model."
    sd["text_token"] = normal((1, 1, D), 1.0)
    sd["image_token"] = normal((1, 1, D), 1.0)
    if spec.name == "siglip_sequential":
        sd["context_pos_embedding"] = normal((1, spec.context_length * (spec.num_patches + 1), D), 1.0)
    sd[tw + "logit_scale"] = _t(np.array([0.5]))
    sd[tw + "logit_bias"] = _t(np.array([-0.25]))

    def tower_layers(prefix):
        for i in range(spec.tower_layers):
            lp = f"{prefix}.encoder.layers.{i}"
            layernorm(lp + ".layer_norm1", D)
            layernorm(lp + ".layer_norm2", D)
            for proj in ("q_proj", "k_proj", "v_proj", "out_proj"):
                if spec.lora and proj in ("q_proj", "v_proj"):
                    # peft key convention (not installed here; see oracle/ref_shims.py)
                    linear(f"{lp}.self_attn.{proj}.base_layer", D, D)
                    sd[f"{lp}.self_attn.{proj}.lora_A.siglip_adapter.weight"] = normal(
                        (spec.r, D), 1.0 / np.sqrt(D))
                    sd[f"{lp}.self_attn.{proj}.lora_B.siglip_adapter.weight"] = normal(
                        (D, spec.r), 0.02)
                else:
                    linear(f"{lp}.self_attn.{proj}", D, D)
            linear(lp + ".mlp.fc1", spec.tower_mlp, D)
            linear(lp + ".mlp.fc2", D, spec.tower_mlp)

    # text tower (HF modeling_siglip.py:189-226, 473-527)
    tp = tw + "text_model"
    sd[tp + ".embeddings.token_embedding.weight"] = normal((spec.vocab_size, D), 0.5)
    sd[tp + ".embeddings.position_embedding.weight"] = normal((spec.text_len, D), 0.5)
    tower_layers(tp)
    layernorm(tp + ".final_layer_norm", D)
    linear(tp + ".head", D, D)  # unused by BiFold
    # vision tower (HF modeling_siglip.py:116-185, 586-625)
    vp = tw + "vision_model"
    p = spec.patch_size
    sd[vp + ".embeddings.patch_embedding.weight"] = normal((D, 3, p, p), 0.7 / np.sqrt(3 * p * p))
    sd[vp + ".embeddings.patch_embedding.bias"] = normal((D,), 0.02)
    sd[vp + ".embeddings.position_embedding.weight"] = normal((spec.num_patches, D), 0.5)
    tower_layers(vp)
    layernorm(vp + ".post_layernorm", D)
    # pooling head (HF modeling_siglip.py:628-660): dead work for BiFold, kept for strict loading
    hp = vp + ".head"
    sd[hp + ".probe"] = normal((1, 1, D), 0.5)
    sd[hp + ".attention.in_proj_weight"] = normal((3 * D, D), 0.7 / np.sqrt(D))
    sd[hp + ".attention.in_proj_bias"] = normal((3 * D,), 0.02)
    linear(hp + ".attention.out_proj", D, D)
    layernorm(hp + ".layernorm", D)
    linear(hp + ".mlp.fc1", spec.tower_mlp, D)
    linear(hp + ".mlp.fc2", D, spec.tower_mlp)

    # fusion transformer (bifold/models/fusion.py:8-34, transformer.py:8-58, modules.py:8-30)
    fp = "pick_place.fusion"
    sd[fp + ".token_type_embeddings.weight"] = normal((2, D), 0.5)
    if spec.fusion_model == "crossattention":
        # torch.nn.MultiheadAttention(embed_dim=D, num_heads=heads, batch_first=True) parameters (fusion.py:75-77)
        sd[fp + ".cross_attention.in_proj_weight"] = normal((3 * D, D), 1.0 / np.sqrt(D))
        sd[fp + ".cross_attention.in_proj_bias"] = normal((3 * D,), 0.05)
        linear(fp + ".cross_attention.out_proj", D, D)
    for l in range(spec.depth if spec.fusion_model == "concat_transformer" else 0):
        lp = f"{fp}.transformer_encoder.layers.{l}"
        layernorm(lp + ".0.norm", D)
        linear(lp + ".0.fn.to_qkv", 3 * D, D, bias=False)
        linear(lp + ".0.fn.to_out.0", D, D)
        layernorm(lp + ".1.norm", D)
        linear(lp + ".1.fn.net.0", D * spec.mlp_ratio, D)
        linear(lp + ".1.fn.net.3", D, D * spec.mlp_ratio)
    # conv decoders (bifold/models/modules.py:33-74, pickplace.py:33-47)
    c1, c2 = int(D / 2), int(D / 4)
    chans = [(D, c1), (c1, c1), (c1, c2), (c2, c2), (c2, 1)]
    for dec in spec.decoder_names:
        for j, (cin, cout) in enumerate(chans):
            sd[f"pick_place.{dec}.decoder_net.{2 * j}.weight"] = normal(
                (cout, cin, 1, 1), 1.2 / np.sqrt(cin))
            sd[f"pick_place.{dec}.decoder_net.{2 * j}.bias"] = normal((cout,), 0.05)
    return sd


def trainable_keys(spec: PolicySpec, sd) -> list:
    """Keys with requires_grad=True in the reference (bifold/models/siglip.py:33-43)."""
    keys = []
    for k in sd:
        if k.startswith("siglip_model."):
            if spec.lora and (".lora_A." in k or ".lora_B." in k):
                keys.append(k)
        else:
            keys.append(k)
    return keys


def gaussian_target(h: int, w: int, cx: int, cy: int, sigma: float = 5.0) -> np.ndarray:
    """Gaussian map, max-normalised (bifold/data/processor.py:405-417)."""
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float32)
    g = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2.0 * sigma * sigma))
    return (g / g.max()).astype(np.float32)


def make_sample(spec: PolicySpec, batch: int, seed: int = 0, first_index: int = 0) -> Dict[str, torch.Tensor]:
    """Synthetic sample dict.  Sample ``first_index + i`` is generated from its own
    stream, so a rank's shard is a slice of the global batch (used by the DDP tests)."""
    H = W = spec.image_size
    out: Dict[str, list] = {}

    def put(k, v):
        out.setdefault(k, []).append(v)

    T = spec.context_length if spec.name == "siglip_sequential" else 0
    for i in range(first_index, first_index + batch):
        rng = np.random.default_rng([seed, i])
        put("rgb", rng.standard_normal((3, H, W), dtype=np.float32))
        put("instruction", rng.integers(0, spec.vocab_size, size=(spec.text_len,), dtype=np.int64))
        if T:
            put("rgb_context", rng.standard_normal((T, 3, H, W), dtype=np.float32))
            m = np.zeros((T,), dtype=np.int64)
            m[: i % (T + 1)] = 1       # leading ones, exercises the query-row mask quirk (F5)
            put("context_attention_mask", m)
        for j, head in enumerate(spec.heads_out):
            cx, cy = rng.integers(0, W), rng.integers(0, H)
            g = gaussian_target(H, W, int(cx), int(cy), sigma=max(1.0, 5.0 * H / 224))
            if i % 4 == 3 and head.startswith(("right", "place")) and j >= len(spec.heads_out) // 2:
                g = np.zeros_like(g)   # missing arm: all-zero map (bifold/data/processor.py:296)
            put(f"{head}_heatmap", g)
        # filled ellipse mask (B,1,H,W)
        ys, xs = np.mgrid[0:H, 0:W].astype(np.float32)
        ax, ay = rng.uniform(0.2, 0.45) * W, rng.uniform(0.2, 0.45) * H
        ex, ey = rng.uniform(0.4, 0.6) * W, rng.uniform(0.4, 0.6) * H
        put("mask", ((((xs - ex) / ax) ** 2 + ((ys - ey) / ay) ** 2) <= 1.0).astype(np.float32)[None])
    return {k: torch.from_numpy(np.stack(v)) for k, v in out.items()}
