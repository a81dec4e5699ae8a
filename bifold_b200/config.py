"""Shape of a BiFold policy as the engine sees it (reference: bifold/conf/model/siglip*.yaml plus the
HF SiglipConfig behind ``automodel_name``, bifold/models/siglip.py:26-29)."""
from __future__ import annotations

import dataclasses


class Cfg(dict):
    """Attribute-access config supporting ``del cfg.name`` and ``**cfg`` — what the factories need from
    an omegaconf DictConfig (hydra/omegaconf are not required)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


@dataclasses.dataclass
class PolicyConfig:
    name: str = "siglip"                 # "siglip" | "siglip_sequential"
    image_size: int = 224
    patch_size: int = 16
    dim: int = 768
    is_bimanual: bool = True
    lora: bool = True
    r: int = 8
    lora_alpha: float = 32
    lora_dropout: float = 0.0
    depth: int = 8
    heads: int = 16
    mlp_ratio: int = 4
    context_length: int = 0
    tower_layers: int = 12
    tower_heads: int = 12
    tower_mlp: int = 3072
    vocab_size: int = 32000
    text_len: int = 64
    tower_ln_eps: float = 1e-6
    compute_mask: bool = False           # PickPlaceConvDecoder(compute_mask=True): mask head + gated pick heatmaps
    detach_mask: bool = False
    fusion_model: str = "concat_transformer"   # or "crossattention" (bifold/models/fusion.py:67-125)

    @property
    def grid(self) -> int:
        return self.image_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid

    @property
    def sequential(self) -> bool:
        return self.name == "siglip_sequential"

    @property
    def heads_out(self):
        # head order = loss order, bifold/losses/__init__.py:74-75,94
        if self.is_bimanual:
            return ["left_pick", "left_place", "right_pick", "right_place"]
        return ["pick", "place"]

    @property
    def decoder_prefixes(self):
        """Parameter prefixes of the ConvDecoders, policy heads first (bifold/models/pickplace.py:27-47)."""
        return [f"pick_place.{h}_decoder" for h in self.heads_out] + (["pick_place.mask_head"] if self.compute_mask else [])

    @property
    def pick_head_bits(self) -> int:
        return sum(1 << i for i, h in enumerate(self.heads_out) if h.endswith("pick"))

    def fusion_len(self, t: int) -> int:
        return (self.text_len + 1) + (t + 1) * (self.num_patches + 1)


# Known SigLIP checkpoints -> tower shape (used when no local checkpoint can be loaded; HF config.json values)
KNOWN_TOWERS = {
    "base": dict(dim=768, tower_layers=12, tower_heads=12, tower_mlp=3072),
    "large": dict(dim=1024, tower_layers=24, tower_heads=16, tower_mlp=4096),
    "so400m": dict(dim=1152, tower_layers=27, tower_heads=16, tower_mlp=4304),
}


def tower_shape_from_name(automodel_name: str, dim: int) -> dict:
    for key, shape in KNOWN_TOWERS.items():
        if key in (automodel_name or "") and shape["dim"] == dim:
            return dict(shape)
    for shape in KNOWN_TOWERS.values():
        if shape["dim"] == dim:
            return dict(shape)
    return dict(dim=dim, tower_layers=12, tower_heads=max(1, dim // 64), tower_mlp=4 * dim)
