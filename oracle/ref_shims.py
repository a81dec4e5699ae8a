"""Shims that let the REAL reference (/root/reference/bifold) import and run in the build container.

TEST INFRASTRUCTURE ONLY, and only usable where /root/reference exists (never on the GPU box).
Used by oracle/make_golden.py to generate tests/golden/*.npz.  Three gaps are bridged
(SURVEY.md §8c):

1. ``timm`` is absent: bifold/models/modules.py:2 imports ``timm...Block`` for the out-of-scope
   TransformerDecoder -> a stub module.
2. ``peft`` is absent: bifold/models/siglip.py:3,33-40 wraps q_proj/v_proj with LoRA -> a minimal
   restatement of peft's published LoRA algorithm with peft's module/key naming
   (``<proj>.base_layer``, ``<proj>.lora_A.<adapter>``, ``<proj>.lora_B.<adapter>``, wrapped model
   under ``.model``).  This part of the golden vectors is therefore pinned to peft's documented
   behaviour, not to peft's code.
3. No network / no checkpoints: ``AutoModel.from_pretrained`` -> ``SiglipModel(SiglipConfig(...))``
   with the shape taken from the active PolicySpec; weights are then overwritten from
   oracle.synth.make_state_dict with ``strict=True``.
"""
from __future__ import annotations

import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = "/root/reference"


class Cfg(dict):
    """Attribute-access mapping that supports ``del cfg.name`` and ``**cfg`` (what the reference's
    factories need from an omegaconf DictConfig)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def _install_timm_stub():
    if "timm" in sys.modules:
        return
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")

    class Block(nn.Module):  # never instantiated on the hot path
        def __init__(self, *a, **k):
            raise RuntimeError("timm stub")

    vt.Block = Block
    timm.models = models
    models.vision_transformer = vt
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.vision_transformer": vt})


def _install_peft_stub():
    if "peft" in sys.modules:
        return

    class LoraConfig:
        def __init__(self, r, lora_alpha, target_modules, lora_dropout):
            self.r, self.lora_alpha = r, lora_alpha
            self.target_modules, self.lora_dropout = list(target_modules), lora_dropout

    class LoraLinear(nn.Module):
        def __init__(self, base: nn.Linear, cfg: LoraConfig, adapter: str):
            super().__init__()
            self.base_layer = base
            self.adapter = adapter
            self.scaling = cfg.lora_alpha / cfg.r
            self.lora_dropout = nn.ModuleDict(
                {adapter: nn.Dropout(cfg.lora_dropout) if cfg.lora_dropout > 0 else nn.Identity()})
            self.lora_A = nn.ModuleDict({adapter: nn.Linear(base.in_features, cfg.r, bias=False)})
            self.lora_B = nn.ModuleDict({adapter: nn.Linear(cfg.r, base.out_features, bias=False)})
            nn.init.kaiming_uniform_(self.lora_A[adapter].weight, a=5 ** 0.5)
            nn.init.zeros_(self.lora_B[adapter].weight)

        def forward(self, x):
            a = self.adapter
            return self.base_layer(x) + self.lora_B[a](self.lora_A[a](self.lora_dropout[a](x))) * self.scaling

    class LoraModel(nn.Module):
        def __init__(self, model, config, adapter_name):
            super().__init__()
            self.model = model
            for p in model.parameters():
                p.requires_grad = False
            for parent in list(model.modules()):
                for name, child in list(parent.named_children()):
                    if name in config.target_modules and isinstance(child, nn.Linear):
                        setattr(parent, name, LoraLinear(child, config, adapter_name))

        def forward(self, *a, **k):
            return self.model(*a, **k)

        def __getattr__(self, name):
            try:
                return super().__getattr__(name)
            except AttributeError:
                return getattr(self._modules["model"], name)

    peft = types.ModuleType("peft")
    tuners = types.ModuleType("peft.tuners")
    lora = types.ModuleType("peft.tuners.lora")
    lora.LoraConfig, lora.LoraModel = LoraConfig, LoraModel
    peft.tuners, tuners.lora = tuners, lora
    sys.modules.update({"peft": peft, "peft.tuners": tuners, "peft.tuners.lora": lora})


_ACTIVE_SPEC = {}


def _install_automodel_shim():
    import transformers
    from transformers import SiglipConfig, SiglipModel

    def from_pretrained(name, *a, **k):
        s = _ACTIVE_SPEC["spec"]
        cfg = SiglipConfig(
            text_config=dict(hidden_size=s.dim, intermediate_size=s.tower_mlp, num_hidden_layers=s.tower_layers,
                             num_attention_heads=s.tower_heads, vocab_size=s.vocab_size,
                             max_position_embeddings=s.text_len, projection_size=s.dim),
            vision_config=dict(hidden_size=s.dim, intermediate_size=s.tower_mlp, num_hidden_layers=s.tower_layers,
                               num_attention_heads=s.tower_heads, image_size=s.image_size,
                               patch_size=s.patch_size))
        return SiglipModel(cfg)

    transformers.AutoModel.from_pretrained = staticmethod(from_pretrained)


def import_reference():
    """Returns the reference's (Models, Losses, Optimizers) factories."""
    import transformers  # noqa: F401  (must precede the timm stub)
    _install_timm_stub()
    _install_peft_stub()
    _install_automodel_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from bifold.losses import Losses
    from bifold.models import Models
    from bifold.optim import Optimizers
    return Models, Losses, Optimizers


def build_reference_model(spec, state_dict):
    """Instantiate the real reference model for ``spec`` and load ``state_dict`` strictly."""
    Models, _, _ = import_reference()
    _ACTIVE_SPEC["spec"] = spec
    cfg = Cfg(name=spec.name, image_size=spec.image_size, is_bimanual=spec.is_bimanual, requires_graph=False,
              patch_size=spec.patch_size, automodel_name="synthetic", dim=spec.dim, emb_dropout=0.0,
              lora=spec.lora, r=spec.r, lora_alpha=spec.lora_alpha, lora_dropout=spec.lora_dropout,
              target_modules=["q_proj", "v_proj"], threshold=0.5, text_encoder=None,
              pick_place_model="pick_place_convdecoder", fusion_model=getattr(spec, "fusion_model", "concat_transformer"),
              depth=spec.depth, heads=spec.heads, mlp_ratio=spec.mlp_ratio, dropout=0.0)
    if spec.name == "siglip_sequential":
        cfg["context_length"] = spec.context_length
    if getattr(spec, "compute_mask", False):
        cfg["compute_mask"], cfg["detach_mask"] = True, spec.detach_mask
    torch.manual_seed(3407)
    model = Models.get_by_name(cfg, device=torch.device("cpu"))
    missing, unexpected = model.load_state_dict(state_dict, strict=True)
    assert not missing and not unexpected
    return model
