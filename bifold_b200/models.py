"""Host-side mirror of ``bifold.models`` for the policy hot path (drop-in interface).

Same factories, constructor kwargs (= keys of bifold/conf/model/siglip*.yaml), ``forward`` /
``get_action`` signatures, attributes and ``state_dict`` key names as the reference:

* ``Models.get_by_name`` / ``Components.get_by_name``  <- bifold/models/__init__.py:12-43
* ``BaseModel``                                         <- bifold/models/__init__.py:46-173
* ``SigLip`` / ``SiglipSequential``                     <- bifold/models/siglip.py:10-137

The nn.Modules below only *hold* parameters (with the reference's names, shapes and default
initialisation).  All arithmetic of ``forward`` runs in ``bifold_b200.engine.Engine`` on our CUDA
kernels; there is no eager fallback, so ``forward`` raises on a non-CUDA device.
"""
from __future__ import annotations

import math
import sys
from typing import Dict

import numpy as np
import torch
from torch import nn

from . import _cabi as K
from . import ops
from .config import PolicyConfig, tower_shape_from_name
from .engine import Engine

DUMMY_PICK = -np.ones(2)


try:                                   # drop-in: hand the caller the reference's own container (bifold/env/__init__.py:1-22)
    from bifold.env import Action
except Exception:                      # the reference package is not importable (GPU box, tests): same keyword contract
    class Action:
        """``Action(pick=, place=)`` or ``Action(left_pick=, right_pick=, left_place=, right_place=)``."""
        _FIELDS = {2: ("pick", "place"), 4: ("left_pick", "right_pick", "left_place", "right_place")}

        def __init__(self, **kwargs):
            want = self._FIELDS.get(len(kwargs))
            if want is None or set(kwargs) != set(want):
                raise ValueError(f"Cannot instantiate action with {kwargs}")
            self.__dict__.update(kwargs)


class PolicyOutput(dict):
    """The reference's output dict plus (non-key) handles used by the fused loss / action paths."""
    fused_loss = None      # tensor [heads] with autograd, or None
    fused_targets = None   # tuple of target data_ptrs the fused loss was computed against
    heat_all = None        # [heads, B, H, W] tensor the per-head views come from


# ----------------------------------------------------------------------------------------------------
# parameter containers with the reference's module tree (names == reference state_dict keys)
# ----------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("parameter container only; the computation runs in bifold_b200.engine")


class _LoraLinear(_Holder):
    """peft ``lora.Linear`` key layout: base_layer + lora_A/lora_B ModuleDicts keyed by adapter name."""

    def __init__(self, in_f, out_f, r, adapter):
        super().__init__()
        self.base_layer = nn.Linear(in_f, out_f)
        self.lora_A = nn.ModuleDict({adapter: nn.Linear(in_f, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter: nn.Linear(r, out_f, bias=False)})
        nn.init.kaiming_uniform_(self.lora_A[adapter].weight, a=math.sqrt(5))
        nn.init.zeros_(self.lora_B[adapter].weight)


def _tower_layer(D, mlp, lora, r, adapter):
    layer = _Holder()
    layer.layer_norm1 = nn.LayerNorm(D)
    attn = _Holder()
    for proj in ("k_proj", "v_proj", "q_proj", "out_proj"):
        if lora and proj in ("q_proj", "v_proj"):
            setattr(attn, proj, _LoraLinear(D, D, r, adapter))
        else:
            setattr(attn, proj, nn.Linear(D, D))
    layer.self_attn = attn
    layer.layer_norm2 = nn.LayerNorm(D)
    mlp_m = _Holder()
    mlp_m.fc1, mlp_m.fc2 = nn.Linear(D, mlp), nn.Linear(mlp, D)
    layer.mlp = mlp_m
    return layer


def _encoder(cfg: PolicyConfig, lora, adapter):
    enc = _Holder()
    enc.layers = nn.ModuleList([_tower_layer(cfg.dim, cfg.tower_mlp, lora, cfg.r, adapter) for _ in range(cfg.tower_layers)])
    return enc


class _SiglipShell(_Holder):
    """HF SiglipModel parameter tree (HF modeling_siglip.py: text_model / vision_model / logit_*)."""

    def __init__(self, cfg: PolicyConfig, lora: bool, adapter: str):
        super().__init__()
        D = cfg.dim
        text = _Holder()
        emb = _Holder()
        emb.token_embedding = nn.Embedding(cfg.vocab_size, D)
        emb.position_embedding = nn.Embedding(cfg.text_len, D)
        text.embeddings = emb
        text.encoder = _encoder(cfg, lora, adapter)
        text.final_layer_norm = nn.LayerNorm(D)
        text.head = nn.Linear(D, D)                    # unused by BiFold (HF:521-522)
        self.text_model = text
        vis = _Holder()
        vemb = _Holder()
        vemb.patch_embedding = nn.Conv2d(3, D, cfg.patch_size, cfg.patch_size)
        vemb.position_embedding = nn.Embedding(cfg.num_patches, D)
        vis.embeddings = vemb
        vis.encoder = _encoder(cfg, lora, adapter)
        vis.post_layernorm = nn.LayerNorm(D)
        head = _Holder()                               # pooling head, unused by BiFold (HF:620)
        head.probe = nn.Parameter(torch.randn(1, 1, D))
        head.attention = nn.MultiheadAttention(D, cfg.tower_heads, batch_first=True)
        head.layernorm = nn.LayerNorm(D)
        hm = _Holder()
        hm.fc1, hm.fc2 = nn.Linear(D, cfg.tower_mlp), nn.Linear(cfg.tower_mlp, D)
        head.mlp = hm
        vis.head = head
        self.vision_model = vis
        self.logit_scale = nn.Parameter(torch.randn(1))
        self.logit_bias = nn.Parameter(torch.randn(1))


class _LoraModelShell(_Holder):
    """peft ``LoraModel`` wrapper: wrapped model under ``.model`` and everything but LoRA frozen."""

    def __init__(self, model):
        super().__init__()
        self.model = model
        for n, p in model.named_parameters():
            p.requires_grad = ".lora_A." in n or ".lora_B." in n


class _PreNorm(_Holder):       # bifold/models/modules.py:8-15
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn


class _Attention(_Holder):     # bifold/models/transformer.py:8-24
    def __init__(self, dim, heads, dim_head):
        super().__init__()
        inner = dim_head * heads
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(0.0))


class _FeedForward(_Holder):   # bifold/models/modules.py:18-30
    def __init__(self, dim, hidden):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(0.0), nn.Linear(hidden, dim), nn.Dropout(0.0))


class _Transformer(_Holder):   # bifold/models/transformer.py:45-58
    def __init__(self, dim, depth, heads, dim_head, mlp_dim):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([_PreNorm(dim, _Attention(dim, heads, dim_head)), _PreNorm(dim, _FeedForward(dim, mlp_dim))])
            for _ in range(depth)])


class ConcatTransformer(_Holder):   # bifold/models/fusion.py:8-34
    def __init__(self, dim, heads, depth, dropout=0.0, mlp_ratio=4, num_modalities=2, num_registers=0, **kwargs):
        super().__init__()
        if dropout != 0.0 or num_registers != 0:
            raise K.BfkError("concat_transformer: dropout and registers are not on the accelerated path")
        self.num_modalities = num_modalities
        self.token_type_embeddings = nn.Embedding(num_modalities, dim)
        self.token_type_embeddings.weight.data.normal_(mean=0.0, std=0.02)   # bifold/models/utils.py:86-89
        self.dim_head = int(dim / heads)
        self.mlp_dim = dim * mlp_ratio
        self.transformer_encoder = _Transformer(dim, depth, heads, self.dim_head, self.mlp_dim)
        self.registers = None


class CrossAttention(_Holder):      # bifold/models/fusion.py:67-78
    """Alternate fusion module: the image tokens (+ type embedding) query the text (+ context) tokens through ONE
    torch.nn.MultiheadAttention layer.  The torch module is only the parameter container (same names and initialisation as the
    reference: in_proj_weight / in_proj_bias / out_proj.*); the math runs in Engine._xattn_fwd / _xattn_bwd."""

    def __init__(self, dim, heads, depth=None, dropout=0.0, num_modalities=2, **kwargs):
        super().__init__()
        if dropout != 0.0:
            raise K.BfkError("crossattention: attention dropout is not on the accelerated path")
        self.num_modalities = num_modalities
        self.token_type_embeddings = nn.Embedding(num_modalities, dim)
        self.token_type_embeddings.weight.data.normal_(mean=0.0, std=0.02)   # bifold/models/utils.py:86-89
        self.num_heads = heads
        self.cross_attention = nn.MultiheadAttention(embed_dim=dim, num_heads=heads, dropout=dropout, batch_first=True)


class ConvDecoder(_Holder):         # bifold/models/modules.py:33-74
    def __init__(self, input_dim, output_dim=1):
        super().__init__()
        c1, c2 = int(input_dim / 2), int(input_dim / 4)
        chans = [(input_dim, c1), (c1, c1), (c1, c2), (c2, c2), (c2, output_dim)]
        mods = []
        for i, (ci, co) in enumerate(chans):
            mods.append(nn.Conv2d(ci, co, kernel_size=1))
            if i != 4:
                mods.append(nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False))
        self.decoder_net = nn.Sequential(*mods)


class PickPlaceConvDecoder(_Holder):   # bifold/models/pickplace.py:10-47
    def __init__(self, dim, is_bimanual, fusion_model, num_patches, patch_size, compute_mask=False, detach_mask=False,
                 **kwargs):
        super().__init__()
        self.is_bimanual = is_bimanual
        self.fusion = Components.get_by_name(fusion_model, dim=dim, **kwargs)
        self.num_patches_sqrt = int(math.sqrt(num_patches))
        if compute_mask:                 # bifold/models/pickplace.py:27-32
            self.detach_mask = detach_mask
            self.mask_head = ConvDecoder(dim)
        else:
            self.mask_head = None
        if is_bimanual:
            self.left_pick_decoder, self.right_pick_decoder = ConvDecoder(dim), ConvDecoder(dim)
            self.left_place_decoder, self.right_place_decoder = ConvDecoder(dim), ConvDecoder(dim)
        else:
            self.pick_decoder, self.place_decoder = ConvDecoder(dim), ConvDecoder(dim)


class Models:
    @staticmethod
    def get_by_name(cfg, *args, **kwargs):
        if cfg.name == "siglip":
            Model = SigLip
        elif cfg.name == "siglip_sequential":
            Model = SiglipSequential
        elif cfg.name in ("rgb_clip", "text_unet"):
            raise ValueError(f"Model {cfg.name} is an ablation baseline outside the accelerated hot path")
        else:
            raise ValueError(f"Model {cfg.name} not recognized")
        del cfg.name
        return Model(*args, **cfg, **kwargs)


class Components:
    @staticmethod
    def get_by_name(name, *args, **kwargs):
        if name == "pick_place_convdecoder":
            return PickPlaceConvDecoder(*args, **kwargs)
        if name == "concat_transformer":
            return ConcatTransformer(*args, **kwargs)
        if name == "crossattention":
            return CrossAttention(*args, **kwargs)
        if name == "pick_place_transdecoder":
            raise ValueError(f"Model {name} is not on the accelerated hot path (no shipped config selects it; its TransformerDecoder "
                             "needs timm)")
        raise ValueError(f"Model {name} not recognized")


# ----------------------------------------------------------------------------------------------------
# The whole policy is ONE torch custom op with a registered backward (bifold_b200/ops.py):
#   torch.ops.bfk.policy_fwd -> Engine.forward,   torch.ops.bfk.policy_bwd -> Engine.backward
# ----------------------------------------------------------------------------------------------------
class BaseModel(nn.Module):
    def __init__(self, image_size, is_bimanual, device, threshold=0.01, requires_graph=False, constrain_pick_mask=True,
                 **kwargs):
        super().__init__()
        self.image_size = image_size
        self.is_bimanual = is_bimanual
        self.device = device
        self.threshold = 0.01          # the reference ignores its argument (bifold/models/__init__.py:61)
        self.requires_graph = requires_graph
        self.constrain_pick_mask = constrain_pick_mask
        if requires_graph:
            raise K.BfkError("requires_graph=True (point-cloud picks) is not on the accelerated path")
        if not self.constrain_pick_mask:
            print("This run will not constrain the pick to the mask")

    def frozen_submodule(self, submodule_name):
        for n, p in self.named_parameters():
            if submodule_name in n:
                p.requires_grad = False

    def frozen_all(self):
        for p in self.parameters():
            p.requires_grad = False

    # -- get_action: bifold/models/__init__.py:79-173 with argmax / mask snap / gate on device -------------
    def action_on_device(self, sample):
        """Device part of ``get_action``: forward + argmax keys + mask snap + arm gate, nothing leaves the GPU and nothing
        synchronises (capturable: ``bifold_b200.graphs.GraphedAction``).  -> (xy int32 [heads,B,2], conf [heads,B], output)."""
        with torch.no_grad():
            output = self.forward(sample, _want_argmax=True)
        names = self.cfg.heads_out
        heat = output.heat_all
        heads, B, H, W = heat.shape
        if self.is_bimanual:
            pick = [names.index("left_pick"), names.index("right_pick")]
            place = [names.index("left_place"), names.index("right_place")]
            n_arms, mask = 2, sample["mask"]
        else:
            pick, place = [names.index("pick"), 0], [names.index("place"), 0]
            n_arms = 1
            mask = sample.get("mask") if self.constrain_pick_mask else None
        if mask is not None:
            mask = mask.to(device=heat.device, dtype=torch.float32).reshape(B, H, W).contiguous()
        xy, conf = torch.ops.bfk.action(heat, self._last_key, mask, pick, place, n_arms, float(self.threshold))
        self.engine.k.launches += 1
        return xy, conf, output

    def get_action(self, sample, return_raw_output=False):
        xy, _, output = self.action_on_device(sample)
        action = self.action_from_table(xy)
        if return_raw_output:
            return action, output
        return action

    def action_from_table(self, xy):
        """(heads,B,2) int32 pixel table (host or device) -> the reference's Action of int64 (B,2) arrays."""
        xy = xy.cpu().numpy().astype(np.int64)
        return Action(**{n: xy[i] for i, n in enumerate(self.cfg.heads_out)})


class SigLip(BaseModel):
    def __init__(self, patch_size, dim, lora, r, lora_alpha, lora_dropout, automodel_name, target_modules,
                 pick_place_model, precision="bf16", tower_config=None, pretrained=True, **kwargs):
        """``pretrained`` (not a reference key; default = the reference's behaviour): the towers are loaded with
        ``AutoModel.from_pretrained(automodel_name)`` (HF cache, local directory or hub id) exactly like
        bifold/models/siglip.py:26 and construction RAISES when that fails.  ``pretrained=False`` is the explicit
        opt-in for randomly initialised towers of the shape ``automodel_name`` / ``tower_config`` describe
        (benchmarks, tests, or when the weights arrive through ``load_state_dict`` of a BiFold checkpoint)."""
        super().__init__(**kwargs)
        self.patch_size = patch_size
        self.hidden_size = dim
        self.num_patches = (self.image_size // self.patch_size) ** 2
        if lora and sorted(target_modules) != ["q_proj", "v_proj"]:
            raise K.BfkError("LoRA target_modules other than [q_proj, v_proj] are not on the accelerated path")
        hf_state = None
        if pretrained:
            hf_state, tower = self._load_pretrained(automodel_name, dim)
        else:
            print(f"bifold_b200: pretrained=False -> SigLIP towers of '{automodel_name}' keep their RANDOM initialisation "
                  "(load a BiFold checkpoint with load_state_dict, or pass pretrained=True)", file=sys.stderr, flush=True)
            tower = dict(tower_shape_from_name(automodel_name, dim))
        tower.update(tower_config or {})
        tower.pop("dim", None)
        fusion_kw = {k: kwargs[k] for k in ("depth", "heads", "mlp_ratio", "fusion_model") if k in kwargs}
        if fusion_kw.get("fusion_model", "concat_transformer") == "crossattention" and kwargs.get("compute_mask", False):
            raise K.BfkError("crossattention with compute_mask=True is not on the accelerated path")
        self.cfg = PolicyConfig(name="siglip", image_size=self.image_size, patch_size=patch_size, dim=dim,
                                is_bimanual=self.is_bimanual, lora=bool(lora), r=r, lora_alpha=lora_alpha,
                                lora_dropout=lora_dropout, compute_mask=bool(kwargs.get("compute_mask", False)),
                                detach_mask=bool(kwargs.get("detach_mask", False)), **fusion_kw, **tower)
        siglip = _SiglipShell(self.cfg, bool(lora), "siglip_adapter")
        if hf_state is not None:
            self._install_pretrained(siglip, hf_state, bool(lora), automodel_name)
        if lora:
            self.siglip_model = _LoraModelShell(siglip)
        else:
            self.siglip_model = siglip
            self.frozen_submodule("siglip_model")
        self.text_token = nn.Parameter(torch.randn(1, 1, dim))
        self.image_token = nn.Parameter(torch.randn(1, 1, dim))
        self.pick_place = Components.get_by_name(pick_place_model, dim=dim, patch_size=patch_size,
                                                 num_patches=self.num_patches, **kwargs)
        self.precision = precision
        self.engine = None
        self._last_key = None

    def _load_pretrained(self, automodel_name, dim):
        """AutoModel.from_pretrained like bifold/models/siglip.py:26-29: returns (state_dict, tower shape); raises on failure."""
        from transformers import AutoModel
        try:
            hf = AutoModel.from_pretrained(automodel_name)
        except Exception as e:
            raise RuntimeError(f"cannot load the SigLIP towers '{automodel_name}' (AutoModel.from_pretrained failed: {type(e).__name__}: "
                               f"{e}); pass pretrained=False to build randomly initialised towers on purpose") from e
        tc, vc = hf.config.text_config, hf.config.vision_config
        assert dim == vc.hidden_size and dim == tc.hidden_size            # bifold/models/siglip.py:28-29
        if vc.image_size != self.image_size or vc.patch_size != self.patch_size:
            raise ValueError(f"'{automodel_name}' is a {vc.image_size}px / patch {vc.patch_size} checkpoint, the model config says "
                             f"{self.image_size}px / patch {self.patch_size}")
        same = ("num_hidden_layers", "num_attention_heads", "intermediate_size", "layer_norm_eps")
        if any(getattr(tc, k) != getattr(vc, k) for k in same):
            raise K.BfkError("text and vision towers of different depth/width are not on the accelerated path")
        if getattr(vc, "hidden_act", "gelu_pytorch_tanh") != "gelu_pytorch_tanh":
            raise K.BfkError(f"tower activation {vc.hidden_act} is not on the accelerated path")
        tower = dict(tower_layers=vc.num_hidden_layers, tower_heads=vc.num_attention_heads, tower_mlp=vc.intermediate_size,
                     vocab_size=tc.vocab_size, text_len=tc.max_position_embeddings, tower_ln_eps=float(vc.layer_norm_eps))
        return {k: v.detach().clone() for k, v in hf.state_dict().items()}, tower

    @staticmethod
    def _install_pretrained(shell, hf_state, lora, automodel_name):
        """Strict load of the HF SiglipModel weights into the parameter tree (peft renames wrapped linears to *.base_layer.*)."""
        sd = {}
        for k, v in hf_state.items():
            if lora:
                for proj in (".q_proj.", ".v_proj."):
                    if proj in k and ".encoder.layers." in k:
                        k = k.replace(proj, proj + "base_layer.")
            sd[k] = v
        res = shell.load_state_dict(sd, strict=False)
        missing = [k for k in res.missing_keys if ".lora_A." not in k and ".lora_B." not in k]
        if missing or res.unexpected_keys:
            raise RuntimeError(f"'{automodel_name}' does not match the SigLIP tower layout: missing {missing[:5]}, "
                               f"unexpected {list(res.unexpected_keys)[:5]}")

    # ------------------------------------------------------------------------------------------------
    def _ensure_engine(self):
        if self.engine is None:
            self._param_names = [n for n, _ in self.named_parameters()]
            self._param_list = [p for _, p in self.named_parameters()]
            self.engine = Engine(self.cfg, dict(zip(self._param_names, self._param_list)), self.precision)
            if getattr(self, "_handle", None) is None:
                self._handle = ops.register_model(self)
        return self.engine

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self.engine = None       # parameters may have been replaced (.to(device))
        return r

    def forward(self, x, _want_argmax=False):
        rgb = x["rgb"]
        if not rgb.is_cuda:
            raise K.BfkError("bifold_b200 runs on CUDA only (no CPU fallback): move the sample to the GPU")
        self._ensure_engine()
        cfg = self.cfg
        # fused BCE when the targets travel with the sample (training loop: bifold/__main__.py:135-137)
        targets = None
        if self.training and torch.is_grad_enabled() and not cfg.compute_mask:
            ts = [x.get(f"{h}_heatmap") for h in cfg.heads_out]
            if all(t is not None and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()
                   and tuple(t.shape) == (rgb.shape[0], cfg.image_size, cfg.image_size) for t in ts):
                targets = ts
        need_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list)
        heat, loss, key = torch.ops.bfk.policy_fwd(self._handle, rgb, x["instruction"], x.get("rgb_context") if cfg.sequential else None,
                                                   x.get("context_attention_mask") if cfg.sequential else None,
                                                   targets or [], self._param_list, self.training, need_grad, _want_argmax)
        self._last_key = key if _want_argmax else None
        # ConcatTransformer returns None here (fusion.py:64); CrossAttention the head-averaged weights (B, Nq, Nc), no autograd
        out = PolicyOutput(attn_weights=getattr(self.engine, "last_attn_weights", None) if cfg.fusion_model == "crossattention" else None)
        if cfg.compute_mask:              # key order of the reference's return_dict (bifold/models/pickplace.py:64-69)
            out["mask_heatmap"] = heat[len(cfg.heads_out)]
        for i, h in enumerate(cfg.heads_out):
            out[f"{h}_heatmap"] = heat[i]
        out.heat_all = heat[:len(cfg.heads_out)]
        if targets is not None:
            out.fused_loss = loss
            out.fused_targets = tuple(t.data_ptr() for t in targets)
        return out


class SiglipSequential(SigLip):
    def __init__(self, context_length, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.cfg.name = "siglip_sequential"
        self.cfg.context_length = context_length
        self.context_pos_embedding = nn.Parameter(torch.randn(1, context_length * (self.num_patches + 1), self.hidden_size))

    def forward(self, x, _want_argmax=False):
        t = x["rgb_context"].shape[1]
        if t != self.cfg.context_length:
            # the reference adds context_pos_embedding of fixed length (bifold/models/siglip.py:117)
            raise ValueError(f"rgb_context has {t} frames, context_length is {self.cfg.context_length}")
        return super().forward(x, _want_argmax)
