"""CPU: the torch custom-op surface (torch.ops.bfk.*) — schemas, fake (meta) kernels, no-CPU-fallback behaviour — and the
pretrained-tower loading contract of the model constructor (bifold/models/siglip.py:26-29)."""
import tempfile

import pytest
import torch
from torch._subclasses.fake_tensor import FakeTensorMode

from bifold_b200 import _cabi as K
from bifold_b200 import ops  # noqa: F401
from bifold_b200.config import Cfg
from oracle.synth import tiny_spec
from tests.util import build_model, model_cfg

EXPECTED = ["policy_fwd", "policy_bwd", "linear", "linear_bwd", "layernorm", "layernorm_bwd", "attention", "attention_bwd",
            "bce_prob", "bce_prob_bwd", "argmax2d", "action", "gauss_targets", "normalize_u8"]


def test_ops_are_registered_with_schemas():
    for name in EXPECTED:
        op = getattr(torch.ops.bfk, name)
        assert op.default._schema.name == f"bfk::{name}"


def test_fake_kernels_give_shapes_without_a_gpu():
    spec = tiny_spec(name="siglip_sequential", is_bimanual=True, lora=True, context_length=2)
    model, _ = build_model(spec, "bf16", device="cpu")
    model._ensure_engine()
    B, H, heads = 3, spec.image_size, 4
    with FakeTensorMode() as m:
        f = lambda *s, dtype=torch.float32: m.from_tensor(torch.empty(*s, dtype=dtype))
        rgb, ids = f(B, 3, H, H), f(B, spec.text_len, dtype=torch.int64)
        ctx, cm = f(B, 2, 3, H, H), f(B, 2, dtype=torch.int64)
        targets = [f(B, H, H) for _ in range(heads)]
        heat, loss, key = torch.ops.bfk.policy_fwd(model._handle, rgb, ids, ctx, cm, targets, [], True, False, True)
        assert heat.shape == (heads, B, H, H) and loss.shape == (heads + 1,) and key.shape == (heads * B,) and key.dtype == torch.int64
        heat, loss, key = torch.ops.bfk.policy_fwd(model._handle, rgb, ids, ctx, cm, [], [], False, False, False)
        assert loss.numel() == 0 and key.numel() == 0
        qkv = f(2 * 10, 3 * 4 * 16, dtype=torch.bfloat16)
        out, lse = torch.ops.bfk.attention(qkv, None, 2, 10, 4, 16, 0.25)
        assert out.shape == (20, 64) and out.dtype == torch.bfloat16 and lse.shape == (2 * 4 * 10,)
        y, mean, rstd = torch.ops.bfk.layernorm(f(7, 64), f(64), f(64), 1e-6, True)
        assert y.shape == (7, 64) and y.dtype == torch.bfloat16 and mean.shape == (7,)
        assert torch.ops.bfk.linear(f(5, 32, dtype=torch.bfloat16), f(48, 32, dtype=torch.bfloat16), f(48), "gelu_tanh").shape == (5, 48)
        assert torch.ops.bfk.bce_prob(f(2, 8, 8), f(2, 8, 8), None).shape == ()
        assert torch.ops.bfk.argmax2d(f(6, 8, 8)).shape == (6,)
        xy, conf = torch.ops.bfk.action(f(4, B, H, H), f(4 * B, dtype=torch.int64), None, [0, 2], [1, 3], 2, 0.01)
        assert xy.shape == (4, B, 2) and xy.dtype == torch.int32 and conf.shape == (4, B)
        maps = torch.ops.bfk.gauss_targets(f(4, B, 2, dtype=torch.int32), H, H, 5.0)
        assert maps.shape == (4, B, H, H)
        img = torch.ops.bfk.normalize_u8(f(B, H, H, 3, dtype=torch.uint8), 0.5, 0.5)
        assert img.shape == (B, 3, H, H) and img.dtype == torch.float32


def test_ops_have_no_cpu_fallback():
    with pytest.raises((NotImplementedError, RuntimeError, K.BfkError)):
        torch.ops.bfk.layernorm(torch.randn(4, 64), torch.ones(64), torch.zeros(64), 1e-6, False)
    with pytest.raises((NotImplementedError, RuntimeError, K.BfkError)):
        torch.ops.bfk.bce_prob(torch.rand(2, 4, 4), torch.rand(2, 4, 4), None)


def _tiny_hf_dir():
    from transformers import SiglipConfig, SiglipModel
    cfg = SiglipConfig(text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, vocab_size=100,
                                        max_position_embeddings=16, projection_size=64),
                       vision_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2, image_size=32,
                                          patch_size=16))
    torch.manual_seed(0)
    hf = SiglipModel(cfg)
    d = tempfile.mkdtemp()
    hf.save_pretrained(d)
    return d, hf.state_dict()


def _cfg(name, lora, **kw):
    c = Cfg(name="siglip", image_size=32, is_bimanual=True, requires_graph=False, patch_size=16, automodel_name=name, dim=64, emb_dropout=0.0,
            lora=lora, r=8, lora_alpha=32, lora_dropout=0.0, target_modules=["q_proj", "v_proj"], threshold=0.5, text_encoder=None,
            pick_place_model="pick_place_convdecoder", fusion_model="concat_transformer", depth=2, heads=4, mlp_ratio=4, dropout=0.0)
    c.update(kw)
    return c


@pytest.mark.parametrize("lora", [True, False])
def test_pretrained_towers_are_loaded_like_the_reference(lora):
    """AutoModel.from_pretrained(automodel_name) (here: a local directory) fills the frozen towers, incl. peft's base_layer names;
    the tower shape comes from the checkpoint's config."""
    from bifold_b200.models import Models
    d, hf_sd = _tiny_hf_dir()
    model = Models.get_by_name(_cfg(d, lora), device=torch.device("cpu"))
    sd = model.state_dict()
    pre = "siglip_model.model." if lora else "siglip_model."
    base = "base_layer." if lora else ""
    for tower in ("text_model", "vision_model"):
        k = f"{tower}.encoder.layers.1.self_attn.v_proj."
        assert torch.equal(sd[pre + k + base + "weight"], hf_sd[k + "weight"])
        assert torch.equal(sd[pre + f"{tower}.encoder.layers.0.mlp.fc1.bias"], hf_sd[f"{tower}.encoder.layers.0.mlp.fc1.bias"])
    assert torch.equal(sd[pre + "vision_model.embeddings.patch_embedding.weight"], hf_sd["vision_model.embeddings.patch_embedding.weight"])
    assert (model.cfg.tower_layers, model.cfg.tower_heads, model.cfg.tower_mlp, model.cfg.vocab_size, model.cfg.text_len) == (2, 2, 128, 100, 16)
    frozen = [n for n, p in model.named_parameters() if "siglip_model" in n and p.requires_grad and ".lora_" not in n]
    assert not frozen


def test_pretrained_failure_raises_unless_opted_out(monkeypatch):
    from bifold_b200.models import Models
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    with pytest.raises(RuntimeError, match="pretrained=False"):
        Models.get_by_name(_cfg("google/siglip-definitely-not-cached", True), device=torch.device("cpu"))
    model = Models.get_by_name(_cfg("google/siglip-definitely-not-cached", True, pretrained=False,
                                    tower_config=dict(tower_layers=2, tower_heads=2, tower_mlp=128, vocab_size=100, text_len=16)),
                               device=torch.device("cpu"))
    assert model.cfg.tower_layers == 2


def test_action_container_contract():
    from bifold_b200.models import Action
    a = Action(left_pick=1, right_pick=2, left_place=3, right_place=4)
    assert (a.left_pick, a.right_place) == (1, 4)
    assert Action(pick=1, place=2).place == 2
    with pytest.raises((ValueError, TypeError)):
        Action(pick=1)
