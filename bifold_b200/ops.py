"""torch custom ops (``torch.ops.bfk.*``) over the C ABI of libbifold_b200.so.

north_star: "host code stays in Python/PyTorch and calls hand-written sm_100a CUDA kernels through a thin C-ABI exposed
as torch custom ops".  Every op below is a ``torch.library.custom_op`` whose CUDA implementation fills a ``bfk_*_args``
struct (``_cabi.py``) and calls the ``extern "C"`` entry point on the CURRENT stream of the calling thread — which is what
makes them safe on PyTorch's autograd worker thread and inside CUDA-graph capture.  Each op has a fake (meta) kernel, and
the differentiable ones register their backward with ``register_autograd``; the backward is itself an op.

Whole-policy ops (what ``bifold_b200.models`` calls):
    bfk::policy_fwd   SigLip/SiglipSequential.forward  (bifold/models/siglip.py:56-137) incl. fused BCE + argmax keys
    bfk::policy_bwd   its explicit backward (all trainable gradients, views of one flat arena)
Single ops (usable from the reference's own nn.Modules):
    bfk::linear       nn.Linear (+GELU tanh/erf)         HF:285-287,309,324-326; bifold/models/transformer.py:19-24
    bfk::layernorm    nn.LayerNorm                        HF:334-336; bifold/models/modules.py:11
    bfk::attention    softmax(QK^T s [query-row mask])V   HF:293-306; bifold/models/transformer.py:31-41
    bfk::bce_prob     nn.BCELoss()(p, t [* mask])         bifold/losses/__init__.py:80-83,112-117
    bfk::dice_loss / bfk::focal_loss  mask losses         bifold/losses/__init__.py:126-155
    bfk::argmax2d     first-max index per map             bifold/models/utils.py:66-73
    bfk::action       argmax + mask snap + arm gate       bifold/models/__init__.py:108-141, utils.py:80-83
    bfk::gauss_targets / bfk::normalize_u8                bifold/data/processor.py:405-439 / :143-148 (GPU input stage)

There is no CPU implementation: calling an op with CPU tensors raises (no fallback).
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _cabi as K

_lib = torch.library.Library("bfk", "FRAGMENT")   # keeps the namespace alive; ops are defined with custom_op below


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return K.F32
    if t.dtype == torch.bfloat16:
        return K.BF16
    raise TypeError(f"bfk ops take float32 or bfloat16 tensors, got {t.dtype}")


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise K.BfkError("bfk ops run on CUDA tensors only (no CPU fallback)")


# =====================================================================================================================
# whole policy
# =====================================================================================================================
_MODELS: "weakref.WeakValueDictionary[int, torch.nn.Module]" = weakref.WeakValueDictionary()
_next_handle = [1]
_BWD_SLOT: dict = {}      # (handle, ticket) -> saved forward context, only for the duration of one policy_bwd call
_TICKET = [0]


def register_model(model) -> int:
    """Custom ops cannot take Python objects: the policy ops address their model (its Engine) by an integer handle."""
    h = _next_handle[0]
    _next_handle[0] += 1
    _MODELS[h] = model
    return h


def _model(handle: int):
    m = _MODELS.get(handle)
    if m is None:
        raise K.BfkError(f"bfk::policy_*: unknown model handle {handle}")
    return m


@torch.library.custom_op("bfk::policy_fwd", mutates_args=(), device_types="cuda")
def policy_fwd(handle: int, rgb: Tensor, instruction: Tensor, rgb_context: Optional[Tensor],
               context_attention_mask: Optional[Tensor], targets: Sequence[Tensor], params: Sequence[Tensor],
               training: bool, need_grad: bool, want_argmax: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (heat [heads,B,H,W] fp32 post-sigmoid (compute_mask: [heads+1,...], pick maps gated, last map = mask_heatmap), loss [heads+1] (per-head BCE, then their sum; empty without targets),
    argmax keys [heads*B] int64 (empty unless want_argmax)).  ``params`` only carries the autograd edges; the engine reads
    the parameters it was built on."""
    model = _model(handle)
    sample = {"rgb": rgb, "instruction": instruction}
    if rgb_context is not None:
        sample["rgb_context"], sample["context_attention_mask"] = rgb_context, context_attention_mask
    heat, loss, key, ectx = model.engine.forward(sample, need_grad, training, list(targets) if len(targets) else None, want_argmax)
    model._pending_ectx = ectx
    if loss is None:
        loss = heat.new_zeros(0)
    if key is None:
        key = torch.zeros(0, device=heat.device, dtype=torch.int64)
    return heat, loss, key


@policy_fwd.register_fake
def _(handle, rgb, instruction, rgb_context, context_attention_mask, targets, params, training, need_grad, want_argmax):
    cfg = _model(handle).cfg
    B, heads, H = rgb.shape[0], len(cfg.heads_out), cfg.image_size
    return (rgb.new_empty((heads + int(cfg.compute_mask), B, H, H), dtype=torch.float32),
            rgb.new_empty((heads + 1 if len(targets) else 0,), dtype=torch.float32),
            rgb.new_empty((heads * B if want_argmax else 0,), dtype=torch.int64))


@torch.library.custom_op("bfk::policy_bwd", mutates_args=(), device_types="cuda")
def policy_bwd(handle: int, ticket: int, dheat: Optional[Tensor], dloss: Optional[Tensor]) -> Tensor:
    """Explicit backward of the policy -> ONE flat fp32 gradient arena; the gradient of the i-th trainable parameter
    (``engine.trainable_names()`` order) is the view ``arena[off_i : off_i + numel_i]`` with every tensor padded to 8 floats
    (``arena_views`` slices it).  ``ticket`` names the saved forward context."""
    model = _model(handle)
    ectx = _BWD_SLOT.pop((handle, ticket), None)
    if ectx is None:
        raise RuntimeError("backward through a policy forward that ran without grad (or was already consumed)")
    model.engine.backward(ectx, dheat, dloss)
    return model.engine.last_arena


@policy_bwd.register_fake
def _(handle, ticket, dheat, dloss):
    eng = _model(handle).engine
    tot = sum((eng.params[n].numel() + 7) // 8 * 8 for n in eng.trainable_names())
    ref = dheat if dheat is not None else dloss
    return ref.new_empty((tot,), dtype=torch.float32)


def arena_views(engine, arena: Tensor):
    """{trainable parameter name: gradient view of the flat arena} (same layout as Engine.backward / FusedAdam / the DDP buckets)."""
    out, off = {}, 0
    for n in engine.trainable_names():
        p = engine.params[n]
        out[n] = arena[off:off + p.numel()].view(p.shape)
        off += (p.numel() + 7) // 8 * 8
    return out


def _policy_setup(ctx, inputs, output):
    handle, params, need_grad = inputs[0], inputs[6], inputs[8]
    model = _model(handle)
    ectx, model._pending_ectx = model._pending_ectx, None
    ctx.handle, ctx.ectx = handle, ectx       # the saved activations live exactly as long as the autograd graph
    ctx.n_params, ctx.n_targets = len(params), len(inputs[5])
    ctx.trainable = [i for i, p in enumerate(params) if p.requires_grad]
    ctx.has_loss = output[1].numel() > 0
    ctx.set_materialize_grads(False)


def _policy_backward(ctx, dheat, dloss, dkey):
    if ctx.ectx is None:
        raise RuntimeError("backward through a policy forward that ran without grad (or a second backward)")
    dl = dloss if (dloss is not None and ctx.has_loss and dloss.numel()) else None
    _TICKET[0] += 1
    ticket = _TICKET[0]
    _BWD_SLOT[(ctx.handle, ticket)], ctx.ectx = ctx.ectx, None     # ops take no Python objects: hand the context over by ticket
    try:
        arena = torch.ops.bfk.policy_bwd(ctx.handle, ticket, dheat, dl)
    finally:
        _BWD_SLOT.pop((ctx.handle, ticket), None)
    gparams: List[Optional[Tensor]] = [None] * ctx.n_params
    model = _model(ctx.handle)
    by_name = arena_views(model.engine, arena)
    for i in ctx.trainable:
        gparams[i] = by_name.get(model._param_names[i])
    return None, None, None, None, None, [None] * ctx.n_targets, gparams, None, None, None


policy_fwd.register_autograd(_policy_backward, setup_context=_policy_setup)


# =====================================================================================================================
# linear (tcgen05 GEMM, fused bias / GELU epilogue)
# =====================================================================================================================
def _gemm(A, B, D, M, N, Kd, lda, ldb, ldd, trans_a=False, trans_b=False, bias=None, act=K.ACT_NONE, ws=None, split_k=0):
    a = K.GemmArgs(M, N, Kd, A.data_ptr(), lda, int(trans_a), B.data_ptr(), ldb, int(trans_b), _dt(A), D.data_ptr(), ldd, _dt(D),
                   K.ptr(bias), act, None, 0, None, 0, 0, 1.0, 0, K.BACKEND_AUTO, K.ptr(ws), ws.numel() * 4 if ws is not None else 0,
                   split_k)
    K.check(K.load().bfk_gemm(C.byref(a), _stream()), "bfk_gemm")


_ACT = {"none": K.ACT_NONE, "gelu_tanh": K.ACT_GELU_TANH, "gelu_erf": K.ACT_GELU_ERF}


@torch.library.custom_op("bfk::linear", mutates_args=(), device_types="cuda")
def linear(x: Tensor, weight: Tensor, bias: Optional[Tensor], act: str = "none") -> Tensor:
    """y = act(x @ weight^T + bias); x [M,K], weight [N,K] (nn.Linear layout), both bf16 (tcgen05) or both fp32 (SIMT);
    bias fp32."""
    _need_cuda(x, weight, bias)
    x, weight = x.contiguous(), weight.contiguous()
    M, Kd = x.shape
    N = weight.shape[0]
    y = torch.empty(M, N, device=x.device, dtype=x.dtype)
    _gemm(x, weight, y, M, N, Kd, Kd, Kd, N, bias=bias.float().contiguous() if bias is not None else None, act=_ACT[act])
    return y


@linear.register_fake
def _(x, weight, bias, act="none"):
    return x.new_empty((x.shape[0], weight.shape[0]))


@torch.library.custom_op("bfk::linear_bwd", mutates_args=(), device_types="cuda")
def linear_bwd(dy: Tensor, x: Tensor, weight: Tensor, need_bias: bool) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (dx [M,K] in x's dtype, dW [N,K] fp32, db [N] fp32 (empty unless need_bias)) for act == none."""
    dy, x, weight = dy.contiguous(), x.contiguous(), weight.contiguous()
    M, Kd = x.shape
    N = weight.shape[0]
    lib = K.load()
    dx = torch.empty_like(x)
    _gemm(dy, weight, dx, M, Kd, N, N, Kd, Kd, trans_b=True)
    dW = torch.empty(N, Kd, device=x.device, dtype=torch.float32)
    ws = torch.empty(16 * 1024 * 1024, device=x.device, dtype=torch.float32) if x.dtype == torch.bfloat16 else None
    _gemm(dy, x, dW, N, Kd, M, N, Kd, Kd, trans_a=True, trans_b=True, ws=ws)
    db = torch.empty(N if need_bias else 0, device=x.device, dtype=torch.float32)
    if need_bias:
        cws = torch.empty(lib.bfk_colsum_workspace_floats(M, N), device=x.device, dtype=torch.float32)
        K.check(lib.bfk_colsum(dy.data_ptr(), _dt(dy), M, N, N, db.data_ptr(), cws.data_ptr(), _stream()), "bfk_colsum")
    return dx, dW, db


@linear_bwd.register_fake
def _(dy, x, weight, need_bias):
    return torch.empty_like(x), weight.new_empty(weight.shape, dtype=torch.float32), \
        weight.new_empty((weight.shape[0] if need_bias else 0,), dtype=torch.float32)


def _linear_setup(ctx, inputs, output):
    x, weight, bias, act = inputs
    if act != "none":
        raise K.BfkError("bfk::linear: autograd is registered for act='none' (the engine's fused GELU backward is internal)")
    ctx.save_for_backward(x, weight)
    ctx.has_bias = bias is not None


def _linear_backward(ctx, dy):
    x, weight = ctx.saved_tensors
    dx, dW, db = torch.ops.bfk.linear_bwd(dy.to(x.dtype), x, weight, ctx.has_bias)
    return dx, dW.to(weight.dtype), (db if ctx.has_bias else None), None


linear.register_autograd(_linear_backward, setup_context=_linear_setup)


# =====================================================================================================================
# LayerNorm
# =====================================================================================================================
@torch.library.custom_op("bfk::layernorm", mutates_args=(), device_types="cuda")
def layernorm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, out_bf16: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """x fp32 [M,D] -> (y [M,D] fp32|bf16, mean [M], rstd [M])."""
    _need_cuda(x, gamma, beta)
    x = x.float().contiguous()
    M, D = x.shape
    y = torch.empty(M, D, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    mean = torch.empty(M, device=x.device, dtype=torch.float32)
    rstd = torch.empty_like(mean)
    a = K.LayerNormFwdArgs(M, D, x.data_ptr(), D, gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(), D, _dt(y), mean.data_ptr(),
                           rstd.data_ptr(), None, 0, 0.0, 0, None, None, 0, None)
    K.check(K.load().bfk_layernorm_fwd(C.byref(a), _stream()), "bfk_layernorm_fwd")
    return y, mean, rstd


@layernorm.register_fake
def _(x, gamma, beta, eps, out_bf16=False):
    M, D = x.shape
    return (x.new_empty((M, D), dtype=torch.bfloat16 if out_bf16 else torch.float32), x.new_empty((M,), dtype=torch.float32),
            x.new_empty((M,), dtype=torch.float32))


@torch.library.custom_op("bfk::layernorm_bwd", mutates_args=(), device_types="cuda")
def layernorm_bwd(dy: Tensor, x: Tensor, mean: Tensor, rstd: Tensor, gamma: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
    """-> (dx [M,D] fp32, dgamma [D], dbeta [D])."""
    dy, x = dy.contiguous(), x.float().contiguous()
    M, D = x.shape
    lib = K.load()
    dx = torch.empty_like(x)
    dg, db = torch.empty(D, device=x.device, dtype=torch.float32), torch.empty(D, device=x.device, dtype=torch.float32)
    ws = torch.empty(lib.bfk_layernorm_bwd_workspace_floats(M, D), device=x.device, dtype=torch.float32)
    a = K.LayerNormBwdArgs(M, D, dy.data_ptr(), D, _dt(dy), x.data_ptr(), D, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                           None, D, dx.data_ptr(), D, None, D, 0, dg.data_ptr(), db.data_ptr(), ws.data_ptr(), None, 0, 0.0, 0, None)
    K.check(lib.bfk_layernorm_bwd(C.byref(a), _stream()), "bfk_layernorm_bwd")
    return dx, dg, db


@layernorm_bwd.register_fake
def _(dy, x, mean, rstd, gamma):
    return x.new_empty(x.shape, dtype=torch.float32), gamma.new_empty(gamma.shape, dtype=torch.float32), \
        gamma.new_empty(gamma.shape, dtype=torch.float32)


def _ln_setup(ctx, inputs, output):
    x, gamma, beta, eps, out_bf16 = inputs
    ctx.save_for_backward(x, output[1], output[2], gamma)
    ctx.mark_non_differentiable(output[1], output[2])


def _ln_backward(ctx, dy, dmean, drstd):
    x, mean, rstd, gamma = ctx.saved_tensors
    dx, dg, db = torch.ops.bfk.layernorm_bwd(dy, x, mean, rstd, gamma)
    return dx, dg, db, None, None


layernorm.register_autograd(_ln_backward, setup_context=_ln_setup)


# =====================================================================================================================
# attention on the packed QKV matrix
# =====================================================================================================================
@torch.library.custom_op("bfk::attention", mutates_args=(), device_types="cuda")
def attention(qkv: Tensor, row_mask: Optional[Tensor], B: int, N: int, H: int, hd: int, scale: float) -> Tuple[Tensor, Tensor]:
    """qkv [B*N, 3*H*hd] = [q(h,d) | k(h,d) | v(h,d)] (bf16: tcgen05, fp32: SIMT) -> (out [B*N, H*hd], lse [B*H*N] fp32).
    ``row_mask`` fp32 [B*N]: rows with 0 get every score filled with -1e5, i.e. uniform attention over ALL keys — the
    reference's query-row mask (bifold/models/transformer.py:33-35), not a key mask."""
    _need_cuda(qkv, row_mask)
    qkv = qkv.contiguous()
    out = torch.empty(B * N, H * hd, device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(B * H * N, device=qkv.device, dtype=torch.float32)
    a = K.AttentionArgs(B, N, H, hd, qkv.data_ptr(), 3 * H * hd, out.data_ptr(), H * hd, lse.data_ptr(), K.ptr(row_mask), scale,
                        -100000.0, _dt(qkv), K.BACKEND_AUTO, None, 0, None, 0, None)
    K.check(K.load().bfk_attention_fwd(C.byref(a), _stream()), "bfk_attention_fwd")
    return out, lse


@attention.register_fake
def _(qkv, row_mask, B, N, H, hd, scale):
    return qkv.new_empty((B * N, H * hd)), qkv.new_empty((B * H * N,), dtype=torch.float32)


@torch.library.custom_op("bfk::attention_bwd", mutates_args=(), device_types="cuda")
def attention_bwd(qkv: Tensor, out: Tensor, lse: Tensor, dout: Tensor, row_mask: Optional[Tensor], B: int, N: int, H: int,
                  hd: int, scale: float) -> Tensor:
    qkv, out, dout = qkv.contiguous(), out.contiguous(), dout.contiguous()
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B * H * N, device=qkv.device, dtype=torch.float32)
    a = K.AttentionArgs(B, N, H, hd, qkv.data_ptr(), 3 * H * hd, out.data_ptr(), H * hd, lse.data_ptr(), K.ptr(row_mask), scale,
                        -100000.0, _dt(qkv), K.BACKEND_AUTO, dout.data_ptr(), H * hd, dqkv.data_ptr(), 3 * H * hd, delta.data_ptr())
    K.check(K.load().bfk_attention_bwd(C.byref(a), _stream()), "bfk_attention_bwd")
    return dqkv


@attention_bwd.register_fake
def _(qkv, out, lse, dout, row_mask, B, N, H, hd, scale):
    return torch.empty_like(qkv)


def _attn_setup(ctx, inputs, output):
    qkv, row_mask, B, N, H, hd, scale = inputs
    ctx.save_for_backward(qkv, output[0], output[1], row_mask if row_mask is not None else qkv.new_zeros(0))
    ctx.dims = (B, N, H, hd, scale, row_mask is not None)
    ctx.mark_non_differentiable(output[1])


def _attn_backward(ctx, dout, dlse):
    qkv, out, lse, rm = ctx.saved_tensors
    B, N, H, hd, scale, has_mask = ctx.dims
    dqkv = torch.ops.bfk.attention_bwd(qkv, out, lse, dout.to(qkv.dtype), rm if has_mask else None, B, N, H, hd, scale)
    return dqkv, None, None, None, None, None, None


attention.register_autograd(_attn_backward, setup_context=_attn_setup)


# =====================================================================================================================
# loss / argmax / action
# =====================================================================================================================
@torch.library.custom_op("bfk::bce_prob", mutates_args=(), device_types="cuda")
def bce_prob(p: Tensor, t: Tensor, t_mask: Optional[Tensor]) -> Tensor:
    """nn.BCELoss()(p, t [* t_mask]): mean over all elements, log clamped at -100 (bifold/losses/__init__.py:61,76-83)."""
    _need_cuda(p, t, t_mask)
    p, t = p.contiguous(), t.to(torch.float32).contiguous()
    tm = t_mask.to(torch.float32).contiguous() if t_mask is not None else None
    loss = torch.empty(1, device=p.device, dtype=torch.float32)
    part = torch.empty(1024, device=p.device, dtype=torch.float32)
    K.check(K.load().bfk_bce_prob(p.data_ptr(), t.data_ptr(), K.ptr(tm), p.numel(), loss.data_ptr(), part.data_ptr(), _stream()),
            "bfk_bce_prob")
    return loss.reshape(())


@bce_prob.register_fake
def _(p, t, t_mask):
    return p.new_empty((), dtype=torch.float32)


@torch.library.custom_op("bfk::bce_prob_bwd", mutates_args=(), device_types="cuda")
def bce_prob_bwd(g: Tensor, p: Tensor, t: Tensor, t_mask: Optional[Tensor]) -> Tensor:
    p, t = p.contiguous(), t.to(torch.float32).contiguous()
    tm = t_mask.to(torch.float32).contiguous() if t_mask is not None else None
    gs = (g.reshape(1).float() / p.numel()).contiguous()
    dp = torch.empty_like(p)
    K.check(K.load().bfk_bce_prob_bwd(p.data_ptr(), t.data_ptr(), K.ptr(tm), p.numel(), gs.data_ptr(), dp.data_ptr(), _stream()),
            "bfk_bce_prob_bwd")
    return dp


@bce_prob_bwd.register_fake
def _(g, p, t, t_mask):
    return torch.empty_like(p)


def _bce_setup(ctx, inputs, output):
    p, t, tm = inputs
    ctx.save_for_backward(p, t, tm if tm is not None else p.new_zeros(0))
    ctx.has_mask = tm is not None


def _bce_backward(ctx, g):
    p, t, tm = ctx.saved_tensors
    return torch.ops.bfk.bce_prob_bwd(g, p, t, tm if ctx.has_mask else None), None, None


bce_prob.register_autograd(_bce_backward, setup_context=_bce_setup)


# ---------------------------------------------------------------------------------------------------------------------
# mask losses on mask_heatmap (compute_mask=True): bifold/losses/__init__.py:126-155
# ---------------------------------------------------------------------------------------------------------------------
@torch.library.custom_op("bfk::dice_loss", mutates_args=(), device_types="cuda")
def dice_loss(p: Tensor, t: Tensor) -> Tuple[Tensor, Tensor]:
    """DiceLoss: p, t [B, ...] -> (sum_b 1 - (2 sum(p t) + 1) / (sum p + sum t + 1), per-sample sums [B,3] for the backward)."""
    _need_cuda(p, t)
    B = p.shape[0]
    p2, t2 = p.reshape(B, -1).float().contiguous(), t.reshape(B, -1).to(torch.float32).contiguous()
    lib = K.load()
    sums = torch.empty(B, 3, device=p.device, dtype=torch.float32)
    loss = torch.empty(1, device=p.device, dtype=torch.float32)
    part = torch.empty(lib.bfk_dice_loss_workspace_floats(B), device=p.device, dtype=torch.float32)
    K.check(lib.bfk_dice_loss_fwd(p2.data_ptr(), t2.data_ptr(), B, p2.shape[1], sums.data_ptr(), loss.data_ptr(), part.data_ptr(), _stream()),
            "bfk_dice_loss_fwd")
    return loss.reshape(()), sums


@dice_loss.register_fake
def _(p, t):
    return p.new_empty((), dtype=torch.float32), p.new_empty((p.shape[0], 3), dtype=torch.float32)


@torch.library.custom_op("bfk::dice_loss_bwd", mutates_args=(), device_types="cuda")
def dice_loss_bwd(g: Tensor, t: Tensor, sums: Tensor) -> Tensor:
    B = t.shape[0]
    t2 = t.reshape(B, -1).to(torch.float32).contiguous()
    dp = torch.empty_like(t2)
    K.check(K.load().bfk_dice_loss_bwd(t2.data_ptr(), B, t2.shape[1], sums.data_ptr(), g.reshape(1).float().contiguous().data_ptr(),
                                       dp.data_ptr(), _stream()), "bfk_dice_loss_bwd")
    return dp


@dice_loss_bwd.register_fake
def _(g, t, sums):
    return t.new_empty((t.shape[0], t[0].numel()), dtype=torch.float32)


def _dice_setup(ctx, inputs, output):
    p, t = inputs
    ctx.save_for_backward(t, output[1])
    ctx.p_shape = p.shape
    ctx.mark_non_differentiable(output[1])


def _dice_backward(ctx, g, gsums):
    t, sums = ctx.saved_tensors
    return torch.ops.bfk.dice_loss_bwd(g, t, sums).reshape(ctx.p_shape), None


dice_loss.register_autograd(_dice_backward, setup_context=_dice_setup)


@torch.library.custom_op("bfk::focal_loss", mutates_args=(), device_types="cuda")
def focal_loss(p: Tensor, t: Tensor, alpha: float, gamma: float) -> Tensor:
    """FocalLoss: alpha_t * BCE(p, t) * (1 - p_t)^gamma, ``.mean(1).sum()`` over p [B, H, W] (= sum / H)."""
    _need_cuda(p, t)
    p2, t2 = p.float().contiguous(), t.to(torch.float32).contiguous()
    loss = torch.empty(1, device=p.device, dtype=torch.float32)
    part = torch.empty(1024, device=p.device, dtype=torch.float32)
    K.check(K.load().bfk_focal_loss_fwd(p2.data_ptr(), t2.data_ptr(), p2.numel(), float(alpha), float(gamma), 1.0 / p.shape[1],
                                        loss.data_ptr(), part.data_ptr(), _stream()), "bfk_focal_loss_fwd")
    return loss.reshape(())


@focal_loss.register_fake
def _(p, t, alpha, gamma):
    return p.new_empty((), dtype=torch.float32)


@torch.library.custom_op("bfk::focal_loss_bwd", mutates_args=(), device_types="cuda")
def focal_loss_bwd(g: Tensor, p: Tensor, t: Tensor, alpha: float, gamma: float) -> Tensor:
    p2, t2 = p.float().contiguous(), t.to(torch.float32).contiguous()
    dp = torch.empty_like(p2)
    K.check(K.load().bfk_focal_loss_bwd(p2.data_ptr(), t2.data_ptr(), p2.numel(), float(alpha), float(gamma), 1.0 / p.shape[1],
                                        g.reshape(1).float().contiguous().data_ptr(), dp.data_ptr(), _stream()), "bfk_focal_loss_bwd")
    return dp


@focal_loss_bwd.register_fake
def _(g, p, t, alpha, gamma):
    return torch.empty_like(p, dtype=torch.float32)


def _focal_setup(ctx, inputs, output):
    p, t, alpha, gamma = inputs
    ctx.save_for_backward(p, t)
    ctx.ag = (alpha, gamma)


def _focal_backward(ctx, g):
    p, t = ctx.saved_tensors
    return torch.ops.bfk.focal_loss_bwd(g, p, t, ctx.ag[0], ctx.ag[1]), None, None, None


focal_loss.register_autograd(_focal_backward, setup_context=_focal_setup)


@torch.library.custom_op("bfk::argmax2d", mutates_args=(), device_types="cuda")
def argmax2d(heat: Tensor) -> Tensor:
    """heat [n, H, W] fp32 -> packed keys [n] int64 = (ordered float bits << 32) | ~index: the maximum key is the FIRST
    maximal row-major index (np.argmax tie rule, bifold/models/utils.py:66-73).  ``decode_argmax`` unpacks."""
    _need_cuda(heat)
    heat = heat.float().contiguous()
    n = heat.shape[0]
    key = torch.empty(n, device=heat.device, dtype=torch.int64)
    K.check(K.load().bfk_argmax2d(heat.data_ptr(), n, heat[0].numel(), key.data_ptr(), _stream()), "bfk_argmax2d")
    return key


@argmax2d.register_fake
def _(heat):
    return heat.new_empty((heat.shape[0],), dtype=torch.int64)


def decode_argmax(key: Tensor, W: int) -> Tensor:
    """keys -> [n, 2] int64 pixel coordinates [x(col), y(row)] like the reference."""
    idx = 0xFFFFFFFF - (key & 0xFFFFFFFF)
    return torch.stack((idx % W, idx // W), dim=1)


@torch.library.custom_op("bfk::action", mutates_args=(), device_types="cuda")
def action(heat: Tensor, key: Tensor, mask: Optional[Tensor], pick_heads: Sequence[int], place_heads: Sequence[int], n_arms: int,
           threshold: float) -> Tuple[Tensor, Tensor]:
    """heat [heads,B,H,W], argmax keys [heads*B] -> (xy [heads,B,2] int32 = [x,y] or [-1,-1] for a dropped arm, conf [heads,B]):
    pick heads snapped to the nearest mask pixel, arm kept iff conf >= threshold or it is the more confident arm
    (bifold/models/__init__.py:108-141, utils.py:80-83)."""
    _need_cuda(heat, key, mask)
    heads, B, H, W = heat.shape
    xy = torch.empty(heads, B, 2, device=heat.device, dtype=torch.int32)
    conf = torch.empty(heads, B, device=heat.device, dtype=torch.float32)
    pk, pl = (list(pick_heads) + [0, 0])[:2], (list(place_heads) + [0, 0])[:2]
    a = K.ActionArgs(B, heads, H, W, n_arms, heat.data_ptr(), key.data_ptr(), K.ptr(mask), (C.c_int32 * 2)(*pk),
                     (C.c_int32 * 2)(*pl), float(threshold), xy.data_ptr(), conf.data_ptr())
    K.check(K.load().bfk_action(C.byref(a), _stream()), "bfk_action")
    return xy, conf


@action.register_fake
def _(heat, key, mask, pick_heads, place_heads, n_arms, threshold):
    heads, B = heat.shape[0], heat.shape[1]
    return heat.new_empty((heads, B, 2), dtype=torch.int32), heat.new_empty((heads, B), dtype=torch.float32)


# =====================================================================================================================
# GPU input stage
# =====================================================================================================================
NO_ACTION = -2 ** 31      # centre x of an arm without an action: all-zero target map (bifold/data/processor.py:296)


@torch.library.custom_op("bfk::gauss_targets", mutates_args=(), device_types="cuda")
def gauss_targets(centers: Tensor, H: int, W: int, sigma: float, points_axis: bool = False) -> Tensor:
    """Processor._make_gaussmap on device (bifold/data/processor.py:405-425).  centers int32 [..., 2] = [cx, cy] -> fp32
    [..., H, W] ("first" strategy); with ``points_axis`` centers is [..., P, 2] and maps with several valid points use the
    "gmm" strategy (sum of Gaussians normalised by the map maximum).  cx == NO_ACTION marks padding; a map without a valid
    point is all zeros."""
    _need_cuda(centers)
    c = centers.to(torch.int32).contiguous()
    P = c.shape[-2] if points_axis else 1
    lead = c.shape[:-2] if points_axis else c.shape[:-1]
    out = torch.empty(*lead, H, W, device=c.device, dtype=torch.float32)
    n = c.numel() // (2 * P)
    K.check(K.load().bfk_gauss_targets(c.data_ptr(), n, P, H, W, float(sigma), out.data_ptr(), _stream()), "bfk_gauss_targets")
    return out


@gauss_targets.register_fake
def _(centers, H, W, sigma, points_axis=False):
    lead = centers.shape[:-2] if points_axis else centers.shape[:-1]
    return centers.new_empty((*lead, H, W), dtype=torch.float32)


@torch.library.custom_op("bfk::normalize_u8", mutates_args=(), device_types="cuda")
def normalize_u8(frames: Tensor, mean: float, std: float) -> Tensor:
    """uint8 [..., H, W, 3] frames at model resolution -> fp32 [..., 3, H, W] = (x/255 - mean)/std (SigLIP image processor,
    bifold/data/processor.py:143-148)."""
    _need_cuda(frames)
    f = frames.contiguous()
    *lead, H, W, ch = f.shape
    if ch != 3 or f.dtype != torch.uint8:
        raise K.BfkError("bfk::normalize_u8 takes uint8 [..., H, W, 3] frames")
    out = torch.empty(*lead, 3, H, W, device=f.device, dtype=torch.float32)
    n = f.numel() // (3 * H * W)
    K.check(K.load().bfk_normalize_u8(f.data_ptr(), n, H, W, float(mean), float(std), out.data_ptr(), _stream()), "bfk_normalize_u8")
    return out


@normalize_u8.register_fake
def _(frames, mean, std):
    *lead, H, W, ch = frames.shape
    return frames.new_empty((*lead, 3, H, W), dtype=torch.float32)
