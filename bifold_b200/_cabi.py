"""ctypes binding of include/bifold_b200.h (the C ABI of libbifold_b200.so).

Fails loudly when the library is missing: there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BFK_LIB_PATH") or os.path.join(_HERE, "lib", "libbifold_b200.so")   # BFK_LIB_PATH: A/B builds (BFK_BUILD_TAG of bifold_b200/build.py, tools/gpu_ab.sh)

F32, BF16 = 0, 1
BACKEND_AUTO, BACKEND_SIMT, BACKEND_TCGEN05, BACKEND_TCGEN05_PERSISTENT = 0, 1, 2, 3
ACT_NONE, ACT_GELU_TANH, ACT_GELU_ERF, ACT_DGELU_TANH, ACT_DGELU_ERF = 0, 1, 2, 3, 4
ACT_GELU_TANH_GRAD, ACT_GELU_ERF_GRAD, ACT_MUL_AUX = 5, 6, 7

vp, i32, i64, f32, u32, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_uint32, C.c_uint64


class GemmArgs(C.Structure):
    _fields_ = [("M", i32), ("N", i32), ("K", i32),
                ("A", vp), ("lda", i64), ("trans_a", i32),
                ("B", vp), ("ldb", i64), ("trans_b", i32),
                ("in_dtype", i32),
                ("D", vp), ("ldd", i64), ("out_dtype", i32),
                ("bias", vp), ("act", i32),
                ("aux", vp), ("ld_aux", i64),
                ("residual", vp), ("ldr", i64), ("res_row_mod", i32),
                ("alpha", f32), ("accumulate", i32), ("backend", i32),
                ("workspace", vp), ("workspace_bytes", i64), ("split_k", i32)]


class LayerNormFwdArgs(C.Structure):
    _fields_ = [("M", i32), ("D", i32), ("x", vp), ("ldx", i64), ("gamma", vp), ("beta", vp), ("eps", f32),
                ("y", vp), ("ldy", i64), ("y_dtype", i32), ("mean", vp), ("rstd", vp),
                ("y2", vp), ("ldy2", i64), ("y2_dropout_p", f32), ("y2_seed", u64),
                ("lora_a", vp), ("lora_u", vp), ("ld_u", i64), ("seed_ptr", vp)]


class LayerNormBwdArgs(C.Structure):
    _fields_ = [("M", i32), ("D", i32), ("dy", vp), ("ld_dy", i64), ("dy_dtype", i32),
                ("x", vp), ("ldx", i64), ("mean", vp), ("rstd", vp), ("gamma", vp),
                ("dres", vp), ("ld_dres", i64), ("dx", vp), ("ld_dx", i64),
                ("dx2", vp), ("ld_dx2", i64), ("dx2_dtype", i32),
                ("dgamma", vp), ("dbeta", vp), ("partials", vp),
                ("dy2", vp), ("ld_dy2", i64), ("dy2_dropout_p", f32), ("dy2_seed", u64), ("seed_ptr", vp)]


class AttentionArgs(C.Structure):
    _fields_ = [("B", i32), ("N", i32), ("H", i32), ("hd", i32),
                ("qkv", vp), ("ld_qkv", i64), ("out", vp), ("ld_out", i64),
                ("lse", vp), ("row_mask", vp), ("scale", f32), ("mask_fill", f32),
                ("dtype", i32), ("backend", i32),
                ("dout", vp), ("ld_dout", i64), ("dqkv", vp), ("ld_dqkv", i64), ("delta", vp)]


class CrossAttentionArgs(C.Structure):
    _fields_ = [("B", i32), ("Nf", i32), ("H", i32), ("hd", i32), ("q0", i32), ("nq", i32), ("k0", i32), ("nk", i32),
                ("dtype", i32), ("qkv", vp), ("ld_qkv", i64), ("scale", f32), ("key_bias", vp),
                ("out", vp), ("ld_out", i64), ("lse", vp), ("dout", vp), ("ld_dout", i64), ("dqkv", vp), ("ld_dqkv", i64),
                ("delta", vp)]


class AssembleArgs(C.Structure):
    _fields_ = [("B", i32), ("L", i32), ("Np", i32), ("T", i32), ("D", i32),
                ("txt", vp), ("img", vp), ("ctx", vp), ("text_token", vp), ("image_token", vp),
                ("type_emb", vp), ("ctx_pos", vp), ("ctx_mask", vp), ("X", vp), ("mask_out", vp)]


class AssembleBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("L", i32), ("Np", i32), ("T", i32), ("D", i32),
                ("dX", vp), ("dtxt", vp), ("dimg", vp), ("dctx", vp),
                ("d_text_token", vp), ("d_image_token", vp), ("d_type_emb", vp), ("d_ctx_pos", vp),
                ("partials", vp)]


class DecoderCollapseArgs(C.Structure):
    _fields_ = [("heads", i32), ("D", i32), ("W", vp), ("b", vp), ("w_eff", vp), ("b_eff", vp), ("chain", vp),
                ("d_w_eff", vp), ("d_b_eff", vp), ("dW", vp), ("db", vp)]


class HeadProjectArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("Np", i32), ("D", i32),
                ("X", vp), ("x_batch_stride", i64), ("row0", i32),
                ("w_eff", vp), ("b_eff", vp), ("grid", vp),
                ("dgrid", vp), ("Nf", i32), ("dX", vp), ("dX2", vp), ("dX2_dtype", i32),
                ("d_w_eff", vp), ("d_b_eff", vp), ("partials", vp)]


class HeatmapFwdArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("g", i32),
                ("grid", vp), ("tap_start", vp), ("tap_w", vp), ("heat", vp),
                ("target", vp), ("target_mask", vp), ("mask_heads", u32),
                ("loss", vp), ("argmax_key", vp), ("partials", vp), ("counters", vp)]


class HeatmapBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("g", i32), ("mode", i32),
                ("heat", vp), ("target", vp), ("target_mask", vp), ("mask_heads", u32),
                ("gscale", vp), ("gscale_mul", f32), ("gscale_has_total", i32),
                ("dheat", vp), ("tap_start", vp), ("tap_w", vp), ("dgrid", vp), ("partials", vp), ("counters", vp)]


class ActionArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("H", i32), ("W", i32), ("n_arms", i32),
                ("heat", vp), ("key", vp), ("mask", vp),
                ("pick_heads", i32 * 2), ("place_heads", i32 * 2), ("threshold", f32),
                ("out_xy", vp), ("out_conf", vp)]


class MaskGateArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("H", i32), ("W", i32), ("pick_heads", u32),
                ("heat_raw", vp), ("mask_p", vp), ("heat_out", vp)]


class MaskGateBwdArgs(C.Structure):
    _fields_ = [("B", i32), ("heads", i32), ("H", i32), ("W", i32), ("pick_heads", u32), ("detach_mask", i32),
                ("dheat_out", vp), ("has_dmask", i32), ("heat_raw", vp), ("mask_p", vp), ("dheat_raw", vp), ("dmask", vp)]


class LoraGradsArgs(C.Structure):
    _fields_ = [("M", i32), ("D", i32), ("r", i32),
                ("dqkv", vp), ("ld_dqkv", i64), ("u", vp), ("ld_u", i64), ("du", vp), ("ld_du", i64), ("x", vp), ("ld_x", i64),
                ("scale", f32), ("dBq", vp), ("dBv", vp), ("dAq", vp), ("dAv", vp),
                ("workspace", vp), ("workspace_bytes", i64)]


class CastSegment(C.Structure):
    _fields_ = [("src", vp), ("dst", vp), ("ld_dst", i64), ("first_chunk", i64),
                ("rows", i32), ("cols", i32), ("dst_dtype", i32), ("scale", f32)]


class AdamWArgs(C.Structure):
    _fields_ = [("n", i64), ("p", vp), ("g", vp), ("m", vp), ("v", vp), ("shadow_bf16", vp),
                ("lr", f32), ("beta1", f32), ("beta2", f32), ("eps", f32), ("weight_decay", f32),
                ("decoupled", i32), ("step", i32), ("grad_scale", f32), ("dyn", vp)]


# name -> (restype, argtypes).  This table is also what tests/test_cabi_symbols.py checks against the header.
_PROTOS = {
    "bfk_version": (i32, []),
    "bfk_last_error": (C.c_char_p, []),
    "bfk_device_sm_count": (i32, []),
    "bfk_gemm": (i32, [C.POINTER(GemmArgs), vp]),
    "bfk_colsum_workspace_floats": (C.c_size_t, [i32, i32]),
    "bfk_colsum": (i32, [vp, i32, i32, i32, i64, vp, vp, vp]),
    "bfk_layernorm_fwd": (i32, [C.POINTER(LayerNormFwdArgs), vp]),
    "bfk_layernorm_bwd_workspace_floats": (C.c_size_t, [i32, i32]),
    "bfk_layernorm_bwd": (i32, [C.POINTER(LayerNormBwdArgs), vp]),
    "bfk_attention_fwd": (i32, [C.POINTER(AttentionArgs), vp]),
    "bfk_attention_bwd": (i32, [C.POINTER(AttentionArgs), vp]),
    "bfk_cross_attention_fwd": (i32, [C.POINTER(CrossAttentionArgs), vp]),
    "bfk_cross_attention_bwd": (i32, [C.POINTER(CrossAttentionArgs), vp]),
    "bfk_cross_attention_weights": (i32, [C.POINTER(CrossAttentionArgs), vp, vp]),
    "bfk_im2col": (i32, [vp, i32, i32, i32, i32, i32, vp, i64, i32, vp]),
    "bfk_text_embed": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "bfk_lora_down": (i32, [vp, i64, i32, i32, i32, vp, f32, u64, vp]),
    "bfk_lora_down_bwd": (i32, [vp, i64, i32, i32, i32, vp, f32, u64, vp]),
    "bfk_dropout_copy": (i32, [vp, i64, vp, i64, i32, i32, i32, f32, u64, vp]),
    "bfk_adamw_advance": (i32, [vp, vp, f32, f32, vp, vp]),
    "bfk_multi_cast": (i32, [vp, i32, i64, vp]),
    "bfk_lora_grads_workspace_bytes": (C.c_size_t, [i32]),
    "bfk_lora_grads": (i32, [C.POINTER(LoraGradsArgs), vp]),
    "bfk_skinny_gram_workspace_floats": (C.c_size_t, [i32, i32, i32]),
    "bfk_skinny_gram": (i32, [vp, i64, i32, vp, i64, i32, i32, i32, f32, vp, i64, i32, f32, u64, vp, vp]),
    "bfk_assemble_fwd": (i32, [C.POINTER(AssembleArgs), vp]),
    "bfk_assemble_bwd": (i32, [C.POINTER(AssembleBwdArgs), vp]),
    "bfk_decoder_collapse_fwd": (i32, [C.POINTER(DecoderCollapseArgs), vp]),
    "bfk_decoder_collapse_bwd": (i32, [C.POINTER(DecoderCollapseArgs), vp]),
    "bfk_head_project_fwd": (i32, [C.POINTER(HeadProjectArgs), vp]),
    "bfk_head_project_bwd_workspace_floats": (C.c_size_t, [i32, i32, i32, i32]),
    "bfk_head_project_bwd": (i32, [C.POINTER(HeadProjectArgs), vp]),
    "bfk_upsample16_taps": (i32, [i32, vp, vp]),
    "bfk_heatmap_bands": (i32, [i32]),
    "bfk_heatmap_fwd_workspace_floats": (C.c_size_t, [i32, i32, i32]),
    "bfk_heatmap_fwd": (i32, [C.POINTER(HeatmapFwdArgs), vp]),
    "bfk_heatmap_bwd_workspace_floats": (C.c_size_t, [i32, i32, i32]),
    "bfk_heatmap_bwd": (i32, [C.POINTER(HeatmapBwdArgs), vp]),
    "bfk_bce_prob": (i32, [vp, vp, vp, i64, vp, vp, vp]),
    "bfk_bce_prob_bwd": (i32, [vp, vp, vp, i64, vp, vp, vp]),
    "bfk_argmax2d": (i32, [vp, i32, i32, vp, vp]),
    "bfk_action": (i32, [C.POINTER(ActionArgs), vp]),
    "bfk_adamw": (i32, [C.POINTER(AdamWArgs), vp]),
    "bfk_gauss_targets": (i32, [vp, i32, i32, i32, i32, f32, vp, vp]),
    "bfk_normalize_u8": (i32, [vp, i64, i32, i32, f32, f32, vp, vp]),
    "bfk_mask_gate_fwd": (i32, [C.POINTER(MaskGateArgs), vp]),
    "bfk_mask_gate_bwd": (i32, [C.POINTER(MaskGateBwdArgs), vp]),
    "bfk_add_inplace": (i32, [vp, vp, i64, vp]),
    "bfk_dice_loss_workspace_floats": (C.c_size_t, [i32]),
    "bfk_dice_loss_fwd": (i32, [vp, vp, i32, i64, vp, vp, vp, vp]),
    "bfk_dice_loss_bwd": (i32, [vp, i32, i64, vp, vp, vp, vp]),
    "bfk_focal_loss_fwd": (i32, [vp, vp, i64, f32, f32, f32, vp, vp, vp]),
    "bfk_focal_loss_bwd": (i32, [vp, vp, i64, f32, f32, f32, vp, vp, vp]),
    "bfk_cast_bf16": (i32, [vp, vp, i64, vp]),
    "bfk_cast_f32_to": (i32, [vp, vp, i32, i64, i64, i64, i64, vp]),
}

_lib = None


class BfkError(RuntimeError):
    pass


def load():
    """Load libbifold_b200.so (built by bifold_b200.build).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BfkError(
            f"{LIB_PATH} is missing: build it with `python -m bifold_b200.build` "
            "(bifold_b200 has no CPU or eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().bfk_last_error().decode("utf-8", "replace")
        raise BfkError(f"{what or 'bfk call'} failed (rc={rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (or None)."""
    return None if t is None else t.data_ptr()
