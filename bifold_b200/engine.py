"""Policy executor: chains the C-ABI kernels of libbifold_b200.so into the BiFold forward and an
explicit backward.  PyTorch is used for device memory and streams only — every arithmetic step is
one of our CUDA kernels (no eager/CPU fallback).

Reference call stack being replaced: SigLip/SiglipSequential.forward
(bifold/models/siglip.py:56-137) -> HF SiglipModel towers -> PickPlaceConvDecoder.forward
(bifold/models/pickplace.py:59-124) -> ConcatTransformer/Transformer (fusion.py:35-64,
transformer.py:26-64) -> ConvDecoder x heads (modules.py:33-74) -> sigmoid.

Data layout in HBM
  * residual streams x: fp32 [rows, D] (rows = samples*tokens, token-major within a sample)
  * GEMM operands / activations: T = bf16 ("bf16" precision, tcgen05 GEMMs) or fp32 ("fp32" parity mode)
  * packed QKV: T [rows, 3D] = [q(h,d) | k(h,d) | v(h,d)];  LoRA rides as 16 extra K columns of the
    LayerNorm output / QKV weight ([W | (alpha/r) B]) so q,v get their rank-r update inside the QKV GEMM
  * heatmaps: fp32 [heads, B, H, W]
"""
from __future__ import annotations

import os

import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from . import _cabi as K
from .config import PolicyConfig


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return K.F32
    if t.dtype == torch.bfloat16:
        return K.BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class Kernels:
    """Typed wrappers over the C ABI working on torch CUDA tensors, launched on the current stream."""

    def __init__(self):
        self.lib = K.load()
        self.launches = 0          # C-ABI calls issued (each is >= 1 kernel launch); bench.py reports it
        # algorithmic work issued so far, by roofline class (bench.py divides it by the measured per-class device time)
        self.work = {"gemm_flops": 0.0, "attn_flops": 0.0, "ln_bytes": 0.0, "tail_bytes": 0.0}
        self.backend = K.BACKEND_AUTO
        self._ws = None
        self.seed_dev = None   # device int64 counter added to every dropout seed (advanced inside captured CUDA graphs)

    def splitk_ws(self, device):
        """fp32 scratch for split-K partial tiles (wgrad GEMMs); one buffer, reused stream-ordered."""
        key = (device, torch.cuda.current_stream().cuda_stream)   # concurrent streams must not share partial tiles
        if self._ws is None:
            self._ws = {}
        if key not in self._ws:
            self._ws[key] = torch.empty(16 * 1024 * 1024, device=device, dtype=torch.float32)   # 64 MB
        return self._ws[key]

    # -- GEMM ------------------------------------------------------------------------------------
    def lora_ws(self, device, D):
        key = ("lora", device, torch.cuda.current_stream().cuda_stream)
        if self._ws is None:
            self._ws = {}
        if key not in self._ws:
            self._ws[key] = torch.empty(self.lib.bfk_lora_grads_workspace_bytes(D) // 4, device=device, dtype=torch.float32)
        return self._ws[key]

    def gemm(self, A, B, D, M, N, Kd, lda, ldb, ldd, trans_a=False, trans_b=False, bias=None, act=K.ACT_NONE,
             aux=None, ld_aux=0, residual=None, ldr=0, res_row_mod=0, alpha=1.0, accumulate=False, backend=None,
             split_k=0):
        ws = self.splitk_ws(A.device) if (trans_a and A.dtype == torch.bfloat16) or split_k > 1 else None
        a = K.GemmArgs(M, N, Kd, A.data_ptr(), lda, int(trans_a), B.data_ptr(), ldb, int(trans_b), _dt(A),
                       D.data_ptr(), ldd, _dt(D), K.ptr(bias), act, K.ptr(aux), ld_aux, K.ptr(residual), ldr,
                       res_row_mod, alpha, int(accumulate), self.backend if backend is None else backend,
                       K.ptr(ws), ws.numel() * 4 if ws is not None else 0, split_k)
        assert _dt(A) == _dt(B)
        K.check(self.lib.bfk_gemm(C.byref(a), _stream()), "bfk_gemm")
        self.launches += 1
        self.work["gemm_flops"] += 2.0 * M * N * Kd

    def colsum(self, X, M, N, ld, out):
        ws = torch.empty(self.lib.bfk_colsum_workspace_floats(M, N), device=X.device, dtype=torch.float32)
        K.check(self.lib.bfk_colsum(X.data_ptr(), _dt(X), M, N, ld, out.data_ptr(), ws.data_ptr(), _stream()), "bfk_colsum")
        self.launches += 2

    # -- LayerNorm ---------------------------------------------------------------------------------
    def layernorm_fwd(self, x, M, D, gamma, beta, eps, y, ldy, mean=None, rstd=None, y2=None, ldy2=0, y2_p=0.0, y2_seed=0,
                      lora_a=None, lora_u=None, ld_u=0):
        a = K.LayerNormFwdArgs(M, D, x.data_ptr(), D, gamma.data_ptr(), beta.data_ptr(), eps, y.data_ptr(), ldy, _dt(y),
                               K.ptr(mean), K.ptr(rstd), K.ptr(y2), ldy2, y2_p, y2_seed, K.ptr(lora_a), K.ptr(lora_u), ld_u,
                               K.ptr(self.seed_dev) if y2_p > 0.0 else None)
        K.check(self.lib.bfk_layernorm_fwd(C.byref(a), _stream()), "bfk_layernorm_fwd")
        self.launches += 2 if lora_a is not None else 1      # + the tensor-core LoRA down-projection over the rows just written
        self.work["ln_bytes"] += M * D * (4.0 + y.element_size() + (y2.element_size() if y2 is not None else 0))
        if lora_a is not None:
            self.work["ln_bytes"] += M * D * 2.0      # the down-projection reads the (dropped) bf16 rows again

    def layernorm_bwd(self, dy, ld_dy, x, M, D, mean, rstd, gamma, dres, dx, dx2=None, dgamma=None, dbeta=None,
                      dy2=None, ld_dy2=0, dy2_p=0.0, dy2_seed=0):
        ws = None
        if dgamma is not None:
            ws = torch.empty(self.lib.bfk_layernorm_bwd_workspace_floats(M, D), device=x.device, dtype=torch.float32)
        a = K.LayerNormBwdArgs(M, D, dy.data_ptr(), ld_dy, _dt(dy), x.data_ptr(), D, mean.data_ptr(), rstd.data_ptr(),
                               gamma.data_ptr(), K.ptr(dres), D, dx.data_ptr(), D, K.ptr(dx2), D,
                               _dt(dx2) if dx2 is not None else 0, K.ptr(dgamma), K.ptr(dbeta), K.ptr(ws),
                               K.ptr(dy2), ld_dy2, dy2_p, dy2_seed, K.ptr(self.seed_dev) if dy2_p > 0.0 else None)
        K.check(self.lib.bfk_layernorm_bwd(C.byref(a), _stream()), "bfk_layernorm_bwd")
        self.launches += 2 if dgamma is not None else 1
        self.work["ln_bytes"] += M * D * (dy.element_size() + 4.0 + (4.0 if dres is not None else 0) + 4.0 +
                                          (dx2.element_size() if dx2 is not None else 0) + (dy2.element_size() if dy2 is not None else 0))

    # -- attention -----------------------------------------------------------------------------------
    def attention_fwd(self, qkv, out, lse, B, N, H, hd, scale, row_mask=None, mask_fill=-100000.0):
        a = K.AttentionArgs(B, N, H, hd, qkv.data_ptr(), 3 * H * hd, out.data_ptr(), H * hd, K.ptr(lse), K.ptr(row_mask),
                            scale, mask_fill, _dt(qkv), self.backend, None, 0, None, 0, None)
        K.check(self.lib.bfk_attention_fwd(C.byref(a), _stream()), "bfk_attention_fwd")
        self.launches += 1
        self.work["attn_flops"] += 4.0 * B * H * N * N * hd            # S = QK^T and O = PV

    def attention_bwd(self, qkv, out, lse, dout, dqkv, B, N, H, hd, scale, row_mask=None, mask_fill=-100000.0):
        delta = torch.empty(B * H * N, device=qkv.device, dtype=torch.float32)
        a = K.AttentionArgs(B, N, H, hd, qkv.data_ptr(), 3 * H * hd, out.data_ptr(), H * hd, lse.data_ptr(), K.ptr(row_mask),
                            scale, mask_fill, _dt(qkv), self.backend, dout.data_ptr(), H * hd, dqkv.data_ptr(), 3 * H * hd,
                            delta.data_ptr())
        K.check(self.lib.bfk_attention_bwd(C.byref(a), _stream()), "bfk_attention_bwd")
        self.launches += 3
        self.work["attn_flops"] += 8.0 * B * H * N * N * hd            # dP, dV, dK, dQ (recomputing S is not algorithmic work)

    def cross_attention(self, qkv, out, lse, B, Nf, H, hd, q0, nq, k0, nk, scale, key_bias=None, dout=None, dqkv=None, weights=None):
        """Alternate fusion module (fusion.py:67-125): rows [q0, q0+nq) of every sample attend to rows [k0, k0+nk) of the packed
        projection.  Forward when ``dout`` is None (``weights``: optional fp32 [B, nq, nk] head-averaged attention weights), else
        backward into ``dqkv``."""
        delta = torch.empty(B * H * nq, device=qkv.device, dtype=torch.float32) if dout is not None else None
        a = K.CrossAttentionArgs(B, Nf, H, hd, q0, nq, k0, nk, _dt(qkv), qkv.data_ptr(), 3 * H * hd, scale, K.ptr(key_bias),
                                 out.data_ptr(), H * hd, lse.data_ptr(), K.ptr(dout), H * hd, K.ptr(dqkv), 3 * H * hd, K.ptr(delta))
        if dout is None:
            K.check(self.lib.bfk_cross_attention_fwd(C.byref(a), _stream()), "bfk_cross_attention_fwd")
            self.launches += 1
            self.work["attn_flops"] += 4.0 * B * H * nq * nk * hd
            if weights is not None:
                K.check(self.lib.bfk_cross_attention_weights(C.byref(a), weights.data_ptr(), _stream()), "bfk_cross_attention_weights")
                self.launches += 1
        else:
            K.check(self.lib.bfk_cross_attention_bwd(C.byref(a), _stream()), "bfk_cross_attention_bwd")
            self.launches += 3
            self.work["attn_flops"] += 8.0 * B * H * nq * nk * hd

    # -- misc ------------------------------------------------------------------------------------------
    def cast(self, src, dst, rows, cols, ld_src, ld_dst):
        K.check(self.lib.bfk_cast_f32_to(src.data_ptr(), dst.data_ptr(), _dt(dst), rows, cols, ld_src, ld_dst, _stream()),
                "bfk_cast_f32_to")
        self.launches += 1


class _Nvtx:
    """BFK_NVTX=1: NVTX ranges around the phases of a step (towers / fusion / tail, forward and backward) for nsys /
    ncu --nvtx timelines.  Off by default: a range push / pop is a host call per phase."""
    on = os.environ.get("BFK_NVTX") == "1"

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _Nvtx.on:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _Nvtx.on:
            torch.cuda.nvtx.range_pop()
        return False


class LayerWeights:
    """Operand-dtype copies ("shadows") of one pre-LN transformer layer's weights."""
    __slots__ = ("ln1_w", "ln1_b", "Wqkv", "bqkv", "Acat", "Wo", "bo", "ln2_w", "ln2_b", "W1", "b1", "W2", "b2",
                 "Dk", "names")


class Engine:
    def __init__(self, cfg: PolicyConfig, params: Dict[str, torch.nn.Parameter], precision: str = "bf16"):
        assert precision in ("bf16", "fp32")
        if cfg.lora and cfg.r > 8:
            raise K.BfkError("LoRA rank > 8 is not supported by the packed QKV layout")
        self.cfg = cfg
        self.params = params
        self.precision = precision
        self.T = torch.bfloat16 if precision == "bf16" else torch.float32
        self.k = Kernels()
        self.device = None
        self._frozen_ready = False
        self._train_version = None
        self._taps = {}
        self._ptr_tables = {}
        self._frozen_version = None
        self.step_seed = 0
        self.tw_prefix = "siglip_model.model." if cfg.lora else "siglip_model."
        self.grad_ready_hook = None   # DDP: called as hook(name_list) when a group of grads is final
        self.overlap_towers = os.environ.get("BFK_OVERLAP_TOWERS", "1") != "0"    # text tower on a side stream, concurrent with the vision tower (fwd and bwd)
        self._side = None
        self.fuse_lora_down = os.environ.get("BFK_FUSE_LORA_DOWN", "1") != "0"
        self.overlap_wgrad = os.environ.get("BFK_OVERLAP_WGRAD", "1") != "0"
        self.use_opt_shadow = os.environ.get("BFK_OPT_SHADOW", "1") != "0"     # fusion GEMM weights: bf16 copies written by the Adam kernel
        self._wgrad_streams, self._wgrad_busy = {}, False

    # ------------------------------------------------------------------------------------------------
    # weight preparation
    # ------------------------------------------------------------------------------------------------
    def _p(self, name) -> torch.Tensor:
        return self.params[name].data

    def _shadow(self, w: torch.Tensor, cols_pad: int = 0) -> torch.Tensor:
        """[rows, cols] fp32 -> T [rows, cols + cols_pad] (pad columns zero)."""
        rows, cols = w.shape[0], w[0].numel()
        if self.T == torch.float32 and cols_pad == 0:
            return w.reshape(rows, cols)
        out = torch.zeros(rows, cols + cols_pad, device=w.device, dtype=self.T) if cols_pad else \
            torch.empty(rows, cols, device=w.device, dtype=self.T)
        self.k.cast(w, out, rows, cols, cols, cols + cols_pad)
        return out

    def _tower_layer_weights(self, prefix: str) -> LayerWeights:
        cfg, D = self.cfg, self.cfg.dim
        lw = LayerWeights()
        a = prefix + ".self_attn."
        lw.names = {}
        lw.ln1_w, lw.ln1_b = self._p(prefix + ".layer_norm1.weight"), self._p(prefix + ".layer_norm1.bias")
        lw.ln2_w, lw.ln2_b = self._p(prefix + ".layer_norm2.weight"), self._p(prefix + ".layer_norm2.bias")
        base = ".base_layer" if cfg.lora else ""
        lw.Dk = D + 16 if cfg.lora else D
        Wqkv = torch.zeros(3 * D, lw.Dk, device=self.device, dtype=self.T)
        bqkv = torch.empty(3 * D, device=self.device, dtype=torch.float32)
        for i, proj in enumerate(("q_proj", "k_proj", "v_proj")):
            b = base if proj != "k_proj" else ""
            self.k.cast(self._p(f"{a}{proj}{b}.weight"), Wqkv[i * D:], D, D, D, lw.Dk)
            bqkv[i * D:(i + 1) * D].copy_(self._p(f"{a}{proj}{b}.bias"))
        lw.Wqkv, lw.bqkv = Wqkv, bqkv
        lw.Acat = torch.zeros(16, D, device=self.device, dtype=self.T) if cfg.lora else None
        if cfg.lora:
            lw.names = {k: f"{a}{p}.lora_{m}.siglip_adapter.weight" for k, p, m in
                        (("Aq", "q_proj", "A"), ("Bq", "q_proj", "B"), ("Av", "v_proj", "A"), ("Bv", "v_proj", "B"))}
        lw.Wo, lw.bo = self._shadow(self._p(a + "out_proj.weight")), self._p(a + "out_proj.bias")
        lw.W1, lw.b1 = self._shadow(self._p(prefix + ".mlp.fc1.weight")), self._p(prefix + ".mlp.fc1.bias")
        lw.W2, lw.b2 = self._shadow(self._p(prefix + ".mlp.fc2.weight")), self._p(prefix + ".mlp.fc2.bias")
        return lw

    def _lora_segments(self, lw: LayerWeights, segs):
        """Trainable LoRA factors -> their slots in the extended QKV weight [W | (alpha/r) B] and the [A_q; A_v] pack."""
        cfg, D, r, Dk = self.cfg, self.cfg.dim, self.cfg.r, lw.Dk
        s = cfg.lora_alpha / cfg.r
        es = lw.Wqkv.element_size()
        n = lw.names
        segs.append((self._p(n["Bq"]), lw.Wqkv.data_ptr() + D * es, D, r, Dk, s, lw.Wqkv.dtype))
        segs.append((self._p(n["Bv"]), lw.Wqkv.data_ptr() + (2 * D * Dk + D + 8) * es, D, r, Dk, s, lw.Wqkv.dtype))
        segs.append((self._p(n["Aq"]), lw.Acat.data_ptr(), r, D, D, 1.0, lw.Acat.dtype))
        segs.append((self._p(n["Av"]), lw.Acat.data_ptr() + 8 * D * lw.Acat.element_size(), r, D, D, 1.0, lw.Acat.dtype))

    def _fusion_layer_weights(self, l: int) -> LayerWeights:
        D = self.cfg.dim
        p = f"pick_place.fusion.transformer_encoder.layers.{l}"
        lw = LayerWeights()
        lw.Dk = D
        lw.names = {"ln1_w": p + ".0.norm.weight", "ln1_b": p + ".0.norm.bias", "Wqkv": p + ".0.fn.to_qkv.weight",
                    "Wo": p + ".0.fn.to_out.0.weight", "bo": p + ".0.fn.to_out.0.bias",
                    "ln2_w": p + ".1.norm.weight", "ln2_b": p + ".1.norm.bias",
                    "W1": p + ".1.fn.net.0.weight", "b1": p + ".1.fn.net.0.bias",
                    "W2": p + ".1.fn.net.3.weight", "b2": p + ".1.fn.net.3.bias"}
        lw.bqkv, lw.Acat = None, None
        return lw

    def _fusion_segments(self, lw: LayerWeights, segs):
        n = lw.names
        lw.ln1_w, lw.ln1_b, lw.ln2_w, lw.ln2_b = (self._p(n[k]) for k in ("ln1_w", "ln1_b", "ln2_w", "ln2_b"))
        lw.bo, lw.b1, lw.b2 = self._p(n["bo"]), self._p(n["b1"]), self._p(n["b2"])
        for k in ("Wqkv", "Wo", "W1", "W2"):
            w = self._p(n[k])
            if self.T == torch.float32:
                setattr(lw, k, w)
                continue
            sh = self._opt_shadow(self.params[n[k]]) if self.use_opt_shadow else None
            if sh is not None:
                # the optimizer's Adam kernel keeps this bf16 copy current (FusedAdam.ensure_shadow): a cast is only needed when
                # the parameter changed some other way (load_state_dict, manual edits)
                setattr(lw, k, sh)
                par = self.params[n[k]]
                if getattr(par, "_bfk_shadow_ver", None) != par._version:
                    segs.append((w, sh.data_ptr(), w.shape[0], w.shape[1], w.shape[1], 1.0, self.T))
                    par._bfk_shadow_ver = par._version
            else:
                cur = getattr(lw, k, None) if hasattr(lw, k) else None
                if cur is None or cur.shape != w.shape:
                    cur = torch.empty_like(w, dtype=self.T)
                    setattr(lw, k, cur)
                segs.append((w, cur.data_ptr(), w.shape[0], w.shape[1], w.shape[1], 1.0, self.T))

    def _opt_shadow(self, par):
        """bf16 view of `par` inside its optimizer's flat shadow buffer, or None (no FusedAdam yet / parameter re-homed)."""
        info = getattr(par, "_bfk_opt", None)
        if info is None:
            return None
        opt, gi, off = info[0](), info[1], info[2]
        st = opt._flat.get(gi) if opt is not None else None
        if st is None or par.data_ptr() != st["p"].data_ptr() + 4 * off or par.dtype != torch.float32:
            return None
        return opt.ensure_shadow(gi)[off:off + par.numel()].view(par.shape)

    def _refresh_trainable(self):
        """ONE launch (bfk_multi_cast) re-derives every operand-dtype copy of the trainable weights."""
        segs = []
        if self.cfg.lora:
            for lw in self.text_layers + self.vis_layers:
                self._lora_segments(lw, segs)
        for lw in self.fusion_layers:
            self._fusion_segments(lw, segs)
        if self.cfg.fusion_model == "crossattention":
            xp = "pick_place.fusion.cross_attention."
            for attr, name in (("xa_Wi", xp + "in_proj_weight"), ("xa_Wo", xp + "out_proj.weight")):
                w = self._p(name)
                if self.T == torch.float32:
                    setattr(self, attr, w)
                else:
                    cur = getattr(self, attr, None)
                    if cur is None or cur.shape != w.shape or cur.device != w.device:
                        cur = torch.empty_like(w, dtype=self.T)
                        setattr(self, attr, cur)
                    segs.append((w, cur.data_ptr(), w.shape[0], w.shape[1], w.shape[1], 1.0, self.T))
        if not segs:
            return
        key = tuple((w.data_ptr(), dst) for (w, dst, *_rest) in segs)
        if getattr(self, "_seg_key", None) != key:
            arr = (K.CastSegment * len(segs))()
            chunk = 0
            for i, (w, dst, rows, cols, ld, scale, dt) in enumerate(segs):
                assert w.dtype == torch.float32 and w.is_contiguous()
                arr[i] = K.CastSegment(w.data_ptr(), dst, ld, chunk, rows, cols, K.BF16 if dt == torch.bfloat16 else K.F32, scale)
                chunk += (rows * cols + 1023) // 1024
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).pin_memory()
            self._seg_host, self._seg_dev = host, host.to(self.device, non_blocking=True)
            self._seg_key, self._seg_n, self._seg_chunks = key, len(segs), chunk
        K.check(self.k.lib.bfk_multi_cast(self._seg_dev.data_ptr(), self._seg_n, self._seg_chunks, _stream()), "bfk_multi_cast")
        self.k.launches += 1

    def prepare(self, device):
        """(Re)build operand-dtype shadows.  Frozen tower weights once; trainable ones when changed."""
        cfg = self.cfg
        if self.device != device:
            self.device, self._frozen_ready, self._train_version = device, False, None
            self._taps, self._ptr_tables = {}, {}
            self.k.seed_dev = torch.zeros(1, device=device, dtype=torch.int64)
        fver = sum(p._version + (p.data_ptr() % 1000003) for p in self.params.values() if not p.requires_grad)
        if fver != self._frozen_version:   # e.g. load_state_dict after a forward
            self._frozen_ready, self._frozen_version, self._train_version = False, fver, None
        if not self._frozen_ready:
            tp, vp = self.tw_prefix + "text_model", self.tw_prefix + "vision_model"
            self.text_layers = [self._tower_layer_weights(f"{tp}.encoder.layers.{i}") for i in range(cfg.tower_layers)]
            self.vis_layers = [self._tower_layer_weights(f"{vp}.encoder.layers.{i}") for i in range(cfg.tower_layers)]
            kpe = 3 * cfg.patch_size * cfg.patch_size
            self.kpe_ld = (kpe + 7) // 8 * 8
            self.W_pe = self._shadow(self._p(vp + ".embeddings.patch_embedding.weight"), self.kpe_ld - kpe)
            self.fusion_layers = [self._fusion_layer_weights(l) for l in range(cfg.depth if cfg.fusion_model == "concat_transformer" else 0)]
            self.xa_Wi = self.xa_Wo = None
            for lw in self.fusion_layers:
                for k in ("Wqkv", "Wo", "W1", "W2"):
                    setattr(lw, k, None)
            self._frozen_ready = True
        version = sum(p._version for p in self.params.values() if p.requires_grad) + \
            sum(p.data_ptr() % 1000003 for p in self.params.values() if p.requires_grad)
        if version != self._train_version:
            self._refresh_trainable()
            self._train_version = version

    def taps(self, g: int):
        if g not in self._taps:
            H = 16 * g
            ts = torch.empty(H, dtype=torch.int32)
            tw = torch.empty(H, 3, dtype=torch.float32)
            K.check(self.k.lib.bfk_upsample16_taps(g, ts.data_ptr(), tw.data_ptr()), "bfk_upsample16_taps")
            self._taps[g] = (ts.to(self.device), tw.to(self.device))
        return self._taps[g]

    # ------------------------------------------------------------------------------------------------
    # one pre-LN transformer layer
    # ------------------------------------------------------------------------------------------------
    def _layer_fwd(self, x, nseq, N, lw: LayerWeights, heads, eps, act, row_mask, save, lora_seed=0):
        """x: fp32 [nseq*N, D] residual stream.  Returns the new stream; appends saved tensors to `save`."""
        k, D, T, dev = self.k, self.cfg.dim, self.T, x.device
        M, Dk, hd = nseq * N, lw.Dk, D // heads
        Mlp = lw.W1.shape[0]
        h1 = torch.empty(M, Dk, device=dev, dtype=T)
        mean1 = torch.empty(M, device=dev, dtype=torch.float32); rstd1 = torch.empty_like(mean1)
        p_drop = self.cfg.lora_dropout if (lw.Acat is not None and self.training) else 0.0
        h1d = None
        if lw.Acat is not None and T == torch.bfloat16 and p_drop > 0.0:
            h1d = torch.empty(M, D, device=dev, dtype=T)   # drop(h1): the LoRA branch input, written by the same LN pass
        lora_bf16 = lw.Acat is not None and T == torch.bfloat16
        fuse_lora = lora_bf16 and self.fuse_lora_down
        # bf16: the LayerNorm pass also computes the LoRA down-projection u = drop(h1) Acat^T into the 16 extra K columns
        k.layernorm_fwd(x, M, D, lw.ln1_w, lw.ln1_b, eps, h1, Dk, mean1, rstd1, y2=h1d, ldy2=D, y2_p=p_drop, y2_seed=lora_seed,
                        lora_a=lw.Acat if fuse_lora else None, lora_u=h1[:, D:] if fuse_lora else None, ld_u=Dk)
        if lora_bf16 and not fuse_lora:      # measurement only: the same product as an N = 16 tensor-core GEMM
            lora_in, ld_in = (h1d, D) if h1d is not None else (h1, Dk)
            k.gemm(lora_in, lw.Acat, h1[:, D:], M, 16, D, ld_in, D, Dk)
        elif lw.Acat is not None and not lora_bf16:      # fp32 parity mode: SIMT kernel
            K.check(k.lib.bfk_lora_down(h1.data_ptr(), Dk, _dt(h1), M, D, lw.Acat.data_ptr(), p_drop, lora_seed, _stream()),
                    "bfk_lora_down")
            k.launches += 1
        qkv = torch.empty(M, 3 * D, device=dev, dtype=T)
        k.gemm(h1, lw.Wqkv, qkv, M, 3 * D, Dk, Dk, Dk, 3 * D, bias=lw.bqkv)
        ctx = torch.empty(M, D, device=dev, dtype=T)
        lse = torch.empty(nseq * heads * N, device=dev, dtype=torch.float32)
        k.attention_fwd(qkv, ctx, lse, nseq, N, heads, hd, hd ** -0.5, row_mask)
        x1 = torch.empty(M, D, device=dev, dtype=torch.float32)
        k.gemm(ctx, lw.Wo, x1, M, D, D, D, D, D, bias=lw.bo, residual=x, ldr=D)
        h2 = torch.empty(M, D, device=dev, dtype=T)
        mean2 = torch.empty(M, device=dev, dtype=torch.float32); rstd2 = torch.empty_like(mean2)
        k.layernorm_fwd(x1, M, D, lw.ln2_w, lw.ln2_b, eps, h2, D, mean2, rstd2)
        pa = torch.empty(M, Mlp, device=dev, dtype=T) if save is not None else None
        a = torch.empty(M, Mlp, device=dev, dtype=T)
        # bf16 training: the fc1 epilogue stores gelu'(pre-activation) instead of the pre-activation, so the fc2 dgrad
        # epilogue is a plain multiply (tanh / exp evaluated once per element instead of twice)
        pa_is_grad = pa is not None and T == torch.bfloat16
        act_f = act if not pa_is_grad else (K.ACT_GELU_TANH_GRAD if act == K.ACT_GELU_TANH else K.ACT_GELU_ERF_GRAD)
        k.gemm(h2, lw.W1, a, M, Mlp, D, D, D, Mlp, bias=lw.b1, act=act_f, aux=pa, ld_aux=Mlp)
        x2 = torch.empty(M, D, device=dev, dtype=torch.float32)
        k.gemm(a, lw.W2, x2, M, D, Mlp, Mlp, Mlp, D, bias=lw.b2, residual=x1, ldr=D)
        if save is not None:
            save.append(dict(x=x, h1=h1, mean1=mean1, rstd1=rstd1, qkv=qkv, ctx=ctx, lse=lse, x1=x1, h2=h2, mean2=mean2,
                             rstd2=rstd2, pa=pa, a=a, p_drop=p_drop, seed=lora_seed, h1d=h1d, pa_is_grad=pa_is_grad))
        return x2

    def _layer_bwd(self, dx2, dx2_T, sv, nseq, N, lw: LayerWeights, heads, eps, act, row_mask, grads, trainable):
        """dx2 (fp32) / dx2_T (operand dtype): gradient w.r.t. the layer output.  Returns (dx, dx_T)."""
        k, D, T, dev = self.k, self.cfg.dim, self.T, dx2.device
        M, Dk, hd = nseq * N, lw.Dk, D // heads
        Mlp = lw.W1.shape[0]
        dact = K.ACT_MUL_AUX if sv.get("pa_is_grad") else (K.ACT_DGELU_TANH if act == K.ACT_GELU_TANH else K.ACT_DGELU_ERF)
        n = lw.names
        # fc2: x2 = a W2^T + b2 + x1
        dpa = torch.empty(M, Mlp, device=dev, dtype=T)
        k.gemm(dx2_T, lw.W2, dpa, M, Mlp, D, D, Mlp, Mlp, trans_b=True, act=dact, aux=sv["pa"], ld_aux=Mlp)
        if trainable:
            def wg_fc2():
                k.gemm(dx2_T, sv["a"], grads[n["W2"]], D, Mlp, M, D, Mlp, Mlp, trans_a=True, trans_b=True)
                k.colsum(dx2, M, D, D, grads[n["b2"]])
            self._off_chain(wg_fc2, dx2_T, sv["a"], dx2)
        # fc1: pa = h2 W1^T + b1
        dh2 = torch.empty(M, D, device=dev, dtype=T)
        k.gemm(dpa, lw.W1, dh2, M, D, Mlp, Mlp, D, D, trans_b=True)
        if trainable:
            def wg_fc1(dpa=dpa):
                k.gemm(dpa, sv["h2"], grads[n["W1"]], Mlp, D, M, Mlp, D, D, trans_a=True, trans_b=True)
                k.colsum(dpa, M, Mlp, Mlp, grads[n["b1"]])
            self._off_chain(wg_fc1, dpa, sv["h2"])
        del dpa
        # LN2 + residual
        dx1 = torch.empty(M, D, device=dev, dtype=torch.float32)
        dx1_T = torch.empty(M, D, device=dev, dtype=T) if T != torch.float32 else None
        k.layernorm_bwd(dh2, D, sv["x1"], M, D, sv["mean2"], sv["rstd2"], lw.ln2_w, dx2, dx1, dx1_T,
                        grads[n["ln2_w"]] if trainable else None, grads[n["ln2_b"]] if trainable else None)
        if dx1_T is None:
            dx1_T = dx1
        # out-proj: x1 = ctx Wo^T + bo + x
        dctx = dh2  # reuse buffer [M, D]
        k.gemm(dx1_T, lw.Wo, dctx, M, D, D, D, D, D, trans_b=True)
        if trainable:
            def wg_out():
                k.gemm(dx1_T, sv["ctx"], grads[n["Wo"]], D, D, M, D, D, D, trans_a=True, trans_b=True)
                k.colsum(dx1, M, D, D, grads[n["bo"]])
            self._off_chain(wg_out, dx1_T, sv["ctx"], dx1)
        # attention
        dqkv = torch.empty(M, 3 * D, device=dev, dtype=T)
        k.attention_bwd(sv["qkv"], sv["ctx"], sv["lse"], dctx, dqkv, nseq, N, heads, hd, hd ** -0.5, row_mask)
        # QKV: qkv = h1 Wqkv^T (+ bqkv); dh1 carries the 16 LoRA columns when present
        dh1 = torch.empty(M, Dk, device=dev, dtype=T)
        k.gemm(dqkv, lw.Wqkv, dh1, M, Dk, 3 * D, 3 * D, Dk, Dk, trans_b=True)
        if trainable:
            self._off_chain(lambda: k.gemm(dqkv, sv["h1"], grads[n["Wqkv"]], 3 * D, D, M, 3 * D, Dk, D, trans_a=True, trans_b=True),
                            dqkv, sv["h1"])
        dy2 = None
        if lw.Acat is not None:
            dy2 = self._lora_bwd(sv, dqkv, dh1, M, lw, grads)
        # LN1 + residual
        dx = torch.empty(M, D, device=dev, dtype=torch.float32)
        dx_T = torch.empty(M, D, device=dev, dtype=T) if T != torch.float32 else None
        k.layernorm_bwd(dh1, Dk, sv["x"], M, D, sv["mean1"], sv["rstd1"], lw.ln1_w, dx1, dx, dx_T,
                        grads[n["ln1_w"]] if trainable else None, grads[n["ln1_b"]] if trainable else None,
                        dy2=dy2, ld_dy2=D, dy2_p=sv["p_drop"] if dy2 is not None else 0.0, dy2_seed=sv["seed"])
        return dx, (dx_T if dx_T is not None else dx)

    # ------------------------------------------------------------------------------------------------
    # alternate fusion module: CrossAttention (bifold/models/fusion.py:67-125)
    # ------------------------------------------------------------------------------------------------
    def _xattn_fwd(self, X, B, Nf, mask, save):
        """X: fp32 [B*Nf, D] = per sample [text token + text | context frames | image token + patches], type embeddings added by
        the assembly kernel.  ONE packed projection of all rows with MultiheadAttention's in_proj (q of the condition rows and
        k / v of the image rows are computed and ignored: 3 x the minimum, at no extra launch), the image rows attend to the
        condition rows, out_proj.  Returns fp32 [B*Nf, D]; only the image rows (the last Np + 1 of a sample) are meaningful -
        the heatmap tail reads exactly those.  `mask` ([B*Nf] fp32, SiglipSequential): its first Nc entries per sample are
        the additive key bias (float attn_mask, fusion.py:107-113)."""
        cfg, k, D, T, dev = self.cfg, self.k, self.cfg.dim, self.T, X.device
        M, H = B * Nf, cfg.heads
        hd, nq = D // H, cfg.num_patches + 1
        nc = Nf - nq
        xp = "pick_place.fusion.cross_attention."
        X_T = X
        if T != torch.float32:
            X_T = torch.empty(M, D, device=dev, dtype=T)
            k.cast(X, X_T, M, D, D, D)
        qkv = torch.empty(M, 3 * D, device=dev, dtype=T)
        k.gemm(X_T, self.xa_Wi, qkv, M, 3 * D, D, D, D, 3 * D, bias=self._p(xp + "in_proj_bias"))
        bias = None
        if mask is not None:
            bias = mask.view(B, Nf)[:, :nc].contiguous()
        att = torch.zeros(M, D, device=dev, dtype=T)          # condition rows stay zero: they meet zero gradients in the wgrad
        lse = torch.empty(B * H * nq, device=dev, dtype=torch.float32)
        # the module's second return value (head-averaged weights, fusion.py:117-125) -> model output "attn_weights"
        self.last_attn_weights = torch.empty(B, nq, nc, device=dev, dtype=torch.float32)
        k.cross_attention(qkv, att, lse, B, Nf, H, hd, nc, nq, 0, nc, hd ** -0.5, key_bias=bias, weights=self.last_attn_weights)
        out = torch.empty(M, D, device=dev, dtype=torch.float32)
        k.gemm(att, self.xa_Wo, out, M, D, D, D, D, D, bias=self._p(xp + "out_proj.bias"))
        if save is not None:
            save.append(dict(X_T=X_T, qkv=qkv, att=att, lse=lse, bias=bias, nc=nc, nq=nq))
        return out

    def _xattn_bwd(self, dx, dx_T, sv, B, Nf, mask, grads):
        """dx: fp32 [B*Nf, D] gradient w.r.t. the module output (zero outside the image rows).  Returns dX fp32 [B*Nf, D]."""
        cfg, k, D, T, dev = self.cfg, self.k, self.cfg.dim, self.T, dx.device
        M, H = B * Nf, cfg.heads
        hd = D // H
        xp = "pick_place.fusion.cross_attention."
        if dx_T is None or dx_T is dx and T != torch.float32:
            dx_T = torch.empty(M, D, device=dev, dtype=T)
            k.cast(dx, dx_T, M, D, D, D)
        # out_proj: out = att Wo^T + bo
        k.colsum(dx, M, D, D, grads[xp + "out_proj.bias"])
        k.gemm(dx_T, sv["att"], grads[xp + "out_proj.weight"], D, D, M, D, D, D, trans_a=True, trans_b=True)
        datt = torch.empty(M, D, device=dev, dtype=T)
        k.gemm(dx_T, self.xa_Wo, datt, M, D, D, D, D, D, trans_b=True)
        dqkv = torch.zeros(M, 3 * D, device=dev, dtype=T)
        k.cross_attention(sv["qkv"], sv["att"], sv["lse"], B, Nf, H, hd, sv["nc"], sv["nq"], 0, sv["nc"], hd ** -0.5,
                          key_bias=sv["bias"], dout=datt, dqkv=dqkv)
        # in_proj: qkv = X Wi^T + bi
        k.colsum(dqkv, M, 3 * D, 3 * D, grads[xp + "in_proj_bias"])
        k.gemm(dqkv, sv["X_T"], grads[xp + "in_proj_weight"], 3 * D, D, M, 3 * D, D, D, trans_a=True, trans_b=True)
        dX = torch.empty(M, D, device=dev, dtype=torch.float32)
        k.gemm(dqkv, self.xa_Wi, dX, M, D, 3 * D, 3 * D, D, D, trans_b=True)
        return dX

    # Weight / bias gradients are leaves of the backward dependency chain: they run on a second stream (in a captured
    # step: a parallel graph branch) and fill the SM time the chain's kernels leave idle (tails, short LayerNorm passes).
    def _off_chain(self, fn, *inputs):
        cur = torch.cuda.current_stream()
        ws = self._wgrad_streams.get(cur.cuda_stream)
        if ws is None:
            ws = self._wgrad_streams[cur.cuda_stream] = torch.cuda.Stream(device=cur.device)
        # Eager multi-process runs keep the gradients on the chain (measured: a third compute stream plus record_stream
        # churn made the eager DDP step slower); inside a captured graph the branch is free again.
        if not self.overlap_wgrad or (self.grad_ready_hook is not None and not torch.cuda.is_current_stream_capturing()):
            fn()
            return
        ws.wait_stream(cur)                       # everything the gradient reads has been enqueued on `cur`
        with torch.cuda.stream(ws):
            fn()
        for t in inputs:
            if t is not None:
                t.record_stream(ws)
        self._wgrad_busy = True

    def _join_off_chain(self, hook_group=None):
        """Make the current stream wait for the off-chain gradients; a DDP hook is announced once they are final."""
        cur = torch.cuda.current_stream()
        ws = self._wgrad_streams.get(cur.cuda_stream)
        if ws is not None and self._wgrad_busy:
            cur.wait_stream(ws)
            self._wgrad_busy = False

    def _hook_after_off_chain(self, group):
        """Announce a gradient group whose members were written on the current stream AND the off-chain stream."""
        cur = torch.cuda.current_stream()
        ws = self._wgrad_streams.get(cur.cuda_stream)
        if ws is None or not self._wgrad_busy:
            self.grad_ready_hook(group)
            return
        ws.wait_stream(cur)
        with torch.cuda.stream(ws):
            self.grad_ready_hook(group)           # the hook orders the all-reduce behind the stream it is called on

    def _lora_bwd(self, sv, dqkv, dh1, M, lw, grads):
        """LoRA factor gradients and the rank-r contribution to dh1 (peft LoRA backward)."""
        k, cfg, D, r = self.k, self.cfg, self.cfg.dim, self.cfg.r
        s = cfg.lora_alpha / cfg.r
        lib, st = k.lib, _stream()
        es = dqkv.element_size()
        h1, Dk = sv["h1"], lw.Dk
        n = lw.names
        if self.T == torch.bfloat16:
            # tensor-core path: split-K "wgrad" GEMMs for the factors, one K=16 GEMM for the dh1 contribution
            x_in, ld_x = (sv["h1d"], D) if sv["h1d"] is not None else (h1, Dk)
            if r <= 8 and D % 8 == 0:
                # both B factors from one dqkv^T u product, both A factors from one du^T x product, one fused reduction
                a = K.LoraGradsArgs(M, D, r, dqkv.data_ptr(), 3 * D, h1.data_ptr() + D * es, Dk, dh1.data_ptr() + D * es, Dk,
                                    x_in.data_ptr(), ld_x, s, grads[n["Bq"]].data_ptr(), grads[n["Bv"]].data_ptr(),
                                    grads[n["Aq"]].data_ptr(), grads[n["Av"]].data_ptr(), None, 0)
                def lora_grads(a=a):
                    ws = k.lora_ws(dqkv.device, D)            # keyed by the stream the products run on
                    a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel() * 4
                    K.check(lib.bfk_lora_grads(C.byref(a), _stream()), "bfk_lora_grads")
                    k.launches += 3
                self._off_chain(lora_grads, dqkv, h1, dh1, x_in)
            else:
                for (Bn, ucol, qoff) in ((n["Bq"], 0, 0), (n["Bv"], 8, 2 * D)):
                    k.gemm(dqkv[:, qoff:], h1[:, D + ucol:], grads[Bn], D, r, M, 3 * D, Dk, r, trans_a=True, trans_b=True, alpha=s)
                for (An, ucol) in ((n["Aq"], 0), (n["Av"], 8)):
                    k.gemm(dh1[:, D + ucol:], x_in, grads[An], r, D, M, Dk, ld_x, D, trans_a=True, trans_b=True)
            dy2 = torch.empty(M, D, device=dqkv.device, dtype=self.T)
            k.gemm(dh1[:, D:], lw.Acat, dy2, M, D, 16, Dk, D, D, trans_b=True)
            return dy2
        ws = torch.empty(max(lib.bfk_skinny_gram_workspace_floats(M, 16, D), 1), device=dqkv.device, dtype=torch.float32)
        # dB_q[d, i] = s * sum_m u_q[m,i] dq[m,d]   (u = h1[:, D:D+16]);  out stored transposed: [D, r]
        for (Bn, ucol, qoff) in ((n["Bq"], 0, 0), (n["Bv"], 8, 2 * D)):
            K.check(lib.bfk_skinny_gram(h1.data_ptr() + (D + ucol) * es, Dk, r, dqkv.data_ptr() + qoff * es, 3 * D, D,
                                        _dt(dqkv), M, s, grads[Bn].data_ptr(), r, 1, 0.0, 0, ws.data_ptr(), st), "bfk_skinny_gram")
        # dA[i, k] = sum_m du[m,i] * drop(h1)[m,k]   (du = dh1[:, D:D+16])
        for (An, ucol) in ((n["Aq"], 0), (n["Av"], 8)):
            K.check(lib.bfk_skinny_gram(dh1.data_ptr() + (D + ucol) * es, Dk, r, h1.data_ptr(), Dk, D, _dt(dqkv), M, 1.0,
                                        grads[An].data_ptr(), D, 0, sv["p_drop"], sv["seed"], ws.data_ptr(), st), "bfk_skinny_gram")
        K.check(lib.bfk_lora_down_bwd(dh1.data_ptr(), Dk, _dt(dh1), M, D, lw.Acat.data_ptr(), sv["p_drop"], sv["seed"], st),
                "bfk_lora_down_bwd")
        k.launches += 9
        return None

    # ------------------------------------------------------------------------------------------------
    # towers
    # ------------------------------------------------------------------------------------------------
    def _final_ln_fwd(self, x, M, wname, bname, save):
        D = self.cfg.dim
        y = torch.empty(M, D, device=x.device, dtype=torch.float32)
        mean = torch.empty(M, device=x.device, dtype=torch.float32); rstd = torch.empty_like(mean)
        self.k.layernorm_fwd(x, M, D, self._p(wname), self._p(bname), self.cfg.tower_ln_eps, y, D, mean, rstd)
        if save is not None:
            save.append(dict(x=x, mean=mean, rstd=rstd))
        return y

    def _vision_fwd(self, images, save):
        cfg, k, D = self.cfg, self.k, self.cfg.dim
        Bv, Np, p = images.shape[0], cfg.num_patches, cfg.patch_size
        vp = self.tw_prefix + "vision_model"
        kpe = 3 * p * p
        cols = torch.empty(Bv * Np, self.kpe_ld, device=images.device, dtype=self.T)
        if self.kpe_ld != kpe:
            cols.zero_()
        K.check(k.lib.bfk_im2col(images.data_ptr(), Bv, 3, cfg.image_size, cfg.image_size, p, cols.data_ptr(), self.kpe_ld,
                                 _dt(cols), _stream()), "bfk_im2col")
        k.launches += 1
        x = torch.empty(Bv * Np, D, device=images.device, dtype=torch.float32)
        k.gemm(cols, self.W_pe, x, Bv * Np, D, self.kpe_ld, self.kpe_ld, self.kpe_ld, D,
               bias=self._p(vp + ".embeddings.patch_embedding.bias"),
               residual=self._p(vp + ".embeddings.position_embedding.weight"), ldr=D, res_row_mod=Np)
        del cols
        for i, lw in enumerate(self.vis_layers):
            x = self._layer_fwd(x, Bv, Np, lw, cfg.tower_heads, cfg.tower_ln_eps, K.ACT_GELU_TANH, None, save,
                                lora_seed=self.step_seed * 4096 + 1000 + i)
        return self._final_ln_fwd(x, Bv * Np, vp + ".post_layernorm.weight", vp + ".post_layernorm.bias", save)

    def _text_fwd(self, ids, save):
        cfg, k, D = self.cfg, self.k, self.cfg.dim
        B, L = ids.shape
        tp = self.tw_prefix + "text_model"
        x = torch.empty(B * L, D, device=ids.device, dtype=torch.float32)
        K.check(k.lib.bfk_text_embed(ids.data_ptr(), B, L, D, cfg.vocab_size,
                                     self._p(tp + ".embeddings.token_embedding.weight").data_ptr(),
                                     self._p(tp + ".embeddings.position_embedding.weight").data_ptr(), x.data_ptr(), _stream()),
                "bfk_text_embed")
        k.launches += 1
        for i, lw in enumerate(self.text_layers):
            x = self._layer_fwd(x, B, L, lw, cfg.tower_heads, cfg.tower_ln_eps, K.ACT_GELU_TANH, None, save,
                                lora_seed=self.step_seed * 4096 + 2000 + i)
        return self._final_ln_fwd(x, B * L, tp + ".final_layer_norm.weight", tp + ".final_layer_norm.bias", save)

    def _tower_bwd(self, dy, saves, layers, nseq, N, final_w, grads):
        """dy fp32 [nseq*N, D] gradient w.r.t. the tower's last_hidden_state; LoRA grads into `grads`."""
        cfg, k, D, T = self.cfg, self.k, self.cfg.dim, self.T
        M = nseq * N
        fin = saves[-1]
        dx = torch.empty(M, D, device=dy.device, dtype=torch.float32)
        dx_T = torch.empty(M, D, device=dy.device, dtype=T) if T != torch.float32 else None
        k.layernorm_bwd(dy, D, fin["x"], M, D, fin["mean"], fin["rstd"], self._p(final_w), None, dx, dx_T)
        dx_T = dx if dx_T is None else dx_T
        for i in range(len(layers) - 1, -1, -1):
            dx, dx_T = self._layer_bwd(dx, dx_T, saves[i], nseq, N, layers[i], cfg.tower_heads, cfg.tower_ln_eps,
                                       K.ACT_GELU_TANH, None, grads, trainable=False)
            saves[i] = None
        # patch/text embeddings are frozen and the inputs need no gradient: stop here.

    # ------------------------------------------------------------------------------------------------
    # full forward
    # ------------------------------------------------------------------------------------------------
    def forward(self, sample, need_grad: bool, training: bool, targets=None, want_argmax=False):
        """Returns (heat [heads,B,H,W] fp32, losses [heads] or None, argmax keys or None, ctx or None)."""
        cfg, k, D = self.cfg, self.k, self.cfg.dim
        rgb = sample["rgb"]
        dev = rgb.device
        self.training = training
        if cfg.compute_mask:
            targets = None          # the pick outputs are sigmoid * mask: the BCE cannot be fused into the sigmoid kernel
        if training and need_grad and cfg.lora and cfg.lora_dropout > 0.0:
            self.step_seed += 1     # every training forward draws fresh LoRA-dropout masks (a captured graph adds seed_dev on top)
        self.prepare(dev)
        pre = self._collapse_fwd(dev, off_chain=True)     # needs only the decoder weights: overlaps the towers
        B, Np, L = rgb.shape[0], cfg.num_patches, cfg.text_len
        T_ctx = sample["rgb_context"].shape[1] if cfg.sequential else 0
        tower_grad = need_grad and cfg.lora
        ctx = dict(B=B, T=T_ctx) if need_grad else None
        vsave = [] if tower_grad else None
        tsave = [] if tower_grad else None
        fsave = [] if need_grad else None

        if T_ctx:
            images = torch.cat((rgb, sample["rgb_context"].reshape(B * T_ctx, *rgb.shape[1:])), dim=0).contiguous()
        else:
            images = rgb.contiguous()
        ids = sample["instruction"].contiguous()
        if self.overlap_towers:
            # the towers are independent until the token assembly: the small text-tower kernels (M = B*64 rows)
            # fill SMs the vision-tower kernels leave idle
            main = torch.cuda.current_stream()
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                txt = self._text_fwd(ids, tsave)                  # [B*L, D]
            with _Nvtx("bfk.vision_tower_fwd"):
                vis = self._vision_fwd(images.float(), vsave)     # [(B + B*T)*Np, D]
            main.wait_stream(self._side)
            txt.record_stream(main)
            ids.record_stream(self._side)
        else:
            vis = self._vision_fwd(images.float(), vsave)
            txt = self._text_fwd(ids, tsave)

        Nf = cfg.fusion_len(T_ctx)
        X = torch.empty(B * Nf, D, device=dev, dtype=torch.float32)
        mask = torch.empty(B * Nf, device=dev, dtype=torch.float32) if cfg.sequential else None
        cmask = sample["context_attention_mask"].to(torch.int64).contiguous() if T_ctx else None
        a = K.AssembleArgs(B, L, Np, T_ctx, D, txt.data_ptr(), vis.data_ptr(),
                           vis.data_ptr() + B * Np * D * 4 if T_ctx else None,
                           self._p("text_token").data_ptr(), self._p("image_token").data_ptr(),
                           self._p("pick_place.fusion.token_type_embeddings.weight").data_ptr(),
                           self._p("context_pos_embedding").data_ptr() if T_ctx else None,
                           K.ptr(cmask), X.data_ptr(), K.ptr(mask))
        K.check(k.lib.bfk_assemble_fwd(C.byref(a), _stream()), "bfk_assemble_fwd")
        k.launches += 1
        del vis, txt

        x = X
        with _Nvtx("bfk.fusion_fwd"):
            if cfg.fusion_model == "crossattention":
                x = self._xattn_fwd(X, B, Nf, mask, fsave)
            for lw in self.fusion_layers:
                x = self._layer_fwd(x, B, Nf, lw, cfg.heads, 1e-5, K.ACT_GELU_ERF, mask, fsave)

        # ---- heatmap tail ----
        heads = len(cfg.heads_out)
        with _Nvtx("bfk.tail_fwd"):
            tail = self._tail_fwd(x, B, Nf, targets, want_argmax, need_grad, pre=pre, X_in=X)
        if need_grad:
            ctx.update(vsave=vsave, tsave=tsave, fsave=fsave, mask=mask, xf=x, Nf=Nf, tail=tail, targets=targets,
                       heads=heads, X_in=X if cfg.compute_mask else None)
        return tail["heat"], tail.get("loss"), tail.get("key"), ctx

    def _decoder_ptr_tables(self, grads=None):
        cfg = self.cfg
        ws, bs, dws, dbs = [], [], [], []
        for pre in cfg.decoder_prefixes:          # policy heads, then the mask head (compute_mask)
            for j in range(5):
                wn, bn = f"{pre}.decoder_net.{2 * j}.weight", f"{pre}.decoder_net.{2 * j}.bias"
                ws.append(self._p(wn).data_ptr()); bs.append(self._p(bn).data_ptr())
                if grads is not None:
                    dws.append(grads[wn].data_ptr()); dbs.append(grads[bn].data_ptr())
        mk = self._ptr_table
        return (mk(ws), mk(bs)) if grads is None else (mk(ws), mk(bs), mk(dws), mk(dbs))

    def _ptr_table(self, ptrs) -> torch.Tensor:
        """Device array of pointers, cached by value (pointer tables are stable in steady state, so no
        per-step H2D copy and CUDA-graph capture stays legal)."""
        key = tuple(ptrs)
        t = self._ptr_tables.get(key)
        if t is None:
            if len(self._ptr_tables) > 256:
                self._ptr_tables.clear()
            host = torch.tensor(list(key), dtype=torch.int64).pin_memory()    # pinned: the upload is a legal graph node
            t = host.to(self.device, non_blocking=True)
            self._ptr_tables[key] = t
            self._ptr_hosts = getattr(self, "_ptr_hosts", [])
            self._ptr_hosts.append(host)                                       # a captured copy re-reads it on every replay
        return t

    def _collapse_fwd(self, dev, off_chain=False):
        """w_eff / b_eff of the linear decoders.  Depends on the decoder weights only, so with `off_chain` it runs on the
        second stream at the START of the forward (four latency-bound CTAs, ~180 us) instead of on the critical path."""
        cfg, k, D = self.cfg, self.k, self.cfg.dim
        heads = len(cfg.decoder_prefixes)
        Wt, bt = self._decoder_ptr_tables()
        w_eff = torch.empty(heads, D, device=dev, dtype=torch.float32)
        b_eff = torch.empty(heads, device=dev, dtype=torch.float32)
        chain = torch.empty(heads, 4, D, device=dev, dtype=torch.float32)
        ca = K.DecoderCollapseArgs(heads, D, Wt.data_ptr(), bt.data_ptr(), w_eff.data_ptr(), b_eff.data_ptr(), chain.data_ptr(),
                                   None, None, None, None)
        ev = None
        cur = torch.cuda.current_stream()
        ws = self._wgrad_streams.get(cur.cuda_stream)
        if ws is None:
            ws = self._wgrad_streams[cur.cuda_stream] = torch.cuda.Stream(device=cur.device)
        if off_chain and self.overlap_wgrad and (self.grad_ready_hook is None or torch.cuda.is_current_stream_capturing()):
            ws.wait_stream(cur)
            with torch.cuda.stream(ws):
                K.check(k.lib.bfk_decoder_collapse_fwd(C.byref(ca), _stream()), "bfk_decoder_collapse_fwd")
                ev = torch.cuda.Event()
                ev.record()
            for t in (w_eff, b_eff, chain):
                t.record_stream(ws)
        else:
            K.check(k.lib.bfk_decoder_collapse_fwd(C.byref(ca), _stream()), "bfk_decoder_collapse_fwd")
        return dict(w_eff=w_eff, b_eff=b_eff, chain=chain, tables=(Wt, bt), ev=ev)

    def _tail_counters(self, n_maps, dev):
        """Zeroed int32 counters the heatmap kernels use to elect the CTA that finishes a reduction (they reset them)."""
        c = getattr(self, "_counters", None)
        if c is None or c.numel() < 1 + n_maps or c.device != dev:
            c = self._counters = torch.zeros(1 + max(n_maps, 4096), device=dev, dtype=torch.int32)
        return c

    def _tail_fwd(self, x, B, Nf, targets, want_argmax, need_grad, pre=None, X_in=None):
        cfg, k, D, dev = self.cfg, self.k, self.cfg.dim, x.device
        heads, Np, g = len(cfg.heads_out), cfg.num_patches, cfg.grid
        H = 16 * g
        lib, st = k.lib, _stream()
        if pre is None:
            pre = self._collapse_fwd(dev)
        elif pre["ev"] is not None:
            torch.cuda.current_stream().wait_event(pre["ev"])
        w_eff, b_eff, chain, (Wt, bt) = pre["w_eff"], pre["b_eff"], pre["chain"], pre["tables"]
        grid = torch.empty(B, heads, Np, device=dev, dtype=torch.float32)
        row0 = Nf - Np
        pa = K.HeadProjectArgs(B, heads, Np, D, x.data_ptr(), Nf * D, row0, w_eff.data_ptr(), b_eff.data_ptr(), grid.data_ptr(),
                               None, Nf, None, None, 0, None, None, None)
        K.check(lib.bfk_head_project_fwd(C.byref(pa), st), "bfk_head_project_fwd")
        ts, tw = self.taps(g)
        counters = self._tail_counters(B * (heads + 1), dev).data_ptr()
        if cfg.compute_mask:
            return self._tail_fwd_masked(X_in, B, Nf, want_argmax, pre, grid, counters)
        heat = torch.empty(heads, B, H, H, device=dev, dtype=torch.float32)
        out = dict(heat=heat, w_eff=w_eff, b_eff=b_eff, chain=chain, grid=grid, tables=(Wt, bt))
        tptr = loss = part = key = None
        if targets is not None:
            tptr = self._ptr_table([t.data_ptr() for t in targets])
            loss = torch.empty(heads + 1, device=dev, dtype=torch.float32)
            out.update(loss=loss, tptr=tptr)
        if want_argmax:
            key = torch.empty(heads * B, device=dev, dtype=torch.int64)
            out["key"] = key
        if targets is not None or want_argmax:
            part = torch.empty(lib.bfk_heatmap_fwd_workspace_floats(B, heads, g), device=dev, dtype=torch.float32)
        ha = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), K.ptr(tptr), None, 0,
                              K.ptr(loss), K.ptr(key), K.ptr(part), counters)
        K.check(lib.bfk_heatmap_fwd(C.byref(ha), st), "bfk_heatmap_fwd")
        k.launches += 3
        # tokens read once (fp32), heatmaps written, targets read when the BCE is fused
        k.work["tail_bytes"] += 4.0 * (B * Np * D + heads * B * H * H * (2 if targets is not None else 1))
        return out

    def _tail_fwd_masked(self, X_in, B, Nf, want_argmax, pre, grid, counters):
        """compute_mask=True (bifold/models/pickplace.py:66-69,78-101): the mask head is a fifth ConvDecoder applied to the
        image rows of the fusion INPUT (image tokens + type embedding: ConcatTransformer adds it in place, fusion.py:52), the pick
        heatmaps are sigmoid(pick) * mask_heatmap.  -> heat [heads + 1, B, H, W], the last map is mask_heatmap."""
        cfg, k, D, dev = self.cfg, self.k, self.cfg.dim, X_in.device
        heads, Np, g = len(cfg.heads_out), cfg.num_patches, cfg.grid
        H = 16 * g
        lib, st = k.lib, _stream()
        w_eff, b_eff = pre["w_eff"], pre["b_eff"]
        ts, tw = self.taps(g)
        grid_m = torch.empty(B, 1, Np, device=dev, dtype=torch.float32)
        pa = K.HeadProjectArgs(B, 1, Np, D, X_in.data_ptr(), Nf * D, Nf - Np, w_eff.data_ptr() + heads * D * 4,
                               b_eff.data_ptr() + heads * 4, grid_m.data_ptr(), None, Nf, None, None, 0, None, None, None)
        K.check(lib.bfk_head_project_fwd(C.byref(pa), st), "bfk_head_project_fwd")
        heat_raw = torch.empty(heads, B, H, H, device=dev, dtype=torch.float32)
        heat = torch.empty(heads + 1, B, H, H, device=dev, dtype=torch.float32)
        for gr, nh, dst in ((grid, heads, heat_raw), (grid_m, 1, heat[heads])):
            ha = K.HeatmapFwdArgs(B, nh, g, gr.data_ptr(), ts.data_ptr(), tw.data_ptr(), dst.data_ptr(), None, None, 0, None, None, None,
                                  counters)
            K.check(lib.bfk_heatmap_fwd(C.byref(ha), st), "bfk_heatmap_fwd")
        ga = K.MaskGateArgs(B, heads, H, H, cfg.pick_head_bits, heat_raw.data_ptr(), heat[heads].data_ptr(), heat.data_ptr())
        K.check(lib.bfk_mask_gate_fwd(C.byref(ga), st), "bfk_mask_gate_fwd")
        out = dict(heat=heat, heat_raw=heat_raw, w_eff=w_eff, b_eff=b_eff, chain=pre["chain"], grid=grid, tables=pre["tables"])
        if want_argmax:
            key = torch.empty(heads * B, device=dev, dtype=torch.int64)
            K.check(lib.bfk_argmax2d(heat.data_ptr(), heads * B, H * H, key.data_ptr(), st), "bfk_argmax2d")
            out["key"] = key
        k.launches += 7
        k.work["tail_bytes"] += 4.0 * (2 * B * Np * D + (3 * heads + 2) * B * H * H)
        return out

    # ------------------------------------------------------------------------------------------------
    # full backward
    # ------------------------------------------------------------------------------------------------
    def trainable_names(self) -> List[str]:
        return [n for n, p in self.params.items() if p.requires_grad]

    def backward(self, ctx, dheat: Optional[torch.Tensor], dloss: Optional[torch.Tensor]) -> Dict[str, torch.Tensor]:
        """dheat: [heads,B,H,W] gradient w.r.t. the heatmaps (or None); dloss: [heads] gradient w.r.t. the fused
        per-head BCE losses (or None).  Returns {param name: fp32 grad}, views of one flat arena."""
        cfg, k, D, T = self.cfg, self.k, self.cfg.dim, self.T
        lib, st = k.lib, _stream()
        B, T_ctx, Nf, heads = ctx["B"], ctx["T"], ctx["Nf"], ctx["heads"]
        Np, L, g = cfg.num_patches, cfg.text_len, cfg.grid
        H = 16 * g
        dev = ctx["xf"].device
        names = self.trainable_names()
        sizes = [self.params[n].numel() for n in names]
        offs, tot = [], 0
        for s in sizes:
            offs.append(tot); tot += (s + 7) // 8 * 8
        arena = torch.empty(tot, device=dev, dtype=torch.float32)
        if os.environ.get("BFK_POISON_ARENA") == "1":       # tests: a trainable gradient that no kernel writes shows up as NaN
            arena.fill_(float("nan"))
        # the flat Adam launch and the DDP all-reduce read the pad slots (<= 7 floats behind a tensor): keep them zero
        pad_key = (tuple(sizes), dev)
        if getattr(self, "_pad_key", None) != pad_key:
            idx = [o + s + j for o, s in zip(offs, sizes) for j in range((s + 7) // 8 * 8 - s)]
            self._pad_idx = torch.tensor(idx, dtype=torch.int64, device=dev) if idx else None
            self._pad_key = pad_key
        if self._pad_idx is not None:
            arena.index_fill_(0, self._pad_idx, 0.0)
        grads = {n: arena[o:o + s].view(self.params[n].shape) for n, o, s in zip(names, offs, sizes)}
        self.last_arena = arena
        tail = ctx["tail"]
        ts, tw = self.taps(g)

        # ---- tail backward: dlogit -> dgrid -> (dX, d w_eff) -> decoder weights ----
        dgrid = torch.empty(B, heads, Np, device=dev, dtype=torch.float32)
        hws = torch.empty(lib.bfk_heatmap_bwd_workspace_floats(B, heads, g), device=dev, dtype=torch.float32)
        counters = self._tail_counters(B * (heads + 1), dev).data_ptr()
        first = True
        heat_p = tail["heat"]              # probabilities the sigmoid backward is taken at
        dgrid_m = None
        if cfg.compute_mask:
            heat_p = tail["heat_raw"]
            if dheat is not None:
                # product rule of pick = sigmoid(pick) * mask (and the mask head's own upstream gradient)
                dheat = dheat.contiguous()
                dheat_raw = torch.empty(heads, B, H, H, device=dev, dtype=torch.float32)
                dmask = torch.empty(B, H, H, device=dev, dtype=torch.float32)
                gb = K.MaskGateBwdArgs(B, heads, H, H, cfg.pick_head_bits, int(cfg.detach_mask), dheat.data_ptr(), 1, heat_p.data_ptr(),
                                       tail["heat"][heads].data_ptr(), dheat_raw.data_ptr(), dmask.data_ptr())
                K.check(lib.bfk_mask_gate_bwd(C.byref(gb), st), "bfk_mask_gate_bwd")
                dgrid_m = torch.empty(B, 1, Np, device=dev, dtype=torch.float32)
                mws = torch.empty(lib.bfk_heatmap_bwd_workspace_floats(B, 1, g), device=dev, dtype=torch.float32)
                mptr = self._ptr_table([dmask.data_ptr()])
                ba = K.HeatmapBwdArgs(B, 1, g, 1, tail["heat"][heads].data_ptr(), None, None, 0, None, 1.0, 0, mptr.data_ptr(),
                                      ts.data_ptr(), tw.data_ptr(), dgrid_m.data_ptr(), mws.data_ptr(), counters)
                K.check(lib.bfk_heatmap_bwd(C.byref(ba), st), "bfk_heatmap_bwd")
                dheat = dheat_raw
                k.launches += 2
        if dloss is not None:
            gscale = dloss.contiguous()     # [heads + 1]: per-head grads then the grad of their sum
            ba = K.HeatmapBwdArgs(B, heads, g, 0, heat_p.data_ptr(), tail["tptr"].data_ptr(), None, 0, gscale.data_ptr(),
                                  1.0 / float(B * H * H), 1, None, ts.data_ptr(), tw.data_ptr(), dgrid.data_ptr(), hws.data_ptr(),
                                  counters)
            K.check(lib.bfk_heatmap_bwd(C.byref(ba), st), "bfk_heatmap_bwd")
            first = False
        if dheat is not None:
            dheat = dheat.contiguous()
            dptr = self._ptr_table([dheat[h].data_ptr() for h in range(heads)])
            tgt = dgrid if first else torch.empty_like(dgrid)
            ba = K.HeatmapBwdArgs(B, heads, g, 1, heat_p.data_ptr(), None, None, 0, None, 1.0, 0, dptr.data_ptr(),
                                  ts.data_ptr(), tw.data_ptr(), tgt.data_ptr(), hws.data_ptr(), counters)
            K.check(lib.bfk_heatmap_bwd(C.byref(ba), st), "bfk_heatmap_bwd")
            if not first:
                dgrid.add_(tgt)   # both a fused-loss and a heatmap gradient (rare): tiny [B,heads,Np] add
            first = False
        if first:
            dgrid.zero_()
        dX = torch.empty(B * Nf, D, device=dev, dtype=torch.float32)
        dX_T = torch.empty(B * Nf, D, device=dev, dtype=T) if T != torch.float32 else None
        nd = len(cfg.decoder_prefixes)
        d_w_eff = torch.empty(nd, D, device=dev, dtype=torch.float32)
        d_b_eff = torch.empty(nd, device=dev, dtype=torch.float32)
        ws = torch.empty(lib.bfk_head_project_bwd_workspace_floats(B, heads, Np, D), device=dev, dtype=torch.float32)
        pa = K.HeadProjectArgs(B, heads, Np, D, ctx["xf"].data_ptr(), Nf * D, Nf - Np, tail["w_eff"].data_ptr(),
                               tail["b_eff"].data_ptr(), None, dgrid.data_ptr(), Nf, dX.data_ptr(), K.ptr(dX_T),
                               _dt(dX_T) if dX_T is not None else 0, d_w_eff.data_ptr(), d_b_eff.data_ptr(), ws.data_ptr())
        K.check(lib.bfk_head_project_bwd(C.byref(pa), st), "bfk_head_project_bwd")
        dX_mask = None
        if cfg.compute_mask:
            if dgrid_m is None:
                d_w_eff[heads:].zero_(); d_b_eff[heads:].zero_()
            else:
                # the mask head reads the fusion INPUT: its token gradient joins the fusion backward's at the assembly
                dX_mask = torch.empty(B * Nf, D, device=dev, dtype=torch.float32)
                ws_m = torch.empty(lib.bfk_head_project_bwd_workspace_floats(B, 1, Np, D), device=dev, dtype=torch.float32)
                pm = K.HeadProjectArgs(B, 1, Np, D, ctx["X_in"].data_ptr(), Nf * D, Nf - Np, tail["w_eff"].data_ptr() + heads * D * 4,
                                       tail["b_eff"].data_ptr() + heads * 4, None, dgrid_m.data_ptr(), Nf, dX_mask.data_ptr(), None, 0,
                                       d_w_eff.data_ptr() + heads * D * 4, d_b_eff.data_ptr() + heads * 4, ws_m.data_ptr())
                K.check(lib.bfk_head_project_bwd(C.byref(pm), st), "bfk_head_project_bwd")
                k.launches += 4
        Wt, bt, dWt, dbt = self._decoder_ptr_tables(grads)
        ca = K.DecoderCollapseArgs(nd, D, Wt.data_ptr(), bt.data_ptr(), None, None, tail["chain"].data_ptr(),
                                   d_w_eff.data_ptr(), d_b_eff.data_ptr(), dWt.data_ptr(), dbt.data_ptr())
        chain_t = tail["chain"]

        def collapse_bwd(ca=ca):      # weight gradients of the decoder chain: a leaf of the backward, off the chain
            K.check(lib.bfk_decoder_collapse_bwd(C.byref(ca), _stream()), "bfk_decoder_collapse_bwd")
        self._off_chain(collapse_bwd, d_w_eff, d_b_eff, chain_t)
        k.launches += 6
        # heatmaps + targets (or dheat) read, tokens read, dX written (fp32 + operand copy)
        k.work["tail_bytes"] += 4.0 * (2 * heads * B * H * H + B * Np * D) + B * Nf * D * (4.0 + (dX_T.element_size() if dX_T is not None else 0))
        if self.grad_ready_hook:
            self._hook_after_off_chain("decoders")

        # ---- fusion transformer ----
        dx, dx_T = dX, (dX_T if dX_T is not None else dX)
        fsave = ctx["fsave"]
        if _Nvtx.on:
            torch.cuda.nvtx.range_push("bfk.fusion_bwd")
        if cfg.fusion_model == "crossattention":
            dx = self._xattn_bwd(dx, dx_T, fsave[0], B, Nf, ctx["mask"], grads)
            dx_T = None
        for l in range(len(self.fusion_layers) - 1, -1, -1):
            dx, dx_T = self._layer_bwd(dx, dx_T, fsave[l], B, Nf, self.fusion_layers[l], cfg.heads, 1e-5, K.ACT_GELU_ERF,
                                       ctx["mask"], grads, trainable=True)
            fsave[l] = None
            if self.grad_ready_hook:
                self._hook_after_off_chain(f"fusion.{l}")
        self._join_off_chain()
        if _Nvtx.on:
            torch.cuda.nvtx.range_pop()
        if dX_mask is not None:
            K.check(lib.bfk_add_inplace(dx.data_ptr(), dX_mask.data_ptr(), dx.numel(), st), "bfk_add_inplace")
            k.launches += 1

        # ---- assembly ----
        need_tower = cfg.lora
        dtxt = torch.empty(B * L, D, device=dev, dtype=torch.float32) if need_tower else None
        dvis = torch.empty((B + B * T_ctx) * Np, D, device=dev, dtype=torch.float32) if need_tower else None
        wsp = torch.empty(2 * 64 * D, device=dev, dtype=torch.float32)
        aa = K.AssembleBwdArgs(B, L, Np, T_ctx, D, dx.data_ptr(), K.ptr(dtxt), K.ptr(dvis),
                               (dvis.data_ptr() + B * Np * D * 4) if (need_tower and T_ctx) else None,
                               grads["text_token"].data_ptr(), grads["image_token"].data_ptr(),
                               grads["pick_place.fusion.token_type_embeddings.weight"].data_ptr(),
                               grads["context_pos_embedding"].data_ptr() if T_ctx else None, wsp.data_ptr())
        K.check(lib.bfk_assemble_bwd(C.byref(aa), st), "bfk_assemble_bwd")
        k.launches += 4
        if cfg.sequential and not T_ctx:
            grads["context_pos_embedding"].zero_()
        if self.grad_ready_hook:
            self.grad_ready_hook("tokens")
        del dx, dx_T

        # ---- towers (only LoRA factors are trainable; gradients flow through the frozen base) ----
        if need_tower:
            tp, vp = self.tw_prefix + "text_model", self.tw_prefix + "vision_model"
            if self.overlap_towers and self._side is not None:
                main = torch.cuda.current_stream()
                self._side.wait_stream(main)
                dtxt.record_stream(self._side)
                arena.record_stream(self._side)
                with torch.cuda.stream(self._side):
                    self._tower_bwd(dtxt, ctx["tsave"], self.text_layers, B, L, tp + ".final_layer_norm.weight", grads)
                    if self.grad_ready_hook:
                        self._hook_after_off_chain("text")
                    self._join_off_chain()
                with _Nvtx("bfk.vision_tower_bwd"):
                    self._tower_bwd(dvis, ctx["vsave"], self.vis_layers, B + B * T_ctx, Np, vp + ".post_layernorm.weight", grads)
                main.wait_stream(self._side)
            else:
                self._tower_bwd(dtxt, ctx["tsave"], self.text_layers, B, L, tp + ".final_layer_norm.weight", grads)
                if self.grad_ready_hook:
                    self._hook_after_off_chain("text")
                self._tower_bwd(dvis, ctx["vsave"], self.vis_layers, B + B * T_ctx, Np, vp + ".post_layernorm.weight", grads)
            if self.grad_ready_hook:
                self._hook_after_off_chain("vision")
            self._join_off_chain()
        return grads
