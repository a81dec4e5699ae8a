"""GPU: each C-ABI kernel against a plain torch fp32 reference of the same op (through the C ABI)."""
import ctypes as C
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kern():
    from bifold_b200.engine import Kernels
    return Kernels()


def _gemm_ref(A, B, ta, tb, bias, act, aux_in, residual, res_mod, alpha):
    a = A.float().t() if ta else A.float()
    b = B.float() if tb else B.float().t()
    out = alpha * (a @ b)
    if bias is not None:
        out = out + bias
    pre = out.clone()
    if act == 1:
        out = F.gelu(out, approximate="tanh")
    elif act == 2:
        out = F.gelu(out)
    elif act in (3, 4):
        h = aux_in.float().requires_grad_(True)
        y = F.gelu(h, approximate="tanh") if act == 3 else F.gelu(h)
        (dg,) = torch.autograd.grad(y.sum(), h)
        out = out * dg
    if residual is not None:
        r = residual
        if res_mod:
            r = residual[torch.arange(out.shape[0], device=out.device) % res_mod]
        out = out + r
    return out, pre


GEMM_CASES = [
    # M, N, K, ta, tb, bias, act, residual, res_mod, out_bf16
    (200, 136, 72, 0, 0, True, 0, False, 0, False),
    (200, 136, 72, 0, 0, True, 1, False, 0, False),
    (129, 256, 64, 0, 0, True, 2, True, 0, False),
    (300, 784, 784, 0, 1, False, 0, False, 0, False),
    (256, 128, 520, 1, 1, False, 0, False, 0, False),
    (130, 264, 200, 0, 1, False, 3, False, 0, False),
    (130, 264, 200, 0, 1, False, 4, False, 0, False),
    (392, 768, 768, 0, 0, True, 0, True, 196, False),
    (1000, 2304, 784, 0, 0, True, 0, False, 0, True),
    (2304, 768, 1048, 1, 1, False, 0, False, 0, False),
    (64, 16, 4304, 0, 0, False, 0, False, 0, False),
    (768, 8, 4100, 1, 1, False, 0, False, 0, False),     # LoRA dB: split-K, N = 8
    (8, 768, 4100, 1, 1, False, 0, False, 0, False),     # LoRA dA: split-K, M = 8
    (300, 768, 16, 0, 1, False, 0, False, 0, True),      # LoRA dh contribution: K = 16
    (300, 16, 768, 0, 0, False, 0, False, 0, True),      # LoRA down-projection: N = 16
    (768, 768, 8200, 1, 1, False, 0, False, 0, False),   # wgrad, few tiles: split-K
]


@pytest.mark.parametrize("dtype,backend", [("f32", 1), ("bf16", 1), ("bf16", 2)])
@pytest.mark.parametrize("case", GEMM_CASES)
def test_gemm(kern, case, dtype, backend):
    M, N, Kd, ta, tb, use_bias, act, use_res, res_mod, out_bf16 = case
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(hash(case) % 2**31)
    T = torch.float32 if dtype == "f32" else torch.bfloat16
    A = (torch.randn((Kd, M) if ta else (M, Kd), generator=g) * 0.5).to(dev).to(T)
    B = (torch.randn((Kd, N) if tb else (N, Kd), generator=g) * 0.5).to(dev).to(T)
    bias = torch.randn(N, generator=g).to(dev) if use_bias else None
    res = torch.randn(res_mod or M, N, generator=g).to(dev) if use_res else None
    aux_in = (torch.randn(M, N, generator=g)).to(dev).to(T) if act in (3, 4) else None
    out_T = torch.bfloat16 if (out_bf16 and dtype == "bf16") else torch.float32
    D = torch.full((M, N), float("nan"), device=dev, dtype=out_T)
    aux = aux_in if act in (3, 4) else (torch.full((M, N), float("nan"), device=dev, dtype=out_T) if act in (1, 2) else None)
    kern.gemm(A, B, D, M, N, Kd, A.shape[1], B.shape[1], N, trans_a=bool(ta), trans_b=bool(tb), bias=bias, act=act,
              aux=aux, ld_aux=N, residual=res, ldr=N, res_row_mod=res_mod, alpha=0.75, backend=backend)
    torch.cuda.synchronize()
    ref, pre = _gemm_ref(A, B, ta, tb, bias, act, aux_in, res, res_mod, 0.75)
    tol = 2e-4 if dtype == "f32" else (2e-2 if out_T == torch.bfloat16 else 2e-3)
    scale = ref.abs().max().item()
    err = (D.float() - ref).abs().max().item()
    assert math.isfinite(err) and err <= tol * scale, f"gemm {case} {dtype} backend={backend}: err {err:.3e} scale {scale:.3e}"
    if act in (1, 2):
        err2 = (aux.float() - pre).abs().max().item()
        assert err2 <= tol * pre.abs().max().item(), f"pre-activation store: {err2:.3e}"


SPECIAL_CASES = [
    # ta, tb, bias, act, residual, out_bf16   (alpha = 1, N % 32 == 0, enough tiles for the CTA-pair kernel)
    (0, 0, True, 0, False, True),      # EPI_STORE_BF16
    (0, 1, False, 0, False, True),     # EPI_STORE_BF16, dgrad layout
    (0, 0, True, 0, True, False),      # EPI_RES_F32
    (0, 0, True, 1, False, True),      # EPI_GELU_TANH_BF16 (+ pre-activation)
    (0, 0, True, 2, False, True),      # EPI_GELU_ERF_BF16
    (0, 1, False, 3, False, True),     # EPI_DGELU_TANH_BF16
    (0, 1, False, 4, False, True),     # EPI_DGELU_ERF_BF16
    (0, 0, True, 5, False, True),      # EPI_GELUG_TANH_BF16: value + derivative
    (0, 0, True, 6, False, True),      # EPI_GELUG_ERF_BF16
    (0, 1, False, 7, False, True),     # EPI_MULAUX_BF16
    (1, 1, False, 0, False, False),    # EPI_STORE_F32, wgrad layout
]


@pytest.mark.parametrize("case", SPECIAL_CASES)
@pytest.mark.parametrize("shape", [(1280, 2048, 320), (1304, 2048, 328), (2560, 784, 328), (2560, 1136, 320), (1280, 4304, 328)])
def test_gemm_specialised_epilogues(kern, case, shape):
    """The compile-time specialised tcgen05 epilogues (alpha == 1, N % 16 == 0) against fp32 torch; M = 1304 leaves a
    ragged last row block, K = 328 a ragged last K block; N = 784 / 1136 end in a thin last n-tile (16 / 112 live columns)
    whose MMA runs with a reduced N; N = 784 / 1136 / 4304 (the so400m MLP width) have N % 32 == 16: the last 32-column chunk
    is half live (guarded vector loads / stores)."""
    ta, tb, use_bias, act, use_res, out_bf16 = case
    M, N, Kd = shape
    dev = "cuda"
    g = torch.Generator(device="cpu").manual_seed(1000 * act + 10 * ta + tb + M)
    T = torch.bfloat16
    A = (torch.randn((Kd, M) if ta else (M, Kd), generator=g) * 0.25).to(dev).to(T)
    B = (torch.randn((Kd, N) if tb else (N, Kd), generator=g) * 0.25).to(dev).to(T)
    bias = torch.randn(N, generator=g).to(dev) if use_bias else None
    res = torch.randn(M, N, generator=g).to(dev) if use_res else None
    reads_aux = act in (3, 4, 7)
    aux_in = torch.randn(M, N, generator=g).to(dev).to(T) if reads_aux else None
    out_T = torch.bfloat16 if out_bf16 else torch.float32
    D = torch.full((M, N), float("nan"), device=dev, dtype=out_T)
    aux = aux_in if reads_aux else (torch.full((M, N), float("nan"), device=dev, dtype=out_T) if act in (1, 2, 5, 6) else None)
    kern.gemm(A, B, D, M, N, Kd, A.shape[1], B.shape[1], N, trans_a=bool(ta), trans_b=bool(tb), bias=bias, act=act,
              aux=aux, ld_aux=N, residual=res, ldr=N, backend=2)
    torch.cuda.synchronize()
    a = A.float().t() if ta else A.float()
    b = B.float() if tb else B.float().t()
    pre = a @ b + (bias if bias is not None else 0)
    tanh_form = act in (1, 3, 5)

    def gelu_and_grad(h):
        h = h.clone().requires_grad_(True)
        y = F.gelu(h, approximate="tanh") if tanh_form else F.gelu(h)
        (dg,) = torch.autograd.grad(y.sum(), h)
        return y.detach(), dg

    aux_ref = None
    if act == 0:
        ref = pre + (res if res is not None else 0)
    elif act in (1, 2):
        ref, aux_ref = gelu_and_grad(pre)[0], pre
    elif act in (3, 4):
        ref = pre * gelu_and_grad(aux_in.float())[1]
    elif act in (5, 6):
        ref, aux_ref = gelu_and_grad(pre)
    else:
        ref = pre * aux_in.float()
    tol = 2e-2 if out_bf16 else 2e-3
    err = (D.float() - ref).abs().max().item()
    assert math.isfinite(err) and err <= tol * ref.abs().max().item(), f"{case} {shape}: err {err:.3e}"
    if aux_ref is not None:
        err2 = (aux.float() - aux_ref).abs().max().item()
        assert math.isfinite(err2) and err2 <= tol * max(aux_ref.abs().max().item(), 1.0), f"{case} {shape}: aux err {err2:.3e}"


@pytest.mark.parametrize("case", [(0, 0, True, 0, True, False), (0, 0, True, 6, False, True), (0, 1, False, 7, False, True), (0, 0, True, 0, False, True)])
def test_gemm_ragged_half_chunk_stays_inside_the_row(kern, case):
    """N % 32 == 16 (so400m MLP width): the fused epilogues store the live half of the last 32-column chunk only.  The outputs
    live in wider buffers (ld = N + 48) pre-filled with a sentinel: nothing beyond column N may change."""
    ta, tb, use_bias, act, use_res, out_bf16 = case
    M, N, Kd, pad = 1280, 4304, 192, 48
    dev, T = "cuda", torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(77 + act)
    A = (torch.randn(M, Kd, generator=g) * 0.25).to(dev).to(T)
    B = (torch.randn((Kd, N) if tb else (N, Kd), generator=g) * 0.25).to(dev).to(T)
    bias = torch.randn(N, generator=g).to(dev) if use_bias else None
    res = torch.randn(M, N + pad, generator=g).to(dev) if use_res else None
    out_T = torch.bfloat16 if out_bf16 else torch.float32
    D = torch.full((M, N + pad), 777.0, device=dev, dtype=out_T)
    aux = None
    if act == 7:
        aux = torch.randn(M, N + pad, generator=g).to(dev).to(T)
    elif act == 6:
        aux = torch.full((M, N + pad), 777.0, device=dev, dtype=T)
    kern.gemm(A, B, D, M, N, Kd, Kd, B.shape[1], N + pad, trans_b=bool(tb), bias=bias, act=act, aux=aux, ld_aux=N + pad,
              residual=res, ldr=N + pad, backend=2)
    torch.cuda.synchronize()
    assert (D[:, N:] == 777.0).all(), "store beyond column N"
    if act == 6:
        assert (aux[:, N:] == 777.0).all(), "derivative store beyond column N"
    pre = A.float() @ (B.float() if tb else B.float().t()) + (bias if bias is not None else 0)
    if act == 0:
        ref = pre + (res[:, :N] if res is not None else 0)
    elif act == 6:
        ref = F.gelu(pre)
    else:
        ref = pre * aux[:, :N].float()
    err = (D[:, :N].float() - ref).abs().max().item()
    assert err <= (2e-2 if out_bf16 else 2e-3) * ref.abs().max().item(), err


def test_gemm_accumulate(kern):
    M, N, Kd = 130, 72, 40
    A = torch.randn(M, Kd, device="cuda"); B = torch.randn(N, Kd, device="cuda")
    D = torch.randn(M, N, device="cuda"); D0 = D.clone()
    kern.gemm(A, B, D, M, N, Kd, Kd, Kd, N, accumulate=True)
    assert torch.allclose(D, D0 + A @ B.t(), atol=1e-4, rtol=1e-4)


def test_colsum(kern):
    for T in (torch.float32, torch.bfloat16):
        for M, N in ((1000, 300), (1001, 768), (515, 3072)):      # 300: scalar kernel; 768 / 3072: 16-byte vector kernel
            X = torch.randn(M, N, device="cuda").to(T)
            out = torch.empty(N, device="cuda")
            kern.colsum(X, M, N, N, out)
            assert torch.allclose(out, X.float().sum(0), atol=2e-3, rtol=1e-4), (T, M, N)


@pytest.mark.parametrize("D,ydt", [(64, torch.float32), (768, torch.bfloat16), (1152, torch.float32), (768, torch.float32)])
def test_layernorm_fwd_bwd(kern, D, ydt):
    M = 777
    x = (torch.randn(M, D, device="cuda") * 2 + 0.3).requires_grad_(True)
    gam = (1 + 0.1 * torch.randn(D, device="cuda")).requires_grad_(True)
    bet = (0.1 * torch.randn(D, device="cuda")).requires_grad_(True)
    ld = D + 16
    y = torch.zeros(M, ld, device="cuda", dtype=ydt)
    mean = torch.empty(M, device="cuda"); rstd = torch.empty(M, device="cuda")
    kern.layernorm_fwd(x.detach(), M, D, gam.detach(), bet.detach(), 1e-6, y, ld, mean, rstd)
    ref = F.layer_norm(x, (D,), gam, bet, 1e-6)
    tol = 1e-5 if ydt == torch.float32 else 1e-2
    assert (y[:, :D].float() - ref).abs().max().item() <= tol * ref.abs().max().item()
    dy = torch.randn(M, D, device="cuda")
    dres = torch.randn(M, D, device="cuda")
    gx, gg, gb = torch.autograd.grad(ref, (x, gam, bet), dy)
    dy_in = torch.zeros(M, ld, device="cuda", dtype=ydt); dy_in[:, :D] = dy.to(ydt)
    dx = torch.empty(M, D, device="cuda"); dx2 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    dgam = torch.empty(D, device="cuda"); dbet = torch.empty(D, device="cuda")
    kern.layernorm_bwd(dy_in, ld, x.detach(), M, D, mean, rstd, gam.detach(), dres, dx, dx2, dgam, dbet)
    tol = 2e-4 if ydt == torch.float32 else 2e-2
    assert (dx - (gx + dres)).abs().max().item() <= tol * (gx + dres).abs().max().item()
    assert (dx2.float() - dx).abs().max().item() <= 1e-2 * dx.abs().max().item()
    assert (dgam - gg).abs().max().item() <= tol * gg.abs().max().item()
    assert (dbet - gb).abs().max().item() <= tol * gb.abs().max().item()
    dx_b = torch.empty(M, D, device="cuda")
    kern.layernorm_bwd(dy_in, ld, x.detach(), M, D, mean, rstd, gam.detach(), None, dx_b)
    assert (dx_b - gx).abs().max().item() <= tol * gx.abs().max().item()


@pytest.mark.parametrize("D", [768, 1152, 512])   # 768 / 1152: tensor-core LoRA down-projection; 512: in-LayerNorm dot products
def test_layernorm_fwd_fused_dropout_copy(kern, D):
    """The optional second LN output equals bfk_dropout_copy applied to the first (same element hash)."""
    from bifold_b200 import _cabi as K
    M, p, seed = 333, 0.1, 1234567
    x = torch.randn(M, D, device="cuda")
    gam = torch.randn(D, device="cuda"); bet = torch.randn(D, device="cuda")
    y = torch.zeros(M, D + 16, device="cuda", dtype=torch.bfloat16)
    y2 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    kern.layernorm_fwd(x, M, D, gam, bet, 1e-6, y, D + 16, y2=y2, ldy2=D, y2_p=p, y2_seed=seed)
    ref = torch.empty_like(y2)
    K.check(kern.lib.bfk_dropout_copy(y.data_ptr(), D + 16, ref.data_ptr(), D, K.BF16, M, D, p, seed,
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert torch.equal(y2, ref)
    frac = (y2 == 0).float().mean().item()
    assert 0.07 < frac < 0.13
    # fused LoRA down-projection: u = drop(y) A^T into the 16 columns behind y (with and without dropout)
    A = (torch.randn(16, D, device="cuda") * 0.1).bfloat16()
    for p_drop in (p, 0.0):
        yb = torch.zeros(M, D + 16, device="cuda", dtype=torch.bfloat16)
        y2b = torch.empty(M, D, device="cuda", dtype=torch.bfloat16) if p_drop > 0 else None
        kern.layernorm_fwd(x, M, D, gam, bet, 1e-6, yb, D + 16, y2=y2b, ldy2=D, y2_p=p_drop, y2_seed=seed,
                           lora_a=A, lora_u=yb[:, D:], ld_u=D + 16)
        torch.cuda.synchronize()
        assert torch.equal(yb[:, :D], y[:, :D])
        src = y2b if p_drop > 0 else yb[:, :D]
        u_ref = src.float() @ A.float().t()
        err = (yb[:, D:].float() - u_ref).abs().max().item()
        assert err <= 1e-2 * u_ref.abs().max().item(), (p_drop, err)


@pytest.mark.parametrize("M,D,r", [(1000, 64, 8), (12544, 768, 8), (777, 128, 4)])
def test_lora_grads_fused(kern, M, D, r):
    """bfk_lora_grads (two split-K tcgen05 GEMMs + one reduction) against fp32 matmuls of the same bf16 inputs."""
    from bifold_b200 import _cabi as K
    g = torch.Generator().manual_seed(M + D)
    Dk = D + 16
    dqkv = torch.randn(M, 3 * D, generator=g).cuda().bfloat16()
    h1 = torch.randn(M, Dk, generator=g).cuda().bfloat16()      # u = h1[:, D:]
    dh1 = torch.randn(M, Dk, generator=g).cuda().bfloat16()     # du = dh1[:, D:]
    x = torch.randn(M, D, generator=g).cuda().bfloat16()
    outs = [torch.full((D, r), float("nan"), device="cuda"), torch.full((D, r), float("nan"), device="cuda"),
            torch.full((r, D), float("nan"), device="cuda"), torch.full((r, D), float("nan"), device="cuda")]
    ws = torch.empty(kern.lib.bfk_lora_grads_workspace_bytes(D) // 4, device="cuda")
    scale = 2.0
    a = K.LoraGradsArgs(M, D, r, dqkv.data_ptr(), 3 * D, h1.data_ptr() + 2 * D, Dk, dh1.data_ptr() + 2 * D, Dk, x.data_ptr(), D,
                        scale, outs[0].data_ptr(), outs[1].data_ptr(), outs[2].data_ptr(), outs[3].data_ptr(), ws.data_ptr(),
                        ws.numel() * 4)
    K.check(kern.lib.bfk_lora_grads(C.byref(a), torch.cuda.current_stream().cuda_stream), "bfk_lora_grads")
    torch.cuda.synchronize()
    u, du = h1[:, D:].float(), dh1[:, D:].float()
    refs = [scale * dqkv[:, :D].float().t() @ u[:, :r], scale * dqkv[:, 2 * D:].float().t() @ u[:, 8:8 + r],
            du[:, :r].t() @ x.float(), du[:, 8:8 + r].t() @ x.float()]
    for o, ref in zip(outs, refs):
        assert torch.isfinite(o).all()
        assert (o - ref).abs().max().item() <= 2e-3 * ref.abs().max().item() + 1e-3


def _attn_ref(qkv, B, N, H, hd, mask, scale):
    q, k, v = qkv.float().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    dots = q @ k.transpose(-1, -2) * scale
    if mask is not None:
        dots = dots.masked_fill(mask[:, None, :, None] == 0, -100000)
    return (torch.softmax(dots, -1) @ v).transpose(1, 2).reshape(B * N, H * hd)


# backend 0 = default dispatch (generic flash forward, merged persistent backward where eligible), 3 = persistent forward too;
# the (20, 196, 12, 64) / (13, 262, 16, 48) / (40, 64, 12, 64) cases give every persistent CTA several items (barrier phases wrap)
@pytest.mark.parametrize("T,backend", [(torch.float32, 1), (torch.bfloat16, 1), (torch.bfloat16, 0), (torch.bfloat16, 3)])
@pytest.mark.parametrize("B,N,H,hd,masked", [(2, 70, 3, 16, True), (3, 197, 2, 64, False), (2, 262, 4, 48, True),
                                              (1, 300, 2, 72, True), (2, 64, 12, 64, False), (1, 853, 2, 48, True),
                                              (20, 196, 12, 64, True), (13, 262, 16, 48, True), (40, 64, 12, 64, False)])
def test_attention_fwd_bwd(kern, T, backend, B, N, H, hd, masked):
    if backend == 1 and B * H * N * N > 4_000_000:
        pytest.skip("SIMT checker: small shapes only")
    kern.backend = backend
    try:
        g = torch.Generator().manual_seed(B * 1000 + N)
        qkv = (torch.randn(B * N, 3 * H * hd, generator=g)).cuda().to(T)
        mask = None
        if masked:
            mask = torch.ones(B, N)
            mask[0, 5:5 + N // 3] = 0
            mask = mask.cuda()
        out = torch.empty(B * N, H * hd, device="cuda", dtype=T)
        lse = torch.empty(B * H * N, device="cuda")
        scale = hd ** -0.5
        kern.attention_fwd(qkv, out, lse, B, N, H, hd, scale, mask)
        q32 = qkv.float().requires_grad_(True)
        ref = _attn_ref(q32, B, N, H, hd, mask, scale)
        tol = 2e-4 if T == torch.float32 else 2e-2
        err = (out.float() - ref).abs().max().item()
        assert math.isfinite(err) and err <= tol * ref.abs().max().item(), f"attention fwd err {err:.3e}"
        dout = torch.randn(B * N, H * hd, generator=g).cuda().to(T)
        (gref,) = torch.autograd.grad(ref, q32, dout.float())
        dqkv = torch.full_like(qkv, float("nan"))
        kern.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, scale, mask)
        err = (dqkv.float() - gref).abs().max().item()
        tol = 5e-4 if T == torch.float32 else 4e-2
        assert math.isfinite(err) and err <= tol * gref.abs().max().item(), f"attention bwd err {err:.3e} vs {gref.abs().max().item():.3e}"
    finally:
        kern.backend = 0


def test_adamw_matches_torch(kern):
    from bifold_b200 import _cabi as K
    n = 100003
    for decoupled, wd in ((0, 0.0), (0, 0.1), (1, 0.01)):
        p = torch.randn(n, device="cuda"); g = torch.randn(n, device="cuda")
        ref = p.clone().requires_grad_(True)
        opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref], lr=1e-3, weight_decay=wd, foreach=False)
        m = torch.zeros(n, device="cuda"); v = torch.zeros(n, device="cuda")
        sh = torch.empty(n, device="cuda", dtype=torch.bfloat16)
        for step in range(1, 4):
            ref.grad = g.clone() * step
            opt.step()
            a = K.AdamWArgs(n, p.data_ptr(), (g * step).data_ptr(), m.data_ptr(), v.data_ptr(), sh.data_ptr(), 1e-3, 0.9, 0.999,
                            1e-8, wd, decoupled, step, 1.0)
            gg = g * step
            a.g = gg.data_ptr()
            K.check(kern.lib.bfk_adamw(C.byref(a), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
        assert torch.allclose(p, ref.detach(), atol=1e-6, rtol=1e-5), (decoupled, wd)
        assert torch.allclose(sh.float(), p, atol=1e-2, rtol=1e-2)


def test_heatmap_fwd_matches_literal_decoder(kern):
    """Collapsed decoder + composite upsample + sigmoid + BCE + argmax against the literal
    ConvDecoder pyramid of the oracle (bifold/models/modules.py:33-74)."""
    from oracle import policy_oracle as po
    from oracle.synth import make_sample, make_state_dict, tiny_spec
    from tests.util import build_model, to_device
    spec = tiny_spec(name="siglip", is_bimanual=True, lora=False)
    model, sd = build_model(spec, "fp32")
    sample = make_sample(spec, 3)
    eng = model._ensure_engine()
    eng.prepare(torch.device("cuda"))
    eng.training = False
    D, Np, B = spec.dim, spec.num_patches, 3
    Nf = spec.fusion_len()
    xf = torch.randn(B * Nf, D, device="cuda")
    targets = [to_device(sample)[f"{h}_heatmap"].contiguous() for h in spec.heads_out]
    tail = eng._tail_fwd(xf, B, Nf, targets, True, False)
    fused = xf.view(B, Nf, D)[:, Nf - Np - 1:, :].cpu()
    logits = po.pick_place_logits(sd, fused, spec)
    for i, h in enumerate(spec.heads_out):
        ref = logits[h].sigmoid()
        got = tail["heat"][i].cpu()
        assert (got - ref).abs().max().item() < 2e-6, h
        loss_ref = F.binary_cross_entropy(ref, sample[f"{h}_heatmap"])
        assert abs(tail["loss"][i].item() - loss_ref.item()) <= 1e-5 * abs(loss_ref.item()) + 1e-7
        key = tail["key"].view(len(spec.heads_out), B).cpu().numpy().astype(np.uint64)
        for b in range(B):
            idx = 0xFFFFFFFF - int(key[i, b] & np.uint64(0xFFFFFFFF))
            assert idx == int(got[b].numpy().argmax()), (h, b)
    total = sum(tail["loss"][i].item() for i in range(len(spec.heads_out)))
    assert abs(tail["loss"][len(spec.heads_out)].item() - total) < 1e-5


def test_argmax2d_first_max_and_action_gate(kern):
    from bifold_b200 import _cabi as K
    B, H, W, heads = 5, 64, 64, 4
    g = torch.Generator().manual_seed(5)
    heat = torch.rand(heads, B, H, W, generator=g) * 0.5
    heat[0, 0, 10, 20] = 0.9; heat[0, 0, 30, 5] = 0.9        # tie -> first (row-major)
    heat[2, 1] *= 0.01                                      # right arm low confidence on sample 1
    heat[0, 2] *= 0.001; heat[2, 2] *= 0.0001               # both below threshold on sample 2
    heat = heat.cuda().contiguous()
    key = torch.empty(heads * B, device="cuda", dtype=torch.int64)
    K.check(kern.lib.bfk_argmax2d(heat.data_ptr(), heads * B, H * W, key.data_ptr(), torch.cuda.current_stream().cuda_stream))
    idx = 0xFFFFFFFF - (key.cpu().numpy().astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
    ref_idx = heat.view(heads * B, -1).cpu().numpy().argmax(1)
    np.testing.assert_array_equal(idx, ref_idx)
    mask = torch.zeros(B, 1, H, W); mask[:, :, 20:40, 25:50] = 1
    from oracle import policy_oracle as po
    from oracle.synth import PolicySpec
    spec = PolicySpec(is_bimanual=True)
    out = {f"{h}_heatmap": heat[i].cpu() for i, h in enumerate(spec.heads_out)}
    ref = po.get_action(out, {"mask": mask}, spec, threshold=0.01)
    xy = torch.empty(heads, B, 2, device="cuda", dtype=torch.int32); conf = torch.empty(heads, B, device="cuda")
    mk = mask.reshape(B, H, W).cuda().contiguous()
    a = K.ActionArgs(B, heads, H, W, 2, heat.data_ptr(), key.data_ptr(), mk.data_ptr(), (C.c_int32 * 2)(0, 2), (C.c_int32 * 2)(1, 3),
                     0.01, xy.data_ptr(), conf.data_ptr())
    K.check(kern.lib.bfk_action(C.byref(a), torch.cuda.current_stream().cuda_stream))
    xy = xy.cpu().numpy()
    for i, h in enumerate(spec.heads_out):
        np.testing.assert_array_equal(xy[i], ref[h].astype(np.int64), err_msg=h)
    assert (xy[2, 1] == -1).all() and (xy[3, 1] == -1).all()      # gated arm
    assert (xy[0, 2] != -1).all() and (xy[2, 2] == -1).all()      # more confident arm kept


def test_bce_prob_clamp_matches_torch(kern):
    from bifold_b200 import ops  # noqa: F401
    p = torch.tensor([0.0, 1.0, 0.5, 1e-30, 1 - 1e-7, 0.3], device="cuda").repeat(100).requires_grad_(True)
    t = torch.tensor([1.0, 0.0, 0.5, 1.0, 0.2, 0.0], device="cuda").repeat(100)
    loss = torch.ops.bfk.bce_prob(p, t, None)
    ref_p = p.detach().clone().requires_grad_(True)
    ref = F.binary_cross_entropy(ref_p, t)
    assert abs(loss.item() - ref.item()) <= 1e-5 * ref.item()
    loss.backward(); ref.backward()
    assert torch.allclose(p.grad, ref_p.grad, rtol=1e-4, atol=1e-6)


def _up16(grid):
    x = grid
    for _ in range(4):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return x


@pytest.mark.parametrize("g,B,heads", [(2, 3, 2), (3, 2, 4), (14, 5, 4), (24, 2, 2), (27, 2, 4)])
@pytest.mark.parametrize("use_counters", [True, False])
def test_heatmap_fwd_bwd_vs_torch(kern, g, B, heads, use_counters):
    """bfk_heatmap_fwd / bfk_heatmap_bwd (all band / group geometries, fused and two-launch reductions, target mask, generic
    upstream gradient) against four torch x2 bilinear stages + sigmoid + BCELoss (bifold/models/modules.py:57-74,
    bifold/losses/__init__.py:76-83) and autograd."""
    from bifold_b200 import _cabi as K
    lib, dev = kern.lib, "cuda"
    H = 16 * g
    gen = torch.Generator().manual_seed(g * 100 + B)
    grid = (torch.randn(B, heads, g, g, generator=gen) * 3).to(dev)
    grid[0, 0, 0, 0] = 30.0          # saturated pixels: the literal (clamped) BCE path
    targets = [torch.rand(B, H, H, generator=gen).to(dev) for _ in range(heads)]
    tmask = (torch.rand(B, H, H, generator=gen) > 0.5).float().to(dev)
    ts_h = torch.empty(H, dtype=torch.int32); tw_h = torch.empty(H, 3)
    K.check(lib.bfk_upsample16_taps(g, ts_h.data_ptr(), tw_h.data_ptr()))
    ts, tw = ts_h.to(dev), tw_h.to(dev)
    tptr = torch.tensor([t.data_ptr() for t in targets], dtype=torch.int64).to(dev)
    counters = torch.zeros(1 + heads * B, device=dev, dtype=torch.int32)
    cptr = counters.data_ptr() if use_counters else None
    st = torch.cuda.current_stream().cuda_stream
    for mask_heads in (0, 0b01):
        heat = torch.empty(heads, B, H, H, device=dev)
        loss = torch.empty(heads + 1, device=dev); key = torch.empty(heads * B, device=dev, dtype=torch.int64)
        part = torch.empty(lib.bfk_heatmap_fwd_workspace_floats(B, heads, g), device=dev)
        fa = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), tptr.data_ptr(),
                              tmask.data_ptr() if mask_heads else None, mask_heads, loss.data_ptr(), key.data_ptr(), part.data_ptr(), cptr)
        for _ in range(2):     # twice: the counters must come back to zero
            K.check(lib.bfk_heatmap_fwd(C.byref(fa), st))
        gref = grid.clone().requires_grad_(True)
        p_ref = _up16(gref).sigmoid()                              # [B, heads, H, H]
        assert (heat.permute(1, 0, 2, 3) - p_ref).abs().max().item() < 2e-6
        tt = [t * tmask if (mask_heads >> h) & 1 else t for h, t in enumerate(targets)]
        losses = [F.binary_cross_entropy(p_ref[:, h], tt[h]) for h in range(heads)]
        for h in range(heads):
            assert abs(loss[h].item() - losses[h].item()) <= 2e-5 * abs(losses[h].item()) + 1e-7, (h, loss[h].item(), losses[h].item())
        assert abs(loss[heads].item() - sum(l.item() for l in losses)) <= 1e-4 * sum(l.item() for l in losses)
        idx = 0xFFFFFFFF - (key.cpu().numpy().astype(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.int64)
        np.testing.assert_array_equal(idx, heat.view(heads * B, -1).cpu().numpy().argmax(1))
        # argmax-only call without a workspace (atomic keys)
        key2 = torch.empty_like(key)
        fa2 = K.HeatmapFwdArgs(B, heads, g, grid.data_ptr(), ts.data_ptr(), tw.data_ptr(), heat.data_ptr(), None, None, 0, None,
                               key2.data_ptr(), None, None)
        K.check(lib.bfk_heatmap_fwd(C.byref(fa2), st))
        assert torch.equal(key, key2)
        # backward, mode 0: d(sum_h w_h loss_h + w_tot total)/dgrid
        gscale = torch.tensor([0.5 + h for h in range(heads)] + [2.0], device=dev)
        (sum((gscale[h] + gscale[heads]) * losses[h] for h in range(heads))).backward()
        dgrid = torch.empty(B, heads, g * g, device=dev)
        ws = torch.empty(lib.bfk_heatmap_bwd_workspace_floats(B, heads, g), device=dev)
        ba = K.HeatmapBwdArgs(B, heads, g, 0, heat.data_ptr(), tptr.data_ptr(), tmask.data_ptr() if mask_heads else None, mask_heads,
                              gscale.data_ptr(), 1.0 / (B * H * H), 1, None, ts.data_ptr(), tw.data_ptr(), dgrid.data_ptr(), ws.data_ptr(), cptr)
        for _ in range(2):
            K.check(lib.bfk_heatmap_bwd(C.byref(ba), st))
        ref = gref.grad.view(B, heads, g * g)
        assert (dgrid - ref).abs().max().item() <= 2e-4 * ref.abs().max().item() + 1e-9
    # backward, mode 1: arbitrary upstream gradient w.r.t. the heatmaps (one head without gradient)
    dheat = [torch.randn(B, H, H, generator=gen).to(dev) for _ in range(heads)]
    dptr = torch.tensor([0 if h == 1 else d.data_ptr() for h, d in enumerate(dheat)], dtype=torch.int64).to(dev)
    gref = grid.clone().requires_grad_(True)
    p_ref = _up16(gref).sigmoid()
    sum((p_ref[:, h] * dheat[h]).sum() for h in range(heads) if h != 1).backward()
    ba = K.HeatmapBwdArgs(B, heads, g, 1, heat.data_ptr(), None, None, 0, None, 1.0, 0, dptr.data_ptr(), ts.data_ptr(), tw.data_ptr(),
                          dgrid.data_ptr(), ws.data_ptr(), cptr)
    K.check(lib.bfk_heatmap_bwd(C.byref(ba), st))
    ref = gref.grad.view(B, heads, g * g)
    assert (dgrid - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    assert int(counters.abs().sum().item()) == 0


@pytest.mark.parametrize("D,heads,B,Np,Nf", [(768, 4, 3, 196, 262), (768, 2, 2, 9, 30), (1152, 4, 2, 50, 120), (64, 4, 3, 16, 40)])
def test_head_project_fwd_vs_torch(kern, D, heads, B, Np, Nf):
    from bifold_b200 import _cabi as K
    gen = torch.Generator().manual_seed(D + Np)
    X = torch.randn(B * Nf, D, generator=gen).cuda()
    w = torch.randn(heads, D, generator=gen).cuda(); b = torch.randn(heads, generator=gen).cuda()
    grid = torch.empty(B, heads, Np, device="cuda")
    pa = K.HeadProjectArgs(B, heads, Np, D, X.data_ptr(), Nf * D, Nf - Np, w.data_ptr(), b.data_ptr(), grid.data_ptr(),
                           None, Nf, None, None, 0, None, None, None)
    K.check(kern.lib.bfk_head_project_fwd(C.byref(pa), torch.cuda.current_stream().cuda_stream))
    ref = torch.einsum("bnd,hd->bhn", X.view(B, Nf, D)[:, Nf - Np:], w) + b[None, :, None]
    assert (grid - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("with_bias", [False, True])
def test_cross_attention_fwd_bwd(kern, dtype, with_bias):
    """bfk_cross_attention_fwd/bwd against torch on the packed projection: queries = rows [q0, q0+nq) of each sample, keys /
    values = rows [k0, k0+nk); the additive key bias is read at row (b*H + h) % B (the reference's tiled float attn_mask,
    bifold/models/fusion.py:107-113)."""
    B, Nf, H, hd, q0, nq, k0, nk = 3, 22, 4, 16, 9, 13, 0, 9
    D = H * hd
    T = torch.float32 if dtype == "fp32" else torch.bfloat16
    g = torch.Generator(device="cpu").manual_seed(5)
    qkv = (torch.randn(B * Nf, 3 * D, generator=g) * 0.7).cuda().to(T)
    dout = torch.randn(B * Nf, D, generator=g).cuda().to(T)
    bias = (torch.rand(B, nk, generator=g) > 0.4).float().cuda() if with_bias else None
    out = torch.zeros(B * Nf, D, device="cuda", dtype=T)
    lse = torch.empty(B * H * nq, device="cuda")
    sc = hd ** -0.5
    wts = torch.empty(B, nq, nk, device="cuda")
    kern.cross_attention(qkv, out, lse, B, Nf, H, hd, q0, nq, k0, nk, sc, key_bias=bias, weights=wts)
    dqkv = torch.zeros_like(qkv)
    kern.cross_attention(qkv, out, lse, B, Nf, H, hd, q0, nq, k0, nk, sc, key_bias=bias, dout=dout, dqkv=dqkv)
    torch.cuda.synchronize()
    x = qkv.float().view(B, Nf, 3, H, hd).requires_grad_(True)
    q = x[:, q0:q0 + nq, 0].transpose(1, 2)
    k = x[:, k0:k0 + nk, 1].transpose(1, 2)
    v = x[:, k0:k0 + nk, 2].transpose(1, 2)
    dots = q @ k.transpose(-1, -2) * sc
    if bias is not None:
        rows = (torch.arange(B)[:, None] * H + torch.arange(H)[None]) % B
        dots = dots + bias[rows.cuda()][:, :, None, :]
    ref = (dots.softmax(-1) @ v).transpose(1, 2).reshape(B, nq, D)
    go = dout.float().view(B, Nf, D)[:, q0:q0 + nq]
    (gx,) = torch.autograd.grad(ref, x, go)
    tol = 2e-5 if dtype == "fp32" else 2e-2
    got = out.float().view(B, Nf, D)
    assert (got[:, q0:q0 + nq] - ref).abs().max().item() <= tol * ref.abs().max().item()
    assert (got[:, :q0] == 0).all()                                        # rows outside the query range stay untouched
    assert (wts - dots.softmax(-1).mean(1)).abs().max().item() <= (1e-5 if dtype == "fp32" else 2e-3)   # head-averaged weights
    gref = gx.reshape(B * Nf, 3 * D)
    assert (dqkv.float() - gref).abs().max().item() <= tol * gref.abs().max().item()
