"""Correctness (vs fp32 torch) of the attention kernels on multi-item persistent shapes, then timing (run on the GPU box).
ATTN_STAGE=fwd|bwd|all selects what runs (a hung kernel must not take the other stage's evidence with it)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
k = Kernels()
stage = os.environ.get("ATTN_STAGE", "all")


def ref_attn(qkv, B, N, H, hd, mask, scale):
    q, kk, v = qkv.float().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)
    dots = q @ kk.transpose(-1, -2) * scale
    if mask is not None:
        dots = dots.masked_fill(mask[:, None, :, None] == 0, -100000)
    return (torch.softmax(dots, -1) @ v).transpose(1, 2).reshape(B * N, H * hd)


cases = [(2, 70, 3, 16, True), (3, 197, 2, 64, False), (2, 262, 4, 48, True), (2, 64, 12, 64, False),
         (20, 196, 12, 64, False), (13, 262, 16, 48, True), (40, 64, 12, 64, False), (40, 196, 12, 64, True), (30, 262, 16, 48, False),
         (3, 130, 2, 32, True), (2, 256, 2, 64, False), (2, 128, 3, 48, True)]
ok = True
for (B, N, H, hd, masked) in cases:
    g = torch.Generator().manual_seed(B * 1000 + N)
    qkv = torch.randn(B * N, 3 * H * hd, generator=g).cuda().bfloat16()
    mask = None
    if masked:
        mask = torch.ones(B, N)
        mask[0, 5:5 + N // 3] = 0
        mask[B - 1, N - 7:] = 0
        mask = mask.cuda()
    out = torch.full((B * N, H * hd), float("nan"), device="cuda", dtype=torch.bfloat16)
    lse = torch.full((B * H * N,), float("nan"), device="cuda")
    scale = hd ** -0.5
    q32 = qkv.float().requires_grad_(True)
    ref = ref_attn(q32, B, N, H, hd, mask, scale)
    msg = f"B{B} N{N} H{H} hd{hd} mask{int(masked)}:"
    if stage in ("fwd", "all"):
        k.attention_fwd(qkv, out, lse, B, N, H, hd, scale, mask)
        torch.cuda.synchronize()
        err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
        dots = (q32.detach().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)[0] @ q32.detach().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)[1].transpose(-1, -2)) * scale
        if mask is not None:
            dots = dots.masked_fill(mask[:, None, :, None] == 0, -100000)
        lse_ref = torch.logsumexp(dots, -1).reshape(-1)
        lm = torch.ones(B, H, N, dtype=torch.bool, device="cuda") if mask is None else (mask[:, None, :] != 0).expand(B, H, N)
        lerr = (lse - lse_ref)[lm.reshape(-1)].abs().max().item()
        good = err == err and err <= 2e-2 and lerr <= 2e-2
        ok &= good
        msg += f" fwd rel err {err:.2e} lse err {lerr:.2e} {'ok' if good else 'FAIL'};"
    else:
        out = ref.detach().bfloat16()
        dots = (q32.detach().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)[0] @ q32.detach().view(B, N, 3, H, hd).permute(2, 0, 3, 1, 4)[1].transpose(-1, -2)) * scale
        lse = torch.logsumexp(dots, -1).reshape(-1).contiguous()
    if stage in ("bwd", "all"):
        dout = torch.randn(B * N, H * hd, generator=g).cuda().bfloat16()
        (gref,) = torch.autograd.grad(ref, q32, dout.float())
        dqkv = torch.full_like(qkv, float("nan"))
        k.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, scale, mask)
        torch.cuda.synchronize()
        D = H * hd
        errs = [((dqkv.float() - gref)[:, s * D:(s + 1) * D].abs().max() / gref[:, s * D:(s + 1) * D].abs().max()).item() for s in range(3)]
        good = all(e == e and e <= 4e-2 for e in errs)
        ok &= good
        msg += f" bwd rel err dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e} {'ok' if good else 'FAIL'}"
    print(msg, flush=True)
print("ATTN CHECK", "OK" if ok else "FAILED", flush=True)

flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
def timeit(name, fn, reps=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"{name:36s} {1e3*ts[len(ts)//2]:8.1f} us", flush=True)
for name, B, N, H, hd in [("vision", 64, 196, 12, 64), ("text", 64, 64, 12, 64), ("fusion", 64, 262, 16, 48), ("vision_ctx", 256, 196, 12, 64)]:
    D = H * hd
    qkv = (torch.randn(B * N, 3 * D, device="cuda") * 0.5).bfloat16()
    out = torch.empty(B * N, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B * H * N, device="cuda")
    dout = torch.randn(B * N, D, device="cuda").bfloat16()
    dqkv = torch.empty_like(qkv)
    mask = torch.ones(B, N, device="cuda") if name == "fusion" else None
    sc = hd ** -0.5
    if stage in ("fwd", "all"):
        timeit(f"{name} fwd  B{B} N{N} H{H} hd{hd}", lambda: k.attention_fwd(qkv, out, lse, B, N, H, hd, sc, row_mask=mask))
    else:
        k.backend = 0
    if stage in ("bwd", "all"):
        if stage == "bwd":
            os.environ["X"] = "1"
        timeit(f"{name} bwd", lambda: k.attention_bwd(qkv, out, lse, dout, dqkv, B, N, H, hd, sc, row_mask=mask))
