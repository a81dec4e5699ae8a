"""Diagnostic for the tcgen05 GEMM: per operand-majorness error statistics against torch (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bifold_b200.engine import Kernels
from bifold_b200 import _cabi as K

k = Kernels()
torch.manual_seed(0)


def run(M, N, Kd, ta, tb, label):
    A = (torch.randn((Kd, M) if ta else (M, Kd)) * 0.5).cuda().bfloat16()
    B = (torch.randn((Kd, N) if tb else (N, Kd)) * 0.5).cuda().bfloat16()
    D = torch.full((M, N), float("nan"), device="cuda")
    try:
        k.gemm(A, B, D, M, N, Kd, A.shape[1], B.shape[1], N, trans_a=ta, trans_b=tb, backend=K.BACKEND_TCGEN05)
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[{label}] M={M} N={N} K={Kd} ta={ta} tb={tb}: EXC {e}")
        return False
    a = A.float().t() if ta else A.float()
    b = B.float() if tb else B.float().t()
    ref = a @ b
    err = (D - ref).abs()
    bad = ~(err <= 2e-2 * ref.abs().max())
    ok = not bad.any().item()
    print(f"[{label}] M={M} N={N} K={Kd} ta={int(ta)} tb={int(tb)}: max err {err.nan_to_num(1e9).max().item():.3e} "
          f"(ref max {ref.abs().max().item():.2f}) bad {bad.float().mean().item():.4f} nan {D.isnan().float().mean().item():.4f} {'OK' if ok else 'FAIL'}")
    if not ok:
        bm = bad[:128, :min(N, 128)].float().view(16, 8, -1, 8).mean((1, 3))
        print("   bad-fraction map (8x8 blocks of first 128x128):")
        for r in range(bm.shape[0]):
            print("   " + " ".join(f"{v:.1f}" for v in bm[r].tolist()))
        # is the result a permutation of K-slices? check against partial sums over first 16/32/64 k
        for kk in (16, 32, 64):
            if kk < Kd:
                part = a[:, :kk] @ b[:kk]
                print(f"   err vs partial K[:{kk}]: {(D - part).abs().nan_to_num(1e9).max().item():.3e}")
    return ok


ok = True
for (ta, tb) in ((False, False), (False, True), (True, True), (True, False)):
    for (M, N, Kd) in ((128, 128, 64), (128, 256, 64), (128, 128, 256), (256, 512, 512), (200, 784, 784), (16768, 768, 768)):
        ok &= run(M, N, Kd, ta, tb, "tc")
print("ALL OK" if ok else "SOME FAILED")
# quick timing of the main shapes
if ok:
    for (M, N, Kd, ta, tb) in ((16768, 3072, 768, False, False), (16768, 768, 3072, False, False), (16768, 2304, 768, False, False),
                               (16768, 3072, 768, False, True), (3072, 768, 16768, True, True), (768, 768, 16768, True, True)):
        A = torch.randn((Kd, M) if ta else (M, Kd), device="cuda").bfloat16()
        B = torch.randn((Kd, N) if tb else (N, Kd), device="cuda").bfloat16()
        D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for _ in range(3):
            k.gemm(A, B, D, M, N, Kd, A.shape[1], B.shape[1], N, trans_a=ta, trans_b=tb)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            k.gemm(A, B, D, M, N, Kd, A.shape[1], B.shape[1], N, trans_a=ta, trans_b=tb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"time M={M} N={N} K={Kd} ta={int(ta)} tb={int(tb)}: {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.1f} TFLOP/s")
        a = A.float().t() if ta else A.float(); b = B.float() if tb else B.float().t()
        torch.cuda.synchronize()
