"""Whole-block FABlock2D kernel in isolation (for ncu): python tools/ncu_fablock_full.py [H W batch]
Runs the default path (pre-staged operands + producer thread, fablock_full2_kernel); LNS_FABLOCK_STAGED=0: the in-kernel-staging
kernel (fablock_full_kernel)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402

H, W, nb = [int(a) for a in (sys.argv[1:4] + ["32", "32", "1024"][len(sys.argv[1:4]):])]
dev = "cuda:0"
dt = torch.float16
u = ops.Act(torch.randn(nb * H * W * 64, device=dev).to(dt), nb, H, W, 64)
gamma, beta = torch.rand(64, device=dev) + 0.5, torch.randn(64, device=dev) * 0.1
w = torch.nn.Parameter(torch.randn(512, 64, device=dev) / 8)
w1 = torch.nn.Parameter(torch.randn(64, 512, 1, 1, device=dev) / 22)
w2 = torch.nn.Parameter(torch.randn(64, 64, 1, 1, device=dev) / 8)
kx = torch.randn(nb, 8, H, H, device=dev) / H ** 0.5
ky = torch.randn(nb, 8, W, W, device=dev) / W ** 0.5
staged = ops._state.fablock_staged
with torch.no_grad(), ops.precision("fp16"):
    if staged:
        sc, sh, _, _, st = ops.fablock_prepass(u, 1e-5, gamma, beta, staged=True)
        w_in16, w1h = ops.fablock_staged_operands(w, w1, 8, dt)
        run = lambda: ops.fablock_full_staged(st, u, w_in16, kx, ky, 8, 1e-5, w1h, w2)  # noqa: E731
    else:
        sc, sh, _, _ = ops.fablock_prepass(u, 1e-5, gamma, beta)
        run = lambda: ops.fablock_full(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2)  # noqa: E731
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"fablock_full ({'staged' if staged else 'in-kernel staging'}) {H}x{W} batch {nb}: {ms:.3f} ms; {nb / ms:.0f} samples/ms")
