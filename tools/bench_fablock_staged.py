"""FABlock2D whole-block kernels timed alone (CUDA events): in-kernel staging (fablock_full) vs pre-staged operands + producer thread
(fablock_full_staged), and the pre-pass with / without the staged copy.
    python tools/bench_fablock_staged.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lns_b200 import ops  # noqa: E402

DEV = "cuda:0"
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4736


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for n in (32, 16):
    for prec, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        u = ops.Act(torch.randn(nb * n * n * 64, device=DEV).to(dt), nb, n, n, 64)
        gamma, beta = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV) * 0.1
        w = torch.nn.Parameter(torch.randn(512, 64, device=DEV) / 8)
        w1 = torch.nn.Parameter(torch.randn(64, 512, 1, 1, device=DEV) / 22)
        w2 = torch.nn.Parameter(torch.randn(64, 64, 1, 1, device=DEV) / 8)
        kx = torch.randn(nb, 8, n, n, device=DEV) / n ** 0.5
        ky = torch.randn(nb, 8, n, n, device=DEV) / n ** 0.5
        fl = nb * (2.0 * n * n * 64 * 512 * 2 + 2.0 * 8 * (2 * n * n * n) * 64 + 2.0 * n * n * 64 * 64)
        with torch.no_grad(), ops.precision(prec):
            sc, sh, _, _, st = ops.fablock_prepass(u, 1e-5, gamma, beta, staged=True)
            w_in16, w1h = ops.fablock_staged_operands(w, w1, 8, dt)
            o_full = ops.fablock_full(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2)
            o_st = ops.fablock_full_staged(st, u, w_in16, kx, ky, 8, 1e-5, w1h, w2)
            torch.cuda.synchronize()
            d = (o_st.t.float() - o_full.t.float()).norm() / o_full.t.float().norm()
            t_full = timed(lambda: ops.fablock_full(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2))
            t_st = timed(lambda: ops.fablock_full_staged(st, u, w_in16, kx, ky, 8, 1e-5, w1h, w2))
            t_p0 = timed(lambda: ops.fablock_prepass(u, 1e-5, gamma, beta))
            t_p1 = timed(lambda: ops.fablock_prepass(u, 1e-5, gamma, beta, staged=True))
        print(f"FABlock2D {n}x{n} x{nb} {prec}: in-kernel staging {t_full:.3f} ms ({fl / t_full / 1e9:.0f} TFLOP/s) | staged + producer thread "
              f"{t_st:.3f} ms ({fl / t_st / 1e9:.0f} TFLOP/s) | rel diff {d:.2e} | pre-pass {t_p0:.3f} -> {t_p1:.3f} ms", flush=True)
