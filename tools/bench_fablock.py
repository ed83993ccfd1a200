"""FABlock2D whole-block kernels timed alone (CUDA events): mma.sync version (fablock_full) vs tcgen05 version (fablock_tc)
    python tools/bench_fablock.py [batch]"""
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
        sc, sh = torch.rand(nb * 64, device=DEV) + 0.5, torch.randn(nb * 64, device=DEV) * 0.1
        w = torch.nn.Parameter(torch.randn(512, 64, device=DEV) / 8)
        w1 = torch.nn.Parameter(torch.randn(64, 512, 1, 1, device=DEV) / 22)
        w2 = torch.nn.Parameter(torch.randn(64, 64, 1, 1, device=DEV) / 8)
        kx = torch.randn(nb, 8, n, n, device=DEV) / n ** 0.5
        ky = torch.randn(nb, 8, n, n, device=DEV) / n ** 0.5
        fl = nb * (2.0 * n * n * 64 * 512 * 2 + 2.0 * 8 * (2 * n * n * n) * 64 + 2.0 * n * n * 64 * 64)
        with torch.no_grad(), ops.precision(prec):
            t_full = timed(lambda: ops.fablock_full(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2))
            t_tc = timed(lambda: ops.fablock_tc(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2))
        print(f"FABlock2D {n}x{n} x{nb} {prec}: mma.sync kernel {t_full:.3f} ms ({fl / t_full / 1e9:.0f} TFLOP/s) | tcgen05 kernel "
              f"{t_tc:.3f} ms ({fl / t_tc / 1e9:.0f} TFLOP/s)", flush=True)
