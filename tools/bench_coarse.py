"""Engine comparison on the coarse-level conv shapes (CUDA events, back-to-back launches):
    python tools/bench_coarse.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lns_b200 import ops  # noqa: E402

DEV = "cuda:0"
SHAPES = [
    # name, B, Cin, Cout, H, W, dil, modes, virt
    ("ns2d 8x8 128->128", 4736, 128, 128, 8, 8, 1, (1, 1), None),
    ("ns2d 8x8 128->128 d2 (propagator, B=1184)", 1184, 128, 128, 8, 8, 2, (1, 1), None),
    ("ns2d up 8->16 128->128", 4736, 128, 128, 8, 8, 1, (1, 1), (16, 16)),
    ("ns2d 16x16 128->64", 4736, 128, 64, 16, 16, 1, (1, 1), None),
    ("ns2d 16x16 64->64", 4736, 64, 64, 16, 16, 1, (1, 1), None),
    ("sw 12x24 128->128 d3 (propagator, B=64)", 64, 128, 128, 12, 24, 3, (0, 1), None),
    ("sw 12x24 128->128 (decoder, B=320)", 320, 128, 128, 12, 24, 1, (0, 1), None),
    ("tp 7x15 128->128 d2 (propagator, B=128)", 128, 128, 128, 7, 15, 2, (0, 0), None),
    ("tp 7x15 128->128 (decoder, B=640)", 640, 128, 128, 7, 15, 1, (0, 0), None),
]


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    only = os.environ.get("BENCH_COARSE_ONLY")       # substring of the shape name
    only_cand = os.environ.get("BENCH_COARSE_CAND")  # substring of the candidate label
    for name, B, Cin, Cout, H, W, dil, modes, virt in SHAPES:
        if only and only != name:
            continue
        w = torch.nn.Parameter(torch.randn(Cout, Cin, 3, 3, device=DEV) / math.sqrt(9 * Cin))
        b = torch.nn.Parameter(torch.zeros(Cout, device=DEV))
        filt = ops.PackedFilter.of(w, b)
        Ho, Wo = virt if virt else (H, W)
        fl = 2.0 * B * Ho * Wo * Cout * 9 * Cin
        x16 = ops.Act(torch.randn(B * H * W * Cin, device=DEV).half(), B, H, W, Cin)
        x32 = ops.Act(torch.randn(B * H * W * Cin, device=DEV), B, H, W, Cin)
        res = []
        kw = dict(dil=dil, pad=(dil,) * 4, pad_mode=modes, virt=virt, act=ops.ACT_GELU)
        with torch.no_grad(), ops.precision("fp16"):
            cands = [("umma f16", x16, ops.ENGINE_UMMA, False, torch.float16),
                     ("coarse f16", x16, ops.ENGINE_COARSE, False, torch.float16),
                     ("coarse f32 x2a", x32, ops.ENGINE_COARSE, False, torch.float32),
                     ("coarse f32 x3", x32, ops.ENGINE_COARSE, True, torch.float32),
                     ("umma f32 x3", x32, ops.ENGINE_UMMA, True, torch.float32)]
            if (H, W) == (8, 8) and virt is None and modes == (1, 1) and Cin == 128 and Cout == 128:
                cands.insert(0, ("latent f16", x16, ops.ENGINE_LATENT, False, torch.float16))
            for label, x, eng, split, odt in cands:
                if only_cand and only_cand not in label:
                    continue
                if eng == ops.ENGINE_COARSE and not ops._coarse_fits(Cin, Cout, dil, x.t.dtype == torch.float32, split):
                    continue
                out = ops.Act.empty(B, Ho, Wo, Cout, odt, DEV)
                try:
                    t = timed(lambda: ops.conv2d(x, filt, engine=eng, split=split, out=out, **kw))
                    res.append(f"{label} {t:7.1f} us ({fl / t / 1e6:6.0f} TF/s)")
                except Exception as ex:  # noqa: BLE001
                    res.append(f"{label} FAILED {str(ex)[:60]}")
        print(f"{name:46s} " + " | ".join(res), flush=True)


if __name__ == "__main__":
    main()
