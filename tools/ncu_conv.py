"""One tcgen05 conv layer in isolation (for ncu --set full):  python tools/ncu_conv.py [H W Cin Cout batch dil]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402

H, W, Cin, Cout, nb, dil = [int(a) for a in (sys.argv[1:7] + ["64", "64", "64", "64", "256", "1"][len(sys.argv[1:7]):])]
ENGINE = {"umma": ops.ENGINE_UMMA, "halo": ops.ENGINE_HALO, "latent": ops.ENGINE_LATENT}[sys.argv[7] if len(sys.argv) > 7 else "umma"]
ACT = int(sys.argv[8]) if len(sys.argv) > 8 else 0
UP = len(sys.argv) > 9 and "up" in sys.argv[9]
RES = len(sys.argv) > 9 and "res" in sys.argv[9]  # + residual (16-bit, output shape)  # nearest x2 folded into the conv: output 2H x 2W
dev = "cuda:0"
x = ops.Act(torch.randn(nb * H * W * Cin, device=dev).bfloat16(), nb, H, W, Cin)
wt = torch.nn.Parameter(torch.randn(Cout, Cin, 3, 3, device=dev) / math.sqrt(9 * Cin))
bs = torch.nn.Parameter(torch.zeros(Cout, device=dev))
filt = ops.PackedFilter.of(wt, bs)
Ho, Wo = (2 * H, 2 * W) if UP else (H, W)
VIRT = (Ho, Wo) if UP else None
out = ops.Act.empty(nb, Ho, Wo, Cout, torch.bfloat16, dev)
res = ops.Act(torch.randn(nb * Ho * Wo * Cout, device=dev).bfloat16(), nb, Ho, Wo, Cout) if RES else None
with ops.precision("bf16"):
    for _ in range(3):
        ops.conv2d(x, filt, dil=dil, pad=(dil,) * 4, pad_mode=(1, 1), virt=VIRT, out=out, engine=ENGINE, act=ACT, residual=res)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv2d(x, filt, dil=dil, pad=(dil,) * 4, pad_mode=(1, 1), virt=VIRT, out=out, engine=ENGINE, act=ACT, residual=res)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"[{sys.argv[7] if len(sys.argv) > 7 else chr(117)+chr(109)+chr(109)+chr(97)}] conv3x3 {Cin}->{Cout} @ {H}x{W} batch {nb} dil {dil}: {ms:.4f} ms, "
      f"{2.0 * nb * Ho * Wo * Cout * 9 * Cin / ms / 1e9:.1f} TFLOP/s" + (" (x2 up)" if UP else "") + (" +res" if RES else ""))
