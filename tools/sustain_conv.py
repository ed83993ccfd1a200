"""Burst vs sustained speed of one halo conv: python tools/sustain_conv.py [up|same] [iters] [data: randn|small|zeros]
Prints ms per launch for consecutive groups of 10 launches, with SM clock / power sampled through NVML in between."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import pynvml  # noqa: E402

from lns_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "up"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
data = sys.argv[3] if len(sys.argv) > 3 else "randn"
rot = int(sys.argv[4]) if len(sys.argv) > 4 else 1  # rotate over this many (input, output) buffer pairs
H = W = 32 if mode == "up" else 64
nb = 4096 if mode == "up" else 1024
dev = "cuda:0"
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
xt = torch.randn(nb * H * W * 64, device=dev)
if data == "small":
    xt = torch.nn.functional.silu(xt) * 0.3
elif data == "zeros":
    xt.zero_()
xs = [ops.Act(xt.bfloat16().clone(), nb, H, W, 64) for _ in range(rot)]
wt = torch.nn.Parameter(torch.randn(64, 64, 3, 3, device=dev) / math.sqrt(576))
bs = torch.nn.Parameter(torch.zeros(64, device=dev))
filt = ops.PackedFilter.of(wt, bs)
Ho, Wo = (2 * H, 2 * W) if mode == "up" else (H, W)
virt = (Ho, Wo) if mode == "up" else None
outs = [ops.Act.empty(nb, Ho, Wo, 64, torch.bfloat16, dev) for _ in range(rot)]
k = 0
flop = 2.0 * nb * Ho * Wo * 64 * 576
with ops.precision("bf16"):
    for _ in range(3):
        ops.conv2d(xs[0], filt, dil=1, pad=(1,) * 4, pad_mode=(1, 1), virt=virt, out=outs[0], engine=ops.ENGINE_HALO)
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters // 10 + 1)]
    smp = []
    evs[0].record()
    for g in range(iters // 10):
        for _ in range(10):
            ops.conv2d(xs[k % rot], filt, dil=1, pad=(1,) * 4, pad_mode=(1, 1), virt=virt, out=outs[k % rot], engine=ops.ENGINE_HALO)
            k += 1
        evs[g + 1].record()
        evs[g + 1].synchronize()
        smp.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
    torch.cuda.synchronize()
for g in range(iters // 10):
    ms = evs[g].elapsed_time(evs[g + 1]) / 10
    print(f"{mode} {data} rot{rot} group {g:3d}: {ms:.4f} ms/launch  {flop / ms / 1e9:7.1f} TFLOP/s  sm {smp[g][0]} MHz  {smp[g][1]:.0f} W")
