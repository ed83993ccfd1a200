"""Decoder output projection (GroupNorm apply -> Swish -> 1x1 64 -> 1, NCHW fp32 out) in isolation: python tools/bench_proj.py [B] [bf16|fp16]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
dev = "cuda:0"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
x = ops.Act(torch.randn(B * 64 * 64 * 64, device=dev).to(torch.float16 if prec == "fp16" else torch.bfloat16), B, 64, 64, 64)
wt = torch.nn.Parameter(torch.randn(1, 64, 1, 1, device=dev) / 8)
bs = torch.nn.Parameter(torch.zeros(1, device=dev))
filt = ops.PackedFilter.of(wt, bs)
sc, sh = torch.rand(B * 64, device=dev) + 0.5, torch.randn(B * 64, device=dev) * 0.1
out = ops.Act(torch.empty(B * 64 * 64, device=dev), B, 64, 64, 1, layout=ops.NCHW)
with ops.precision(prec):
    for _ in range(3):
        ops.conv2d(x, filt, pro=(sc, sh, ops.ACT_SILU), out=out, out_layout=ops.NCHW)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    ev[0].record()
    for i in range(5):
        ops.conv2d(x, filt, pro=(sc, sh, ops.ACT_SILU), out=out, out_layout=ops.NCHW)
        ev[i + 1].record()
    torch.cuda.synchronize()
print(f"proj 64->1 @64x64 B={B} {prec}:", " ".join(f"{ev[i].elapsed_time(ev[i + 1]) * 1e3:.0f}" for i in range(5)), "us")
