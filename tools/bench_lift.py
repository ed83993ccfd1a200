"""Propagator in_proj (1x1 conv 16 -> 128 on the fp32 latent, lift1x1_kernel) in isolation: python tools/bench_lift.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
dev = "cuda:0"
x = ops.Act(torch.randn(B * 8 * 8 * 16, device=dev), B, 8, 8, 16)
wt = torch.nn.Parameter(torch.randn(128, 16, 1, 1, device=dev) / 4)
bs = torch.nn.Parameter(torch.zeros(128, device=dev))
filt = ops.PackedFilter.of(wt, bs)
out = ops.Act.empty(B, 8, 8, 128, torch.bfloat16, dev)
with ops.precision("bf16"):
    for _ in range(3):
        ops.conv2d(x, filt, out=out)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            ops.conv2d(x, filt, out=out)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
print(f"lift 16->128 @8x8 B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
