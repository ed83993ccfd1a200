"""GroupNorm(+act) single-kernel path in isolation: python tools/bench_gn.py [H W C B G]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402

H, W, C, B, G = [int(a) for a in (sys.argv[1:6] + ["8", "8", "128", "1024", "1"][len(sys.argv[1:6]):])]
ACT = int(sys.argv[6]) if len(sys.argv) > 6 else 0
dev = "cuda:0"
x = ops.Act(torch.randn(B * H * W * C, device=dev).bfloat16(), B, H, W, C)
gamma = torch.nn.Parameter(torch.rand(C, device=dev) + 0.5)
beta = torch.nn.Parameter(torch.randn(C, device=dev) * 0.1)
reps = 50
for _ in range(3):
    ops.LazyNorm(x, G, 1e-5, gamma, beta, None, ACT).materialize()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(reps):
        y = ops.LazyNorm(x, G, 1e-5, gamma, beta, None, ACT).materialize()
g.replay()
torch.cuda.synchronize()
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"gn_act {H}x{W}x{C} B={B} G={G}: {us:.1f} us per call; {2 * B * H * W * C * 2 / us / 1e3:.0f} GB/s (read+write)")
