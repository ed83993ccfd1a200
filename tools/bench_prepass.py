"""FABlock2D pre-pass in isolation: python tools/bench_prepass.py [H W B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402

H, W, B = [int(a) for a in (sys.argv[1:4] + ["32", "32", "4096"][len(sys.argv[1:4]):])]
dev = "cuda:0"
x = ops.Act(torch.randn(B * H * W * 64, device=dev).bfloat16(), B, H, W, 64)
gamma = torch.nn.Parameter(torch.rand(64, device=dev) + 0.5)
beta = torch.nn.Parameter(torch.randn(64, device=dev) * 0.1)
reps = 20
for _ in range(3):
    ops.fablock_prepass(x, 1e-5, gamma, beta)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(reps):
        r = ops.fablock_prepass(x, 1e-5, gamma, beta)
g.replay()
torch.cuda.synchronize()
e0.record()
g.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
print(f"fablock_prepass {H}x{W}x64 B={B}: {us:.1f} us per call; {B * H * W * 64 * 2 / us / 1e3:.0f} GB/s (read)")
