"""FABlock2D pooled-branch kernel in isolation: python tools/bench_fa_axis.py [n B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402
from modules.factorized_attention import FABlock2D  # noqa: E402

n, B = [int(a) for a in (sys.argv[1:3] + ["32", "4096"][len(sys.argv[1:3]):])]
dev = "cuda:0"
torch.manual_seed(0)
blk = FABlock2D(64, 64, 64, 8, 64).to(dev).eval()
mx = ops.Act(torch.randn(B * n * 64, device=dev), B, n, 1, 64)
with torch.no_grad(), ops.precision("bf16"):
    for _ in range(2):
        blk._axis_kernels(mx, mx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        blk._axis_kernels(mx, mx)
    e1.record()
    torch.cuda.synchronize()
print(f"fa_axis n={n} B={B}: {e0.elapsed_time(e1) / 6 * 1e3:.1f} us per axis call")
