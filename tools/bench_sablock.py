"""SABlock whole-block kernel in isolation: python tools/bench_sablock.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lns_b200 import ops  # noqa: E402
from modules.basics import SABlock  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dev = "cuda:0"
torch.manual_seed(0)
blk = SABlock(128, 8, 64, use_pe=True, block_size=64).to(dev).eval()
x = ops.Act(torch.randn(B * 64 * 128, device=dev).bfloat16(), B, 8, 8, 128)
with torch.no_grad(), ops.precision("bf16"):
    for _ in range(2):
        blk._fwd(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        blk._fwd(x)
    e1.record()
    torch.cuda.synchronize()
print(f"sablock_fused 8x8x128 B={B}: {e0.elapsed_time(e1) / 3 * 1e3:.1f} us")
