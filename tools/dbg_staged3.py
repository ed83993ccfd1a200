"""debug: run-to-run determinism of the whole-block FABlock kernels over many launches: python tools/dbg_staged3.py n B reps"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lns_b200 import ops
DEV = "cuda:0"
dt = torch.float16
n, nb, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(1)
u = ops.Act(torch.randn(nb * n * n * 64, device=DEV).to(dt), nb, n, n, 64)
gamma, beta = torch.rand(64, device=DEV) + 0.5, torch.randn(64, device=DEV) * 0.1
w = torch.nn.Parameter(torch.randn(512, 64, device=DEV) / 8)
w1 = torch.nn.Parameter(torch.randn(64, 512, 1, 1, device=DEV) / 22)
w2 = torch.nn.Parameter(torch.randn(64, 64, 1, 1, device=DEV) / 8)
kx = torch.randn(nb, 8, n, n, device=DEV) / n ** 0.5
ky = torch.randn(nb, 8, n, n, device=DEV) / n ** 0.5
with torch.no_grad(), ops.precision("fp16"):
    sc, sh, px, py, st = ops.fablock_prepass(u, 1e-5, gamma, beta, staged=True)
    w_in16, w1h = ops.fablock_staged_operands(w, w1, 8, dt)
    ref = ops.fablock_full(u, sc, sh, w, kx, ky, 8, 1e-5, w1, w2).t.float().view(nb, -1)
    bad = []
    for _ in range(reps):
        o = ops.fablock_full_staged(st, u, w_in16, kx, ky, 8, 1e-5, w1h, w2).t.float().view(nb, -1)
        per = (o - ref).norm(dim=1) / ref.norm(dim=1)
        bad.append(int((per > 2e-3).sum()))
    torch.cuda.synchronize()
print(f"{n}x{n} B={nb} dbg={os.environ.get('LNS_DBG_FULL2', '0')}: wrong samples per launch over {reps} launches: total {sum(bad)} {bad}", flush=True)
