"""Per-op timeline of one eager rollout (CUDA events around every library call):
    python tools/timeline.py [workload] [B] [R]    -> table of op label, calls, total ms, share"""
import collections
import os
import sys

os.environ.setdefault("LNS_ROLLOUT_PIPELINE", "0")  # serial order: with the decode pipelined on a second stream the per-call events overlap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import lns_oracle as O  # noqa: E402
from lns_b200 import ops  # noqa: E402
from lns_b200.configs import get_config  # noqa: E402
from lns_b200.latent_dynamics import LatentDynamics  # noqa: E402
from lns_b200.rollout import Rollout  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ns2d"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
R = int(sys.argv[3]) if len(sys.argv) > 3 else 20
cfg = get_config(name)
torch.manual_seed(1234)
model = LatentDynamics(cfg).eval()
model.load_state_dict(O.randomize_zero_init(model.state_dict()))
model = model.to("cuda:0")
x, p = O.make_inputs(cfg, B, seed=0)
x = x.to("cuda:0")
p = p.to("cuda:0") if p is not None else None
ro = Rollout(model, batch=B, steps=R, to_x=True, precision=os.environ.get("LNS_TL_PREC", "bf16"), use_graph=False)
with torch.no_grad():
    ro.build()
    torch.cuda.synchronize()
    ops._state.timeline = []
    ro(x, p)
    torch.cuda.synchronize()
tl, ops._state.timeline = ops._state.timeline, None
agg = collections.defaultdict(lambda: [0, 0.0])
for label, e0, e1, _fl, _by in tl:
    agg[label][0] += 1
    agg[label][1] += e0.elapsed_time(e1)
tot = sum(v[1] for v in agg.values())
print(f"{name} B={B} R={R}: {len(tl)} timed calls, {tot:.2f} ms (event-bracketed, eager)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t:9.3f} ms {100 * t / tot:5.1f}%  n={n:4d} avg={1e3 * t / n:8.1f} us  {k}")
