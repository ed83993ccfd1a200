"""One eager (no CUDA graph) rollout for kernel-level profiling:  ncu ... python tools/ncu_rollout.py [workload] [B] [R]
Prints per-call launch count; with LNS_TIMELINE=1 also prints a CUDA-event timeline per op family (not under ncu)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import lns_oracle as O  # noqa: E402
from lns_b200.configs import get_config  # noqa: E402
from lns_b200.latent_dynamics import LatentDynamics  # noqa: E402
from lns_b200.rollout import Rollout  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "ns2d"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
R = int(sys.argv[3]) if len(sys.argv) > 3 else 20
prec = sys.argv[4] if len(sys.argv) > 4 else "bf16"
cfg = get_config(name)
torch.manual_seed(1234)
model = LatentDynamics(cfg).eval()
model.load_state_dict(O.randomize_zero_init(model.state_dict()))
model = model.to("cuda:0")
x, p = O.make_inputs(cfg, B, seed=0)
x = x.to("cuda:0")
p = p.to("cuda:0") if p is not None else None
ro = Rollout(model, batch=B, steps=R, to_x=True, precision=prec, use_graph=False)
with torch.no_grad():
    ro.build()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro(x, p)
    e1.record()
    torch.cuda.synchronize()
print(f"{name} B={B} R={R} {prec}: {ro.launches_per_call} launches, eager {e0.elapsed_time(e1):.2f} ms "
      f"-> {B * R / e0.elapsed_time(e1) * 1e3:.0f} trajectory-steps/s")
