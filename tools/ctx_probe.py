"""Is the in-rollout slowdown of the up-sampling halo conv caused by its DATA or by its CONTEXT (preceding kernel)?
Runs the eager NS2d rollout; the 32x32->64x64 conv is launched three times in a row (same arguments) and each launch is
timed with its own CUDA events; afterwards the same conv is repeated on a copy of the captured input / on randn data."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402

import lns_oracle as O  # noqa: E402
from lns_b200 import ops  # noqa: E402
from lns_b200.configs import get_config  # noqa: E402
from lns_b200.latent_dynamics import LatentDynamics  # noqa: E402
from lns_b200.rollout import Rollout  # noqa: E402

cfg = get_config("ns2d")
torch.manual_seed(1234)
model = LatentDynamics(cfg).eval()
model.load_state_dict(O.randomize_zero_init(model.state_dict()))
model = model.to("cuda:0")
x, p = O.make_inputs(cfg, 1024, seed=0)
x = x.to("cuda:0")
ro = Rollout(model, batch=1024, steps=20, to_x=True, precision="bf16", use_graph=False)
orig = ops.conv2d
rec = []
keep = {}


def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    return r, (e0, e1)


def patched(xa, filt, **kw):
    if kw.get("virt") is not None and xa.H == 32 and xa.B == 4096:
        evs = []
        for _ in range(3):
            r, ev = timed(lambda: orig(xa, filt, **kw))
            evs.append(ev)
        rec.append(evs)
        keep["x"], keep["filt"], keep["kw"] = xa, filt, kw
        return r
    return orig(xa, filt, **kw)


ops.conv2d = patched
with torch.no_grad():
    ro.build()
    torch.cuda.synchronize()
    rec.clear()
    ro(x, p)
    torch.cuda.synchronize()
for evs in rec:
    print("in rollout, 3 launches in a row:", " ".join(f"{a.elapsed_time(b) * 1e3:.0f} us" for a, b in evs))
xa, filt, kw = keep["x"], keep["filt"], keep["kw"]
with torch.no_grad(), ops.precision("bf16"):
    for name, t in (("captured input", xa.t.clone()), ("randn", torch.randn_like(xa.t.float()).bfloat16()),
                    ("captured input again", xa.t.clone())):
        xb = ops.Act(t, xa.B, xa.H, xa.W, xa.C)
        ts = []
        for _ in range(4):
            _, (a, b) = timed(lambda: orig(xb, filt, **kw))
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"after the rollout, {name}:", " ".join(f"{v:.0f} us" for v in ts),
              f"| absmax {t.float().abs().max().item():.3g} mean|x| {t.float().abs().mean().item():.3g}")

# what makes the FIRST launch slow?  same conv after (i) a spin kernel, (ii) a 4 GB memset, (iii) a different halo conv
with torch.no_grad(), ops.precision("bf16"):
    xb = ops.Act(xa.t.clone(), xa.B, xa.H, xa.W, xa.C)
    big = torch.empty(1 << 30, dtype=torch.float32, device="cuda:0")
    x2 = ops.Act(torch.randn(1024 * 64 * 64 * 64, device="cuda:0").bfloat16(), 1024, 64, 64, 64)
    kw2 = dict(kw)
    kw2["virt"] = None

    def other_conv():
        orig(x2, filt, **kw2)

    for name, pre in (("spin 1 ms", lambda: torch.cuda._sleep(2_000_000)), ("4 GB memset", lambda: big.zero_()),
                      ("another halo conv (64x64, 1024 samples)", other_conv), ("itself", lambda: orig(xb, filt, **kw)),
                      ("5 x 4 GB memset", lambda: [big.zero_() for _ in range(5)]),
                      ("3 x another halo conv", lambda: [other_conv() for _ in range(3)])):
        ts = []
        for _ in range(3):
            pre()
            _, (a, b) = timed(lambda: orig(xb, filt, **kw))
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"conv after {name}:", " ".join(f"{v:.0f} us" for v in ts))

# first read of a FRESHLY WRITTEN input: which state makes it slow?
with torch.no_grad(), ops.precision("bf16"):
    def fresh():
        return ops.Act(xa.t.clone(), xa.B, xa.H, xa.W, xa.C)

    def run(xc):
        _, (a, b) = timed(lambda: orig(xc, filt, **kw))
        torch.cuda.synchronize()
        return a.elapsed_time(b) * 1e3

    for name, prep in (("clone -> conv", lambda xc: None), ("clone -> sum(x) -> conv", lambda xc: xc.t.float().sum()),
                       ("clone -> 4 GB memset -> conv", lambda xc: big.zero_()),
                       ("clone -> 4 GB memset -> sum(x) -> conv", lambda xc: (big.zero_(), xc.t.sum())),
                       ("clone -> conv (fp32 out: no TMA store)", None)):
        ts = []
        for _ in range(3):
            xc = fresh()
            torch.cuda.synchronize()
            if prep is None:
                kw3 = dict(kw)
                kw3["out_dtype"] = torch.float32
                _, (a, b) = timed(lambda: orig(xc, filt, **kw3))
                torch.cuda.synchronize()
                t1 = a.elapsed_time(b) * 1e3
                _, (a, b) = timed(lambda: orig(xc, filt, **kw3))
                torch.cuda.synchronize()
                ts.append((t1, a.elapsed_time(b) * 1e3))
            else:
                prep(xc)
                torch.cuda.synchronize()
                ts.append((run(xc), run(xc)))
        print(f"{name}:", " ".join(f"{a:.0f}/{b:.0f}" for a, b in ts), "us (first/second launch)")
