"""CPU study (no GPU): which GEMM call sites of a stage carry the 16-bit operand-rounding error.

The oracle (fp64) is run with the operands of selected conv / linear / einsum call sites rounded to IEEE half (or
bf16), everything else exact; the per-site contribution to the stage's teacher-forced rel-L2 is printed.  This is test
tooling: it drives oracle/lns_oracle.py only.

  python tools/precision_study.py ns2d decode
"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lns_oracle as O  # noqa: E402
from lns_b200.configs import get_config  # noqa: E402

RT = {"f16": torch.float16, "bf16": torch.bfloat16}


class Sites:
    """Counts GEMM call sites in call order and applies a per-site rounding policy:
    policy(idx, name) -> (round_act, round_weight) with values None | 'f16' | 'bf16' | 'split' (hi+lo of f16)."""

    def __init__(self, policy):
        self.policy = policy
        self.log = []
        self.i = 0

    def rnd(self, t, how):
        if how is None:
            return t
        if how == "split":
            hi = t.to(torch.float16).to(t.dtype)
            lo = (t - hi).to(torch.float16).to(t.dtype)
            return hi + lo
        return t.to(RT[how]).to(t.dtype)

    def site(self, name, a, w):
        ra, rw = self.policy(self.i, name)
        self.log.append((self.i, name, tuple(a.shape), tuple(w.shape)))
        self.i += 1
        return self.rnd(a, ra), self.rnd(w, rw)


def install(sites):
    orig = dict(conv2d=F.conv2d, linear=F.linear, einsum=torch.einsum)

    def conv2d(x, w, b=None, **kw):
        cin = w.shape[1]
        if cin < 32:  # tiny-channel layers run on the fp32 CUDA-core path
            return orig["conv2d"](x, w, b, **kw)
        x, w = sites.site(f"conv{w.shape[-1]}x{w.shape[-1]} {w.shape[1]}->{w.shape[0]} @{x.shape[-2]}x{x.shape[-1]}", x, w)
        return orig["conv2d"](x, w, b, **kw)

    def linear(x, w, b=None):
        if w.shape[1] < 32:
            return orig["linear"](x, w, b)
        x, w = sites.site(f"linear {w.shape[1]}->{w.shape[0]} n={x.shape[-2]}", x, w)
        return orig["linear"](x, w, b)

    def einsum(eq, a, b):
        if eq == "...i,j->...ij":
            return orig["einsum"](eq, a, b)
        a, b = sites.site(f"einsum {eq} {tuple(a.shape)}", a, b)
        return orig["einsum"](eq, a, b)

    F.conv2d, F.linear, torch.einsum = conv2d, linear, einsum
    return orig


def restore(orig):
    F.conv2d, F.linear, torch.einsum = orig["conv2d"], orig["linear"], orig["einsum"]


def build(name):
    from lns_b200.latent_dynamics import LatentDynamics
    cfg = get_config(name)
    torch.manual_seed(1234)
    model = LatentDynamics(cfg).eval()
    sd = O.randomize_zero_init(model.state_dict())
    return cfg, O.to_dtype(sd, torch.float64)


def stage_fn(name, stage, cfg, sd64, B=2, seed=11):
    x, param = O.make_inputs(cfg, B, seed=seed)
    ae = O.ae_name(cfg)
    if stage == "encode":
        return lambda: O.encode(sd64, cfg, x.double(), ae)
    z = O.encode(sd64, cfg, x.double(), ae)
    cond = O.cond_embedding(sd64, cfg, param.double(), torch.float64) if param is not None else None
    if stage == "step":
        return lambda: O.propagator_step(sd64, cfg, z, cond)
    z1 = O.propagator_step(sd64, cfg, z, cond)
    return lambda: O.decode(sd64, cfg, z1, ae)


def run(fn, policy):
    s = Sites(policy)
    orig = install(s)
    try:
        with torch.no_grad():
            y = fn()
    finally:
        restore(orig)
    return y, s.log


def main():
    name, stage = sys.argv[1], sys.argv[2]
    fmt = sys.argv[3] if len(sys.argv) > 3 else "f16"
    cfg, sd64 = build(name)
    fn = stage_fn(name, stage, cfg, sd64)
    ref, log = run(fn, lambda i, n: (None, None))
    err = lambda y: O.rel_l2(y, ref).max().item()
    full, _ = run(fn, lambda i, n: (fmt, fmt))
    print(f"{name} {stage}: all sites {fmt}: {err(full):.3e}")
    wex, _ = run(fn, lambda i, n: (fmt, None))
    print(f"  weights exact, activations {fmt}: {err(wex):.3e}")
    aex, _ = run(fn, lambda i, n: (None, fmt))
    print(f"  activations exact, weights {fmt}: {err(aex):.3e}")
    tot = 0.0
    rows = []
    for (i, nm, sa, sw) in log:
        y, _ = run(fn, lambda j, n, i=i: (fmt, fmt) if j == i else (None, None))
        ya, _ = run(fn, lambda j, n, i=i: (fmt, None) if j == i else (None, None))
        e, ea = err(y), err(ya)
        tot += e * e
        rows.append((i, nm, e, ea))
    for (i, nm, e, ea) in rows:
        print(f"  site {i:3d} {nm:55s} both {e:.2e} ({100 * e * e / tot:4.1f}%)  act-only {ea:.2e}")
    print(f"  rss of sites: {tot ** 0.5:.3e}")


if __name__ == "__main__":
    main()
