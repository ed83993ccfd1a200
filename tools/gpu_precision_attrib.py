"""GPU experiment: teacher-forced per-stage error of the 16-bit modes and of variants of the split-operand mode 'fp16s'
(which layers are split, exact CUDA-core layers instead of split tensor-core layers) -- locates the 16-bit error.
  python tools/gpu_precision_attrib.py [configs] [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import lns_oracle as O  # noqa: E402
from lns_b200 import ops  # noqa: E402
from lns_b200.configs import get_config  # noqa: E402
from lns_b200.latent_dynamics import LatentDynamics  # noqa: E402

DEV = "cuda:0"
VARIANTS = [
    ("fp16", "fp16", {}),
    ("fp16s", "fp16s", {}),
    ("fp16s hi: +filter split", "fp16s", dict(hi_wsplit=True)),
    ("fp16s wsplit all", "fp16s", dict(wsplit_policy="all")),
    ("fp16s wsplit all + hi w", "fp16s", dict(wsplit_policy="all", hi_wsplit=True)),
    ("fp16s wsplit none", "fp16s", dict(wsplit_policy="none")),
    ("fp16s gather x3 (no coarse)", "fp16s", dict(coarse=False)),
    ("fp16s hi x4 (3 levels)", "fp16s", dict(hi_scale=4.0)),
    ("fp16s hi 1 level", "fp16s", dict(hi_scale=0.25)),
]


def main():
    names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["ns2d", "sw", "twophase", "twophase_cond"]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    for name in names:
        cfg = get_config(name)
        torch.manual_seed(1234)
        model = LatentDynamics(cfg).eval()
        sd = O.randomize_zero_init(model.state_dict())
        model.load_state_dict(sd, strict=True)
        model = model.to(DEV)
        sd64 = O.to_dtype(sd, torch.float64)
        x, param = O.make_inputs(cfg, B, seed=11)
        ae = O.ae_name(cfg)
        z_ref = O.encode(sd64, cfg, x.double(), ae)
        cond = O.cond_embedding(sd64, cfg, param.double(), torch.float64) if param is not None else None
        z1_ref = O.propagator_step(sd64, cfg, z_ref, cond)
        y_ref = O.decode(sd64, cfg, z1_ref, ae)
        for label, prec, kw in VARIANTS:
            saved = {k: getattr(ops._state, k) for k in kw}
            for k, v in kw.items():
                setattr(ops._state, k, v)
            try:
                with torch.no_grad(), ops.precision(prec):
                    z = model.autoencoder.encode(x.to(DEV))
                    y = model.autoencoder.decode(z1_ref.float().to(DEV))
                    z1 = model.propagator(z_ref.float().to(DEV)) if param is None else \
                        model.propagator(z_ref.float().to(DEV), param.to(DEV))
                torch.cuda.synchronize()
                e = [O.rel_l2(a.cpu(), b) for a, b in ((z, z_ref), (z1, z1_ref), (y, y_ref))]
                print(f"[{name}] {label:26s} encode {e[0].max():.2e} (mean {e[0].mean():.2e})  "
                      f"step {e[1].max():.2e} ({e[1].mean():.2e})  decode {e[2].max():.2e} ({e[2].mean():.2e})", flush=True)
            except Exception as ex:  # noqa: BLE001
                print(f"[{name}] {label:26s} FAILED: {ex}", flush=True)
            finally:
                for k, v in saved.items():
                    setattr(ops._state, k, v)


if __name__ == "__main__":
    main()
