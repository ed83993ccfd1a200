"""Training-rollout benchmark: one optimisation step = LatentDynamics.forward(z_in, z_out,
smooth_l1) + backward + AdamW.step at the reference's training shape (configs/ns2d_stage2_prop.yml: batch 32, out_tw 2) and at a
throughput batch, next to the same step in PyTorch eager on the same GPU (autograd of the oracle's restatement, fp32)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "oracle")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)


def _timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def run(cases=((32, 2), (1024, 2)), iters=10, eager=True):
    import lns_oracle as O
    from lns_b200 import ops
    from lns_b200.configs import get_config
    from lns_b200.latent_dynamics import LatentDynamics
    cfg = get_config("ns2d")
    out = {"workload": "NS2d latent training step: forward(z_in, z_out, smooth_l1) + backward + AdamW (configs/ns2d_stage2_prop.yml)",
           "unit": "trajectory-steps/s", "cases": []}
    for B, T in cases:
        torch.manual_seed(1234)
        model = LatentDynamics(cfg)
        model.load_state_dict(O.randomize_zero_init(model.state_dict()))
        model = model.cuda()
        for p in model.autoencoder.parameters():
            p.requires_grad_(False)
        opt = torch.optim.AdamW(model.propagator.parameters(), lr=1e-5)
        z_in, z_out = O.train_inputs(cfg, B, T, seed=0)
        z_in, z_out = z_in.cuda(), z_out.cuda()
        row = {"batch": B, "t_out": T}
        for prec in ("fp16s", "fp32"):
            def step():
                opt.zero_grad(set_to_none=True)
                with ops.precision(prec):
                    loss = model(z_in, z_out, F.smooth_l1_loss)
                    loss.backward()
                opt.step()
            ms = _timed(step, iters)
            row[prec] = {"ms_per_iter": round(ms, 3), "value": round(B * T / ms * 1e3, 1)}
        # the same step (forward + backward + AdamW) captured once as a CUDA graph and replayed (lns_b200.train.GraphedTrainStep)
        try:
            from lns_b200.train import GraphedTrainStep
            optg = torch.optim.AdamW(model.propagator.parameters(), lr=1e-5, capturable=True)
            gs = GraphedTrainStep(model, z_in, z_out, F.smooth_l1_loss, optimizer=optg, precision="fp16s")
            ms = _timed(lambda: gs(z_in, z_out), iters)
            row["fp16s_cuda_graph"] = {"ms_per_iter": round(ms, 3), "value": round(B * T / ms * 1e3, 1)}
            del gs
        except Exception as ex:  # noqa: BLE001
            row["fp16s_cuda_graph"] = {"error": repr(ex)[:200]}
        if eager:
            sd = {k: v.detach().clone().requires_grad_(k.startswith("propagator.")) for k, v in model.state_dict().items()}
            params = [v for k, v in sd.items() if k.startswith("propagator.")]
            opt2 = torch.optim.AdamW(params, lr=1e-5)

            def estep():
                opt2.zero_grad(set_to_none=True)
                F.smooth_l1_loss(O.train_rollout(sd, cfg, z_in[:, 0], T), z_out).backward()
                opt2.step()
            ms = _timed(estep, iters)
            row["torch_eager_fp32_same_gpu"] = {"ms_per_iter": round(ms, 3), "value": round(B * T / ms * 1e3, 1)}
        out["cases"].append(row)
    # the other three stage-2 scripts at their own training shapes (configs/*_stage2*_prop.yml: batch 32, out_tw 5), graphed step
    out["other_configs"] = {}
    for name in ("sw", "twophase", "twophase_cond"):
        try:
            from lns_b200.train import GraphedTrainStep
            c2 = get_config(name)
            torch.manual_seed(1234)
            m2 = LatentDynamics(c2)
            m2.load_state_dict(O.randomize_zero_init(m2.state_dict()))
            m2 = m2.cuda()
            for p in m2.autoencoder.parameters():
                p.requires_grad_(False)
            zi, zo = O.train_inputs(c2, 32, 5, seed=0)
            zi, zo = zi.cuda(), zo.cuda()
            par = torch.linspace(0.3, 0.9, 32).cuda() if name == "twophase_cond" else None
            optg = torch.optim.AdamW(m2.propagator.parameters(), lr=1e-5, capturable=True)
            gs = GraphedTrainStep(m2, zi, zo, F.smooth_l1_loss, param=par, optimizer=optg, precision="fp16s")
            ms = _timed((lambda: gs(zi, zo, par)) if par is not None else (lambda: gs(zi, zo)), iters)
            out["other_configs"][name] = {"batch": 32, "t_out": 5, "fp16s_cuda_graph_ms_per_iter": round(ms, 3),
                                          "value": round(32 * 5 / ms * 1e3, 1)}
            del gs, m2
            torch.cuda.empty_cache()
        except Exception as ex:  # noqa: BLE001
            out["other_configs"][name] = {"error": repr(ex)[:200]}
    return out


if __name__ == "__main__":
    print(json.dumps(run()))
