"""Generate tests/golden/train_grads.pt: loss and parameter gradients of the UNMODIFIED reference's training rollout
``LatentDynamics.forward(z_in, z_out, F.smooth_l1_loss)`` (train_stage2_ns2d.py:126-141 and the SW / two-phase scripts),
fp64 copy of the seed-1234 model, seeded latents.  Run in the build container only:

    python oracle/make_golden_train.py

Full gradients of the 3x3 filters would be tens of MB, so per parameter the fixture stores the gradient's L2 norm and its inner
product with a fixed pseudo-random direction (lns_oracle.grad_probe); small parameters (<= 4096 elements) are stored in full."""
import copy
import os
import sys

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402
import lns_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CASES = {"ns2d": (3, 3), "sw": (2, 2), "twophase": (2, 2), "twophase_cond": (2, 2)}  # (batch, t_out)


def main():
    torch.set_num_threads(os.cpu_count())
    fix = {}
    for name, (B, T) in CASES.items():
        model, _ = ref_shims.build_reference(name, seed=1234)
        sd = O.randomize_zero_init(model.state_dict())
        model.load_state_dict(sd, strict=True)
        from lns_b200_cfg import get_config
        cfg = get_config(name)
        z_in, z_out = O.train_inputs(cfg, B, T, seed=0)
        m64 = copy.deepcopy(model).double()
        for p in m64.parameters():
            p.requires_grad_(p is not None)
        m64.zero_grad()
        if name == "twophase_cond":
            # the reference's fourier_embedding always returns fp32 (modules/cond_utils.py:34): cast it up for the fp64 copy
            script = ref_shims.load_script(name)
            orig = script.fourier_embedding
            script.fourier_embedding = lambda t, dim: orig(t, dim).double()
            param = torch.linspace(0.3, 0.9, B)
            loss = m64(z_in.double(), z_out.double(), param.double(), F.smooth_l1_loss)
            script.fourier_embedding = orig
        else:
            loss = m64(z_in.double(), z_out.double(), F.smooth_l1_loss)
        loss.backward()
        entry = {"batch": B, "t_out": T, "loss": float(loss), "norm": {}, "probe": {}, "full": {}}
        for k, p in m64.propagator.named_parameters():
            g = p.grad.detach()
            entry["norm"][k] = float(g.norm())
            entry["probe"][k] = float((g * O.grad_probe(k, g.shape)).sum())
            if g.numel() <= 4096:
                entry["full"][k] = g.clone()
        fix[name] = entry
        print(name, "loss %.6f" % entry["loss"], len(entry["norm"]), "parameters")
    torch.save(fix, os.path.join(OUT, "train_grads.pt"))


if __name__ == "__main__":
    import importlib.util
    p = os.path.join(os.path.dirname(HERE), "lns-latent-neural-pde-solver_b200", "configs.py")
    spec = importlib.util.spec_from_file_location("lns_b200_cfg", p)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["lns_b200_cfg"] = mod
    spec.loader.exec_module(mod)
    main()
