"""CPU tests: the oracle restatement (oracle/lns_oracle.py) against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare the CUDA path with it."""
import hashlib
import os

import pytest
import torch

import lns_oracle as O
from lns_b200.configs import get_config
from lns_b200.latent_dynamics import LatentDynamics

CONFIGS = ["ns2d", "sw", "twophase", "twophase_cond"]


def _sha(t):
    return hashlib.sha1(t.detach().cpu().contiguous().numpy().tobytes()).hexdigest()


def build_state(name):
    """Weights exactly as make_golden.py used them: seed 1234 through the drop-in constructors + redrawn zero-inits."""
    cfg = get_config(name)
    torch.manual_seed(1234)
    model = LatentDynamics(cfg).eval()
    return cfg, model, O.randomize_zero_init(model.state_dict())


@pytest.mark.parametrize("name", CONFIGS)
def test_weights_reproduce_reference_init(name, golden_dir):
    """Same seed -> the drop-in modules hold bit-identical parameters to the reference's (names, shapes, values)."""
    fix = torch.load(os.path.join(golden_dir, f"{name}_predict.pt"))
    _, _, sd = build_state(name)
    assert set(sd) == set(fix["state_sha1"])
    bad = [k for k in sd if _sha(sd[k]) != fix["state_sha1"][k]]
    assert not bad, bad[:5]


@pytest.mark.parametrize("name", CONFIGS)
def test_oracle_matches_reference_fp32(name, golden_dir):
    fix = torch.load(os.path.join(golden_dir, f"{name}_predict.pt"))
    cfg, _, sd = build_state(name)
    x, param = O.make_inputs(cfg, fix["batch"], seed=fix["input_seed"])
    y, z = O.predict(sd, cfg, x, fix["steps"], param=param, to_x=True, return_latents=True)
    # same ATen CPU kernels in the same order: equal up to fp32 re-association inside oneDNN (thread count dependent)
    ez = O.rel_l2(z.flatten(0, 1), fix["latent_fp32"].flatten(0, 1)).max().item()
    ey = O.rel_l2(y[..., ::2, ::2].flatten(0, 1), fix["field_fp32_sub2"].flatten(0, 1)).max().item()
    assert ez < 2e-5 and ey < 2e-5, (ez, ey)


@pytest.mark.parametrize("name", CONFIGS)
def test_oracle_matches_reference_fp64(name, golden_dir):
    """fp64 oracle vs the fp64 copy of the reference: equal to ~1e-12 -> the restatement is the same function."""
    fix = torch.load(os.path.join(golden_dir, f"{name}_predict.pt"))
    cfg, _, sd = build_state(name)
    x, param = O.make_inputs(cfg, fix["batch"], seed=fix["input_seed"])
    sd64 = O.to_dtype(sd, torch.float64)
    y, z = O.predict(sd64, cfg, x.double(), fix["steps"], param=None if param is None else param.double(),
                     to_x=True, return_latents=True)
    ez = O.rel_l2(z.flatten(0, 1), fix["latent_fp64"].flatten(0, 1)).max().item()
    ey = O.rel_l2(y[..., ::2, ::2].flatten(0, 1), fix["field_fp64_sub2"].flatten(0, 1)).max().item()
    # the conditional model's sinusoidal embedding is fp32 in the reference -> ~1e-8 there
    tol = 1e-6 if name == "twophase_cond" else 1e-10
    assert ez < tol and ey < tol, (ez, ey)


def test_oracle_fourier_blocks(golden_dir):
    fix = torch.load(os.path.join(golden_dir, "fourier_blocks.pt"))
    x, emb = fix["x"], fix["emb"]
    # the reference allocates out_ft as torch.cfloat regardless of the input dtype (modules/basics.py:134-141), so even
    # its fp64 copy carries complex64 rounding in the spectral term; the oracle keeps complex128 -> agree to ~1e-8
    y = O.fourier_basic_block(x.double(), O.SD(O.to_dtype(fix["fourier_sd"], torch.float64)))
    assert O.rel_l2(y, fix["fourier_out_fp64"]).max().item() < 1e-7
    y = O.cond_fourier_basic_block(x.double(), emb.double(), O.SD(O.to_dtype(fix["cond_sd"], torch.float64)))
    assert O.rel_l2(y, fix["cond_out_fp64"]).max().item() < 1e-7
    y = O.fourier_basic_block(x, O.SD(fix["fourier_sd"]))
    assert O.rel_l2(y, fix["fourier_out"]).max().item() < 1e-5


@pytest.mark.parametrize("circular", [(True, True), (False, False), (False, True)])
def test_upsample_phase_decomposition(circular):
    """nearest x2 -> conv3x3 == four 2x2 convs of the source image (oracle/upsample_phases.py: the algebra behind the planned
    phase-decomposed up-sampling conv, 2.25x fewer MACs); zeros, circular and half-periodic padding, odd and even sizes"""
    import upsample_phases as U
    g = torch.Generator().manual_seed(5)
    for (H, W) in ((8, 8), (7, 15), (16, 5)):
        x = torch.randn(2, 6, H, W, generator=g, dtype=torch.float64)
        w = torch.randn(5, 6, 3, 3, generator=g, dtype=torch.float64)
        b = torch.randn(5, generator=g, dtype=torch.float64)
        ref = U.conv3x3_of_up2(x, w, b, circular)
        got = U.conv_up2_by_phases(x, w, b, circular)
        assert got.shape == ref.shape == (2, 5, 2 * H, 2 * W)
        assert (got - ref).abs().max().item() < 1e-12


@pytest.mark.parametrize("name", ["ns2d", "sw", "twophase", "twophase_cond"])
def test_oracle_training_gradients_vs_reference(name, golden_dir):
    """Gradient oracle (torch autograd of O.train_rollout, fp64) vs loss / parameter gradients of the unmodified reference's
    LatentDynamics.forward(z_in, z_out, F.smooth_l1_loss) (tests/golden/train_grads.pt, oracle/make_golden_train.py)."""
    import torch.nn.functional as F
    from lns_b200.configs import get_config
    from lns_b200.latent_dynamics import LatentDynamics
    fix = torch.load(os.path.join(golden_dir, "train_grads.pt"))[name]
    cfg = get_config(name)
    torch.manual_seed(1234)
    sd = O.randomize_zero_init(LatentDynamics(cfg).state_dict())
    sd64 = {k: v.double().requires_grad_(k.startswith("propagator.")) for k, v in sd.items()}
    z_in, z_out = O.train_inputs(cfg, fix["batch"], fix["t_out"], seed=0)
    param = torch.linspace(0.3, 0.9, fix["batch"]).double() if name == "twophase_cond" else None
    loss = F.smooth_l1_loss(O.train_rollout(sd64, cfg, z_in[:, 0].double(), fix["t_out"], param=param), z_out.double())
    loss.backward()
    # (the conditional model's sinusoidal embedding is computed in fp32 by the reference: ~1e-8 there)
    tol = 1e-6 if name == "twophase_cond" else 1e-9
    assert abs(float(loss) - fix["loss"]) < tol * 1e-3
    for k, n in fix["norm"].items():
        g = sd64["propagator." + k].grad
        assert abs(float(g.norm()) - n) <= tol * max(n, 1e-30), k
        pr = float((g * O.grad_probe(k, g.shape)).sum())
        assert abs(pr - fix["probe"][k]) <= tol * max(n, 1e-30) * g.numel() ** 0.5, k
        if k in fix["full"]:
            assert (g - fix["full"][k]).abs().max().item() <= tol * max(n, 1e-30), k
