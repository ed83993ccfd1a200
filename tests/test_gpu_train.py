"""GPU parity tests of the training rollout: the backward kernels of csrc/backward.cu one by one
against torch autograd in fp64, then LatentDynamics.forward(z_in, z_out, loss_fn).backward() against autograd of the oracle
(O.train_rollout, itself pinned to the unmodified reference's gradients by tests/golden/train_grads.pt).

Tolerances: 'fp32' mode -- loss and every parameter gradient within 1e-5 relative L2 of the fp64 oracle (measured <= 1.5e-6; torch's own
fp32 autograd is at 5e-7 ... 9e-7); 'fp16s' (forward on split-operand tcgen05 convs, fp32 storage) 5e-5 (measured <= 9.3e-6)."""
import pytest
import torch
import torch.nn.functional as F

import lns_oracle as O
from lns_b200.configs import get_config
from lns_b200.latent_dynamics import LatentDynamics

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def ops_mod():
    from lns_b200 import ops
    return ops


def rel(a, b):
    a, b = a.double().cpu().flatten(), b.double().cpu().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(t):  # NCHW torch -> Act fp32
    ops = ops_mod()
    B, C, H, W = t.shape
    return ops.Act(t.permute(0, 2, 3, 1).contiguous().float().to(DEV).reshape(-1), B, H, W, C)


PAD = {"circular": (1, 1), "zeros": (0, 0), "half": (0, 1)}


def ref_conv(x, w, b, mode, dil):
    k = w.shape[-1]
    p = dil * (k - 1) // 2
    if p:
        mh, mw = PAD[mode]
        x = F.pad(x, (p, p, 0, 0), mode="circular" if mw else "constant")
        x = F.pad(x, (0, 0, p, p), mode="circular" if mh else "constant")
    return F.conv2d(x, w, b, dilation=dil)


@pytest.mark.parametrize("shape", [(3, 8, 8, 128, 128, 3, 1, "circular"), (3, 8, 8, 128, 128, 3, 2, "circular"),
                                   (2, 7, 15, 64, 128, 3, 2, "zeros"), (2, 12, 24, 128, 128, 3, 3, "half"),
                                   (5, 8, 8, 16, 128, 1, 1, "zeros"), (5, 8, 8, 128, 16, 1, 1, "zeros"),
                                   (130, 8, 8, 128, 128, 3, 1, "circular")])
@pytest.mark.parametrize("with_pro", [False, True])
@pytest.mark.parametrize("tc", [False, True])
def test_conv_wgrad_and_bias_grad(shape, with_pro, tc):
    ops = ops_mod()
    B, H, W, Ci, Co, k, dil, mode = shape
    g = torch.Generator().manual_seed(B * 31 + Ci + dil)
    x = torch.randn(B, Ci, H, W, generator=g, dtype=torch.float64)
    w = (torch.randn(Co, Ci, k, k, generator=g, dtype=torch.float64) / (Ci * k * k) ** 0.5).requires_grad_(True)
    b = torch.randn(Co, generator=g, dtype=torch.float64).requires_grad_(True)
    dy = torch.randn(B, Co, H, W, generator=g, dtype=torch.float64)
    sc = torch.rand(B, Ci, generator=g, dtype=torch.float64) + 0.5
    sh = torch.randn(B, Ci, generator=g, dtype=torch.float64)
    xin = F.gelu(x * sc[:, :, None, None] + sh[:, :, None, None]) if with_pro else x
    ref_conv(xin, w, b, mode, dil).backward(dy)
    p = dil * (k - 1) // 2
    dW = torch.full((Co, Ci, k, k), 0.25, dtype=torch.float32, device=DEV)  # accumulates into existing values
    db = torch.full((Co,), -1.0, dtype=torch.float32, device=DEV)
    pro = (sc.float().to(DEV).contiguous(), sh.float().to(DEV).contiguous(), ops.ACT_GELU) if with_pro else None
    dya = nhwc(dy)
    # tc: TF32 mma.sync with hi + lo split operands; out_scale: the 1 / loss-scale factor of a scaled backward pass
    ops.conv2d_wgrad(nhwc(x), dya, dW, KH=k, KW=k, dil=dil, pad=(p, p, p, p), pad_mode=PAD[mode], pro=pro, tensor_core=tc,
                     out_scale=0.5 if tc else 1.0)
    ops.chan_sum_accum(dya, db, out_scale=0.5 if tc else 1.0)
    torch.cuda.synchronize()
    f = 0.5 if tc else 1.0
    assert rel(dW - 0.25, f * w.grad) < (5e-6 if tc else 3e-6)
    assert rel(db + 1.0, f * b.grad) < 3e-6


def test_absmax():
    ops = ops_mod()
    g = torch.Generator().manual_seed(2)
    t = torch.randn(3, 5, 16, 8, 8, generator=g) * 1e-5
    t[1, 2, 3, 4, 5] = -7.25e-3
    assert ops.absmax(t.to(DEV).contiguous()) == pytest.approx(7.25e-3, rel=1e-7)


@pytest.mark.parametrize("act", ["gelu", "silu"])
def test_act_bwd(act):
    ops = ops_mod()
    g = torch.Generator().manual_seed(3)
    x = (3 * torch.randn(4, 128, 8, 8, generator=g, dtype=torch.float64)).requires_grad_(True)
    dy = torch.randn(4, 128, 8, 8, generator=g, dtype=torch.float64)
    (F.gelu(x) if act == "gelu" else F.silu(x)).backward(dy)
    out = ops.act_bwd(nhwc(dy), nhwc(x.detach()), ops.ACT_GELU if act == "gelu" else ops.ACT_SILU)
    got = out.to_torch_nhwc().permute(0, 3, 1, 2)
    assert rel(got, x.grad) < 2e-6


@pytest.mark.parametrize("case", [(3, 128, 8, 8, 1, 1e-5), (3, 128, 8, 8, 32, 1e-6), (2, 128, 12, 24, 1, 1e-5),
                                  (2, 64, 7, 15, 32, 1e-6), (2, 16, 8, 8, 1, 1e-5)])
@pytest.mark.parametrize("skip", [False, True])
def test_group_norm_bwd(case, skip):
    ops = ops_mod()
    B, C, H, W, G, eps = case
    g = torch.Generator().manual_seed(C + G)
    x = (2 * torch.randn(B, C, H, W, generator=g, dtype=torch.float64) + 0.7).requires_grad_(True)
    gamma = (torch.rand(C, generator=g, dtype=torch.float64) + 0.5).requires_grad_(True)
    beta = torch.randn(C, generator=g, dtype=torch.float64).requires_grad_(True)
    dy = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    ds = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    F.group_norm(x, G, gamma, beta, eps).backward(dy)
    dgam = torch.zeros(C, dtype=torch.float32, device=DEV)
    dbet = torch.zeros(C, dtype=torch.float32, device=DEV)
    out = ops.group_norm_bwd(nhwc(x.detach()), nhwc(dy), G, eps, gamma.detach().float().to(DEV), dskip=nhwc(ds) if skip else None,
                             dgamma=dgam, dbeta=dbet, out_scale=1.0)
    got = out.to_torch_nhwc().permute(0, 3, 1, 2)
    want = x.grad + (ds if skip else 0)
    assert rel(got, want) < 3e-6
    assert rel(dgam, gamma.grad) < 3e-6 and rel(dbet, beta.grad) < 3e-6


_cache = {}


def build(name):
    if name not in _cache:
        cfg = get_config(name)
        torch.manual_seed(1234)
        model = LatentDynamics(cfg)
        sd = O.randomize_zero_init(model.state_dict())
        model.load_state_dict(sd, strict=True)
        _cache[name] = (cfg, model.to(DEV), sd)
    return _cache[name]


def oracle_grads(cfg, sd, z_in, z_out, loss_fn, param=None):
    sd64 = {k: v.double().requires_grad_(k.startswith("propagator.")) for k, v in sd.items()}
    z0 = z_in[:, 0].double().requires_grad_(True)
    loss = loss_fn(O.train_rollout(sd64, cfg, z0, z_out.shape[1], param=None if param is None else param.double()), z_out.double())
    loss.backward()
    return float(loss), {k[len("propagator."):]: v.grad for k, v in sd64.items() if k.startswith("propagator.")}, z0.grad


def rel_loss(pred, gt):  # training_utils.py:9-23 with reduce_all=True
    d = ((pred - gt) ** 2).sum(dim=(-1, -2, -3)) / (gt ** 2).sum(dim=(-1, -2, -3))
    return d.sqrt().mean()


@pytest.mark.parametrize("name,B,T", [("ns2d", 3, 3), ("sw", 2, 2), ("twophase", 2, 2), ("ns2d", 12, 2), ("twophase_cond", 3, 2)])
@pytest.mark.parametrize("mode,tol", [("fp32", 1e-5), ("fp16s", 5e-5)])
def test_training_rollout_gradients(name, B, T, mode, tol):
    """loss and d loss / d parameter of LatentDynamics.forward vs fp64 autograd of the oracle (every parameter of the propagator)."""
    ops = ops_mod()
    cfg, model, sd = build(name)
    z_in, z_out = O.train_inputs(cfg, B, T, seed=B)
    loss_fn = F.smooth_l1_loss if B != 12 else rel_loss
    param = torch.linspace(0.25, 0.85, B) if name == "twophase_cond" else None
    want_loss, want, _ = oracle_grads(cfg, sd, z_in, z_out, loss_fn, param)
    for p in model.parameters():
        p.requires_grad_(True)
    for p in model.autoencoder.parameters():
        p.requires_grad_(False)
    model.zero_grad(set_to_none=True)
    with ops.precision(mode):
        if param is None:
            loss = model(z_in.to(DEV), z_out.to(DEV), loss_fn)
        else:
            loss = model(z_in.to(DEV), z_out.to(DEV), param.to(DEV), loss_fn)
        loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - want_loss) <= tol * abs(want_loss)
    errs = []
    for k, p in model.propagator.named_parameters():
        assert p.grad is not None, k
        errs.append((rel(p.grad, want[k]), k, want[k].norm().item()))
    errs.sort(reverse=True)
    worst = (errs[0][1], errs[0][0])
    print("\n" + "\n".join(f"   {e:.2e} {k} |g|={n:.3e}" for e, k, n in errs[:6]))
    print(f"\n[{mode} {name} B={B} T={T}] loss {loss.item():.6f} (oracle {want_loss:.6f}); worst parameter-gradient rel-L2 "
          f"{worst[1]:.2e} ({worst[0]})")
    assert worst[1] < tol, worst
    assert all(p.grad is None for p in model.autoencoder.parameters())


def test_training_rollout_input_gradient_and_accumulation():
    """d loss / d z_in, and .grad accumulation over two backward passes (optimizer semantics)."""
    ops = ops_mod()
    cfg, model, sd = build("ns2d")
    z_in, z_out = O.train_inputs(cfg, 2, 2, seed=7)
    _, want, want_z = oracle_grads(cfg, sd, z_in, z_out, F.mse_loss)
    for p in model.propagator.parameters():
        p.requires_grad_(True)
    model.zero_grad(set_to_none=True)
    zi = z_in.to(DEV).requires_grad_(True)
    with ops.precision("fp32"):
        model(zi, z_out.to(DEV), F.mse_loss).backward()
        model(zi, z_out.to(DEV), F.mse_loss).backward()
    assert rel(zi.grad[:, 0], 2 * want_z) < 2e-5
    k, p = next(iter(model.propagator.named_parameters()))
    assert rel(p.grad, 2 * want[k]) < 2e-5


def test_training_step_reduces_loss():
    """Ten SGD-with-momentum steps on a fixed batch through the unmodified torch optimizer: the loss goes down, and predict()
    afterwards sees the updated weights (packed-filter caches follow the parameter versions)."""
    ops = ops_mod()
    cfg = get_config("ns2d")
    torch.manual_seed(5)
    model = LatentDynamics(cfg).to(DEV)
    for p in model.autoencoder.parameters():
        p.requires_grad_(False)
    opt = torch.optim.AdamW(model.propagator.parameters(), lr=2e-4)
    z_in, z_out = O.train_inputs(cfg, 8, 3, seed=1)
    z_in, z_out = z_in.to(DEV), (0.1 * z_out).to(DEV)
    losses = []
    with ops.precision("fp16s"):
        for _ in range(10):
            opt.zero_grad(set_to_none=True)
            loss = model(z_in, z_out, F.smooth_l1_loss)
            loss.backward()
            opt.step()
            losses.append(loss.item())
    print("\nlosses", " ".join(f"{v:.4f}" for v in losses))
    assert losses[-1] < 0.8 * losses[0]


def test_pixel_dot_and_scale_add():
    ops = ops_mod()
    g = torch.Generator().manual_seed(9)
    B, C, H, W = 3, 128, 7, 15
    dy = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    x = torch.randn(B, C, H, W, generator=g, dtype=torch.float64)
    sc = torch.randn(B, C, generator=g, dtype=torch.float64)
    out = torch.full((B * C,), 2.0, dtype=torch.float32, device=DEV)
    ops.pixel_dot(nhwc(dy), nhwc(x), out, accumulate=True)
    assert rel(out.view(B, C) - 2.0, (dy * x).sum(dim=(2, 3))) < 3e-6
    ops.pixel_dot(nhwc(dy), None, out, accumulate=False)
    assert rel(out.view(B, C), dy.sum(dim=(2, 3))) < 3e-6
    y = ops.scale_add(nhwc(dy), scale=sc.float().to(DEV).reshape(-1).contiguous(), skip=nhwc(x))
    assert rel(y.to_torch_nhwc().permute(0, 3, 1, 2), dy * sc[:, :, None, None] + x) < 1e-6


def test_graphed_train_step_matches_eager_and_follows_weight_updates():
    """The whole step as a CUDA graph: gradients equal the eager step's bit for bit, and after an optimizer update (outside the
    graph) the replay uses the NEW weights (the filter packing kernels are inside the captured sequence)."""
    ops = ops_mod()
    from lns_b200.train import GraphedTrainStep
    cfg = get_config("ns2d")
    torch.manual_seed(11)
    model = LatentDynamics(cfg).to(DEV)
    for p in model.autoencoder.parameters():
        p.requires_grad_(False)
    z_in, z_out = O.train_inputs(cfg, 32, 2, seed=4)
    z_in, z_out = z_in.to(DEV), (0.1 * z_out).to(DEV)
    opt = torch.optim.SGD(model.propagator.parameters(), lr=1e-2)

    def eager():
        model.zero_grad(set_to_none=True)
        with ops.precision("fp16s"):
            loss = model(z_in, z_out, F.smooth_l1_loss)
            loss.backward()
        return loss.item(), [p.grad.clone() for p in model.propagator.parameters()]

    step = GraphedTrainStep(model, z_in, z_out, F.smooth_l1_loss, precision="fp16s")
    for it in range(2):
        want_loss, want = eager()
        loss = step(z_in, z_out)
        torch.cuda.synchronize()
        assert loss.item() == want_loss
        for p, g in zip(model.propagator.parameters(), want):
            assert torch.equal(p.grad, g)
        opt.step()   # weights change: the next replay (and the next eager step) must see them
    z2 = z_in * 1.5  # new inputs through the static buffers
    loss = step(z2, z_out)
    with ops.precision("fp16s"):
        ref = model(z2, z_out, F.smooth_l1_loss)
    torch.cuda.synchronize()
    assert loss.item() == ref.item()
