"""GPU parity tests of the individual C-ABI kernels (called through lns_b200.ops -> ctypes -> liblns_b200.so) against
fp64 PyTorch CPU references of the same op.  fp32 results must agree to <= 1e-5 relative L2 (the fp32 validation
path's bar), bf16 results to bf16 rounding."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def relerr(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def ops_mod():
    from lns_b200 import ops
    return ops


def nhwc(x):  # NCHW cpu tensor -> contiguous NHWC cuda fp32 flat
    return x.permute(0, 2, 3, 1).contiguous()


def act_from(x_nchw, dtype=torch.float32):
    ops = ops_mod()
    B, C, H, W = x_nchw.shape
    t = nhwc(x_nchw).to(DEV).to(dtype).reshape(-1)
    return ops.Act(t, B, H, W, C)


def act_to_nchw(a):
    return a.to_torch_nhwc().float().permute(0, 3, 1, 2).cpu()


def ref_conv(x, w, b, stride, dil, pad, modes, virt=None):
    """fp64 reference of lns_conv2d's index map using torch ops."""
    x = x.double()
    if virt is not None:
        x = F.interpolate(x, size=virt, mode="nearest")
    pt, pb, pl, pr = pad
    if modes[1] == 1:
        x = F.pad(x, (pl, pr, 0, 0), mode="circular")
    else:
        x = F.pad(x, (pl, pr, 0, 0))
    if modes[0] == 1:
        x = F.pad(x, (0, 0, pt, pb), mode="circular")
    else:
        x = F.pad(x, (0, 0, pt, pb))
    return F.conv2d(x, w.double(), None if b is None else b.double(), stride=stride, dilation=dil)


class Holder:
    """minimal nn.Conv2d-like parameter holder for PackedFilter.of"""

    def __init__(self, w, b):
        self.weight = torch.nn.Parameter(w.to(DEV))
        self.bias = torch.nn.Parameter(b.to(DEV)) if b is not None else None


CONV_CASES = [
    # B, Cin, Cout, H, W, k, stride, dil, pad(t,b,l,r), modes(h,w), virt
    (2, 64, 64, 8, 8, 3, 1, 1, (1, 1, 1, 1), (1, 1), None),         # circular
    (2, 128, 128, 8, 8, 3, 1, 2, (2, 2, 2, 2), (1, 1), None),       # circular, dilation 2 (NS2d propagator)
    (3, 128, 128, 7, 15, 3, 1, 2, (2, 2, 2, 2), (0, 0), None),      # zeros, dilation 2, ragged (two-phase)
    (2, 128, 128, 12, 24, 3, 1, 3, (3, 3, 3, 3), (0, 1), None),     # half periodic, dilation 3 (shallow water)
    (2, 64, 64, 16, 16, 3, 2, 1, (1, 1, 1, 1), (1, 1), None),       # circular down-sample
    (2, 64, 64, 15, 30, 3, 2, 1, (0, 1, 0, 1), (0, 0), None),       # zeros down-sample pad (0,1,0,1)
    (2, 64, 64, 12, 24, 3, 2, 1, (1, 1, 1, 1), (0, 1), None),       # half-periodic stride 2 pad 1
    (2, 64, 64, 8, 8, 3, 1, 1, (1, 1, 1, 1), (1, 1), (16, 16)),     # nearest x2 folded
    (2, 64, 64, 14, 30, 3, 1, 1, (1, 1, 1, 1), (0, 0), (31, 61)),   # nearest to odd size
    (2, 64, 128, 9, 9, 1, 1, 1, (0, 0, 0, 0), (0, 0), None),        # 1x1 channel_up
    (1, 128, 16, 8, 8, 1, 1, 1, (0, 0, 0, 0), (0, 0), None),        # latent projection (N=16)
    (2, 64, 512, 16, 16, 1, 1, 1, (0, 0, 0, 0), (0, 0), None),      # FABlock in_proj (N tiles)
    (2, 512, 64, 16, 16, 1, 1, 1, (0, 0, 0, 0), (0, 0), None),      # FABlock to_out (8 K slabs)
    (5, 64, 64, 64, 64, 3, 1, 1, (1, 1, 1, 1), (1, 1), None),       # many tiles
]


def _conv_case(case, seed=0):
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    b = torch.randn(Cout, generator=g) * 0.1
    return x, w, b


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_simt_fp32(case):
    ops = ops_mod()
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    x, w, b = _conv_case(case)
    ref = ref_conv(x, w, b, stride, dil, pad, modes, virt)
    with ops.precision("fp32"):
        y = ops.conv2d(act_from(x), ops.PackedFilter.of(Holder(w, b).weight, Holder(w, b).bias), stride=stride, dil=dil,
                       pad=pad, pad_mode=modes, virt=virt, engine=ops.ENGINE_SIMT)
    torch.cuda.synchronize()
    assert tuple(act_to_nchw(y).shape) == tuple(ref.shape)
    assert relerr(act_to_nchw(y), ref) < 2e-6


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_umma_bf16(case):
    """tcgen05 engine vs an fp64 conv of the SAME bf16-rounded operands: only fp32 accumulation order differs."""
    ops = ops_mod()
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    x, w, b = _conv_case(case, seed=1)
    xb, wb = x.bfloat16().float(), w.bfloat16().float()
    ref = ref_conv(xb, wb, b, stride, dil, pad, modes, virt)
    h = Holder(w, b)
    with ops.precision("bf16"):
        y = ops.conv2d(act_from(x, torch.bfloat16), ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil,
                       pad=pad, pad_mode=modes, virt=virt, engine=ops.ENGINE_UMMA, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 5e-6
    with ops.precision("bf16"):
        y16 = ops.conv2d(act_from(x, torch.bfloat16), ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil,
                         pad=pad, pad_mode=modes, virt=virt, engine=ops.ENGINE_UMMA)
    torch.cuda.synchronize()
    assert y16.t.dtype == torch.bfloat16
    assert relerr(act_to_nchw(y16), ref) < 4e-3  # bf16 output rounding (2^-9 per element)


@pytest.mark.parametrize("engine", ["simt", "umma"])
def test_conv_epilogue_and_prologue(engine):
    """bias + per-sample bias + pre-activation addend + GELU + residual; SiLU(affine(x)) prologue with zero padding"""
    ops = ops_mod()
    B, C, H, W = 3, 64, 10, 12
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) / math.sqrt(C * 9)
    b = torch.randn(C, generator=g) * 0.1
    sb = torch.randn(B, C, generator=g) * 0.1
    pre = torch.randn(B, C, H, W, generator=g)
    res = torch.randn(B, C, H, W, generator=g)
    sc = torch.rand(B, C, generator=g) + 0.5
    sh = torch.randn(B, C, generator=g) * 0.2
    dt = torch.float32 if engine == "simt" else torch.bfloat16
    rd = (lambda t: t) if engine == "simt" else (lambda t: t.bfloat16().float())
    xin = F.silu(rd(x).double() * sc[:, :, None, None].double() + sh[:, :, None, None].double())
    if engine == "umma":
        xin = xin.float().bfloat16().double()  # the umma path materialises the normalised activation in bf16
    ref = F.conv2d(F.pad(xin, (1, 1, 1, 1)), rd(w).double(), b.double())
    ref = F.gelu(ref + sb[:, :, None, None].double() + rd(pre).double()) + rd(res).double()
    h = Holder(w, b)
    with ops.precision("fp32" if engine == "simt" else "bf16"):
        y = ops.conv2d(act_from(x, dt), ops.PackedFilter.of(h.weight, h.bias), pad=(1, 1, 1, 1),
                       sample_bias=sb.to(DEV), pro=(sc.to(DEV).reshape(-1), sh.to(DEV).reshape(-1), ops.ACT_SILU),
                       act=ops.ACT_GELU, pre_add=act_from(pre, dt), residual=act_from(res, dt),
                       out_dtype=torch.float32,
                       engine=ops.ENGINE_SIMT if engine == "simt" else ops.ENGINE_UMMA)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < (2e-6 if engine == "simt" else 2e-5)


def test_conv_nchw_ends():
    """Cin=1 NCHW fp32 lift (+Swish) and Cout=1 NCHW fp32 projection with a batch-strided output"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(2)
    x = torch.randn(3, 1, 16, 16, generator=g)
    w = torch.randn(64, 1, 1, 1, generator=g)
    b = torch.randn(64, generator=g)
    h = Holder(w, b)
    with ops.precision("fp32"):
        y = ops.conv2d(ops.Act.from_nchw(x.to(DEV)), ops.PackedFilter.of(h.weight, h.bias), act=ops.ACT_SILU)
    ref = F.silu(F.conv2d(x.double(), w.double(), b.double()))
    assert relerr(act_to_nchw(y), ref) < 2e-6
    w2 = torch.randn(1, 64, 1, 1, generator=g) / 8
    b2 = torch.randn(1, generator=g)
    h2 = Holder(w2, b2)
    out = torch.zeros(3, 2, 1, 16, 16, device=DEV)  # [B, K=2, C, H, W]; write slot k=1
    dst = ops.Act(out.view(-1)[256:], 3, 16, 16, 1, bstride=2 * 256, layout=ops.NCHW)
    with ops.precision("fp32"):
        ops.conv2d(y, ops.PackedFilter.of(h2.weight, h2.bias), out=dst, out_layout=ops.NCHW)
    ref2 = F.conv2d(ref, w2.double(), b2.double())
    assert relerr(out[:, 1].cpu(), ref2) < 2e-6
    assert float(out[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,groups,eps", [((3, 64, 8, 8), 32, 1e-6), ((2, 128, 7, 15), 1, 1e-5),
                                               ((2, 64, 64, 64), 8, 1e-5), ((2, 512, 16, 16), 512, 1e-5),
                                               ((2, 128, 1, 1), 1, 1e-5), ((1, 64, 96, 192), 32, 1e-6)])
def test_group_norm(shape, groups, eps, dtype):
    ops = ops_mod()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(*shape, generator=g) * 1.7 + 0.3
    gamma = torch.rand(shape[1], generator=g) + 0.5
    beta = torch.randn(shape[1], generator=g)
    xr = x.to(dtype).double()
    ref = F.silu(F.group_norm(xr, groups, gamma.double(), beta.double(), eps))
    a = act_from(x, dtype)
    s, t = ops.group_norm_affine(a, groups, eps, gamma.to(DEV), beta.to(DEV))
    y = ops.affine_act(a, s, t, ops.ACT_SILU, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 3e-6


def test_group_norm_prescale():
    """GroupNorm(1) of x*(1+g) from the statistics of x (conditional propagator gate)"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 128, 7, 15, generator=g)
    gate = torch.randn(2, 128, generator=g) * 0.3
    gamma, beta = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    ref = F.group_norm(x.double() * (1 + gate.double())[:, :, None, None], 1, gamma.double(), beta.double(), 1e-5)
    a = act_from(x)
    s, t = ops.group_norm_affine(a, 1, 1e-5, gamma.to(DEV), beta.to(DEV), prescale=(1 + gate).to(DEV).reshape(-1))
    y = ops.affine_act(a, s, t)
    assert relerr(act_to_nchw(y), ref) < 3e-6


def test_layernorm_pe():
    ops = ops_mod()
    g = torch.Generator().manual_seed(6)
    x = torch.randn(3, 128, 8, 8, generator=g)
    gamma, beta = torch.rand(128, generator=g) + 0.5, torch.randn(128, generator=g)
    pe = torch.randn(1, 80, 128, generator=g) * 0.02
    tok = x.permute(0, 2, 3, 1).reshape(3, 64, 128).double()
    ref = F.layer_norm(tok, (128,), gamma.double(), beta.double(), 1e-5) + pe[:, :64].double()
    y = ops.layernorm(act_from(x), gamma.to(DEV), beta.to(DEV), 1e-5, pe=pe.to(DEV))
    assert relerr(y.as_tokens().cpu(), ref) < 2e-6


@pytest.mark.parametrize("n_hw", [(8, 8), (12, 24), (7, 15)])
def test_attention(n_hw):
    ops = ops_mod()
    H, W = n_hw
    n, heads, dh = H * W, 8, 64
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(2, n, 3 * heads * dh, generator=g)
    q, k, v = [t.view(2, n, heads, dh).transpose(1, 2).double() for t in qkv.split(heads * dh, dim=-1)]
    attn = F.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, dim=-1)
    ref = (attn @ v).transpose(1, 2).reshape(2, n, heads * dh)
    a = ops.Act(qkv.to(DEV).reshape(-1), 2, H, W, 3 * heads * dh)
    y = ops.attention(a, heads, dh, dh ** -0.5)
    assert relerr(y.as_tokens().cpu(), ref) < 2e-6


def test_axis_mean_lowrank_contract():
    ops = ops_mod()
    import lns_oracle as O
    g = torch.Generator().manual_seed(8)
    B, H, W, heads, ch = 2, 12, 24, 8, 64
    x = torch.randn(B, 64, H, W, generator=g)
    a = act_from(x)
    m1 = ops.axis_mean(a, axis=1).t.view(B, H, 64).cpu()
    m0 = ops.axis_mean(a, axis=0).t.view(B, W, 64).cpu()
    assert relerr(m1, x.double().mean(dim=3).permute(0, 2, 1)) < 1e-6
    assert relerr(m0, x.double().mean(dim=2).permute(0, 2, 1)) < 1e-6
    # low-rank kernel vs the oracle's (reference modules/factorized_attention.py:43-69)
    from modules.factorized_attention import LowRankKernel
    torch.manual_seed(0)
    lrk = LowRankKernel(64, 128, heads, use_rotary_emb=True).to(DEV)
    u = torch.randn(B, W, 64, generator=g)
    sd = {k: v.cpu().double() for k, v in lrk.state_dict().items()}
    ref = O.low_rank_kernel(u.double(), O.SD(sd), heads)
    K = lrk._fwd(ops.Act(u.to(DEV).reshape(-1), B, W, 1, 64))
    assert relerr(K.cpu(), ref) < 3e-6
    # axial contractions
    up = torch.randn(B, heads * ch, H, W, generator=g)
    kx = torch.randn(B, heads, H, H, generator=g)
    ky = torch.randn(B, heads, W, W, generator=g)
    u5 = up.double().view(B, heads, ch, H, W)
    r1 = torch.einsum("bhij,bhcjm->bhcim", kx.double(), u5)
    r2 = torch.einsum("bhlm,bhcim->bhcil", ky.double(), r1)
    a1 = ops.axial_contract(act_from(up), kx.to(DEV).contiguous(), heads, axis=0)
    a2 = ops.axial_contract(a1, ky.to(DEV).contiguous(), heads, axis=1)
    assert relerr(act_to_nchw(a1), r1.reshape(B, heads * ch, H, W)) < 2e-6
    assert relerr(act_to_nchw(a2), r2.reshape(B, heads * ch, H, W)) < 3e-6


def test_layout_roundtrip_and_embedding():
    ops = ops_mod()
    import lns_oracle as O
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 5, 7, 11, generator=g)
    a = ops.nchw_to_act(x.to(DEV), torch.float32)
    assert torch.equal(a.to_torch_nhwc().cpu(), x.permute(0, 2, 3, 1))
    assert torch.equal(a.to_nchw().cpu(), x)
    p = torch.rand(6, generator=g)
    e = ops.fourier_embedding(p.to(DEV), 64)
    assert relerr(e.cpu(), O.fourier_embedding(p, 64)) < 1e-6


def test_spectral_blocks_vs_golden(golden_dir):
    """FourierBasicBlock / CondFourierBasicBlock (truncated-DFT kernels) vs outputs of the reference modules"""
    import os
    ops = ops_mod()
    from modules.basics import FourierBasicBlock
    from modules.fourier_cond import CondFourierBasicBlock
    fix = torch.load(os.path.join(golden_dir, "fourier_blocks.pt"))
    blk = FourierBasicBlock(8, 8, modes=[4, 5])
    blk.load_state_dict(fix["fourier_sd"], strict=True)
    cblk = CondFourierBasicBlock(8, 8, modes=[4, 5])
    cblk.load_state_dict(fix["cond_sd"], strict=True)
    blk, cblk = blk.to(DEV).eval(), cblk.to(DEV).eval()
    with torch.no_grad(), ops.precision("fp32"):
        y = blk(fix["x"].to(DEV))
        yc = cblk(fix["x"].to(DEV), fix["emb"].to(DEV))
    assert relerr(y.cpu(), fix["fourier_out_fp64"]) < 1e-5
    assert relerr(yc.cpu(), fix["cond_out_fp64"]) < 1e-5


@pytest.mark.parametrize("H,W,m1,m2", [(64, 64, 16, 16), (61, 121, 16, 31), (32, 32, 6, 6)])
def test_spectral_conv_shapes(H, W, m1, m2):
    ops = ops_mod()
    import lns_oracle as O
    from modules.basics import SpectralConv2d
    torch.manual_seed(1)
    m = SpectralConv2d(16, 16, m1, m2).to(DEV)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(2, 16, H, W, generator=g)
    sd = {k: v.cpu().double() for k, v in m.state_dict().items()}
    ref = O.spectral_conv2d(x.double(), O.SD(sd))
    with torch.no_grad(), ops.precision("fp32"):
        y = m(x.to(DEV))
    assert relerr(y.cpu(), ref) < 1e-5


@pytest.mark.parametrize("H,W", [(16, 16), (32, 32), (24, 48), (48, 96), (12, 24), (7, 15)])
def test_axial_contract_tensor_core_bf16(H, W):
    """mma.sync path (bf16 in/out) vs fp64 einsum of the same bf16-rounded operands"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(21)
    B, heads, ch = 3, 8, 64
    up = torch.randn(B, heads * ch, H, W, generator=g)
    kx = torch.randn(B, heads, H, H, generator=g)
    ky = torch.randn(B, heads, W, W, generator=g)
    u5 = up.bfloat16().double().view(B, heads, ch, H, W)
    r1 = torch.einsum("bhij,bhcjm->bhcim", kx.bfloat16().double(), u5)
    a1 = ops.axial_contract(act_from(up, torch.bfloat16), kx.to(DEV).contiguous(), heads, axis=0)
    assert a1.t.dtype == torch.bfloat16
    assert relerr(act_to_nchw(a1), r1.reshape(B, heads * ch, H, W)) < 4e-3
    r1b = act_to_nchw(a1).double().view(B, heads, ch, H, W)  # feed the kernel's own (bf16) intermediate to stage 2
    r2 = torch.einsum("bhlm,bhcim->bhcil", ky.bfloat16().double(), r1b)
    a2 = ops.axial_contract(a1, ky.to(DEV).contiguous(), heads, axis=1)
    assert relerr(act_to_nchw(a2), r2.reshape(B, heads * ch, H, W)) < 4e-3


HALO_CASES = [
    # B, Cout, H, W, dil, modes(h,w), virt
    (2, 64, 64, 64, 1, (1, 1), None),        # NS2d full resolution, circular
    (3, 64, 32, 32, 1, (1, 1), None),
    (2, 64, 16, 16, 1, (1, 1), (32, 32)),    # nearest x2 folded into the halo fill
    (2, 128, 16, 16, 1, (1, 1), None),       # 64 -> 128 (encoder ResBlock)
    (2, 64, 28, 60, 1, (0, 0), (61, 121)),   # two-phase: general nearest to an odd size, zero padding, ragged tiles
    (2, 64, 61, 121, 1, (0, 0), None),       # ragged tiles in both directions
    (2, 64, 48, 96, 1, (0, 1), None),        # shallow water: zeros in H, circular in W
    (1, 64, 96, 192, 1, (0, 1), None),
    (2, 64, 24, 24, 2, (1, 1), None),        # dilation 2
    (2, 64, 24, 40, 3, (0, 1), None),        # dilation 3
    (150, 64, 16, 16, 1, (1, 1), None),      # more tiles than SMs: persistent loop, ring wrap-around
]


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_halo_bf16(case):
    """halo engine (resident filter, shifted-descriptor taps) vs fp64 conv of the same bf16-rounded operands, with the
    full epilogue (bias, GELU, residual)"""
    ops = ops_mod()
    B, Cout, H, W, dil, modes, virt = case
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(Cout, 64, 3, 3, generator=g) / 24.0
    b = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = virt if virt is not None else (H, W)
    res = torch.randn(B, Cout, Ho, Wo, generator=g)
    ref = ref_conv(x.bfloat16().float(), w.bfloat16().float(), b, 1, dil, (dil,) * 4, modes, virt)
    ref = F.gelu(ref) + res.bfloat16().double()
    h = Holder(w, b)
    with ops.precision("bf16"):
        y = ops.conv2d(act_from(x, torch.bfloat16), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4,
                       pad_mode=modes, virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.bfloat16),
                       engine=ops.ENGINE_HALO, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert tuple(act_to_nchw(y).shape) == tuple(ref.shape)
    assert relerr(act_to_nchw(y), ref) < 5e-6
    with ops.precision("bf16"):  # bf16 output: the shared-memory staged, full-line store path of the epilogue
        y16 = ops.conv2d(act_from(x, torch.bfloat16), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4,
                         pad_mode=modes, virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.bfloat16),
                         engine=ops.ENGINE_HALO)
    torch.cuda.synchronize()
    assert y16.t.dtype == torch.bfloat16
    assert torch.equal(act_to_nchw(y16), act_to_nchw(y).bfloat16().float())  # same values, rounded once


@pytest.mark.parametrize("n_hw", [(8, 8), (12, 24), (7, 15)])
def test_attention_tensor_core_bf16(n_hw):
    """mma.sync flash-style attention (bf16 in/out) vs fp64 softmax attention of the same bf16-rounded q, k, v"""
    ops = ops_mod()
    H, W = n_hw
    n, heads, dh = H * W, 8, 64
    g = torch.Generator().manual_seed(17)
    qkv = torch.randn(3, n, 3 * heads * dh, generator=g)
    qr = qkv.bfloat16().double()
    q, k, v = [t.view(3, n, heads, dh).transpose(1, 2) for t in qr.split(heads * dh, dim=-1)]
    attn = F.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, dim=-1)
    ref = (attn @ v).transpose(1, 2).reshape(3, n, heads * dh)
    a = ops.Act(qkv.to(DEV).bfloat16().reshape(-1), 3, H, W, 3 * heads * dh)
    y = ops.attention(a, heads, dh, dh ** -0.5)
    assert y.t.dtype == torch.bfloat16
    assert relerr(y.as_tokens().float().cpu(), ref) < 6e-3  # P and the output are rounded to bf16


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pointwise_projection(dtype):
    """GroupNorm -> Swish -> Conv1x1(64 -> 3) decoder tail as one kernel, NCHW fp32 output into a strided slot"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(23)
    B, C, H, W, Co = 3, 64, 20, 33, 3
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Co, C, 1, 1, generator=g) / 8
    b = torch.randn(Co, generator=g)
    sc, sh = torch.rand(B, C, generator=g) + 0.5, torch.randn(B, C, generator=g) * 0.2
    xin = F.silu(x.to(dtype).double() * sc[:, :, None, None].double() + sh[:, :, None, None].double())
    ref = F.conv2d(xin, w.double(), b.double())
    h = Holder(w, b)
    out = torch.zeros(B, 2, Co, H, W, device=DEV)
    dst = ops.Act(out.view(-1)[Co * H * W:], B, H, W, Co, bstride=2 * Co * H * W, layout=ops.NCHW)
    ops.conv2d(act_from(x, dtype), ops.PackedFilter.of(h.weight, h.bias),
               pro=(sc.to(DEV).reshape(-1), sh.to(DEV).reshape(-1), ops.ACT_SILU), out=dst, out_layout=ops.NCHW)
    # bf16 storage: Swish through tanh.approx.f32 (one special-function op per element, abs error < 2^-10.9)
    assert relerr(out[:, 1].cpu(), ref) < (3e-6 if dtype == torch.float32 else 3e-4)
    assert float(out[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("H,W", [(16, 16), (32, 32), (24, 48), (12, 20), (8, 8), (30, 17)])
def test_fablock_fused_vs_unfused_and_oracle(H, W):
    """Fused FABlock2D core (shared-memory resident u_phi) == the unfused kernel sequence (bf16 tolerance) and agrees
    with the fp64 oracle of the reference block at bf16 accuracy."""
    ops = ops_mod()
    import lns_oracle as O
    from modules.factorized_attention import FABlock2D
    torch.manual_seed(3)
    blk = FABlock2D(64, 64, 64, 8, 64).to(DEV).eval()
    g = torch.Generator().manual_seed(41)
    x = torch.randn(3, 64, H, W, generator=g)
    sd = {k: v.cpu().double() for k, v in blk.state_dict().items()}
    ref = O.fa_block(x.double(), O.SD(sd))
    a = act_from(x, torch.bfloat16)
    assert ops.fablock_core_supported(a, 64)
    with torch.no_grad(), ops.precision("bf16"):
        fused = act_to_nchw(blk._fwd(a))  # whole-block kernel (fablock_full) when H, W <= 32, else the fused core
        orig_full, orig = ops.fablock_full_supported, ops.fablock_core_supported
        ops.fablock_full_supported = lambda *aa, **kk: False
        try:
            core = act_to_nchw(blk._fwd(a))  # fused core + two conv launches
            ops.fablock_core_supported = lambda *aa, **kk: False
            unfused = act_to_nchw(blk._fwd(a))
        finally:
            ops.fablock_full_supported, ops.fablock_core_supported = orig_full, orig
    e_f, e_c, e_u = relerr(fused, ref), relerr(core, ref), relerr(unfused, ref)
    print(f"[FABlock2D {H}x{W} bf16] whole-block kernel vs fp64 oracle {e_f:.2e}, fused core {e_c:.2e}, unfused {e_u:.2e}, "
          f"whole-block vs unfused {relerr(fused, unfused):.2e}")
    assert e_f < 3e-2 and e_c < 3e-2 and e_u < 3e-2
    assert relerr(fused, unfused) < 2e-2 and relerr(core, unfused) < 2e-2


@pytest.mark.parametrize("n", [16, 32, 24, 48, 15])
def test_lowrank_kernel_tensor_core_bf16(n):
    """mma.sync LowRankKernel (bf16 q|k) vs the fp64 oracle evaluated on the same bf16-rounded q|k"""
    ops = ops_mod()
    import lns_oracle as O
    from modules.factorized_attention import LowRankKernel
    torch.manual_seed(0)
    lrk = LowRankKernel(64, 128, 8, use_rotary_emb=True).to(DEV)
    g = torch.Generator().manual_seed(51)
    qk = torch.randn(3, n, 2048, generator=g)
    cos_t, sin_t = lrk._tables(n, torch.device(DEV))
    K = ops.lowrank_kernel(ops.Act(qk.to(DEV).bfloat16().reshape(-1), 3, n, 1, 2048), 8, 128, cos_t, sin_t, 1.0)
    # reference: rotary + q k^T in fp64 on the bf16-rounded projections
    qr = qk.bfloat16().double()
    q, k = qr.split(1024, dim=-1)
    q = q.view(3, n, 8, 128).transpose(1, 2)
    k = k.view(3, n, 8, 128).transpose(1, 2)
    pos = torch.linspace(0, 1, n).double() * 64.0
    fr = pos[:, None] * lrk.pos_emb.inv_freq.cpu().double()[None, :]
    fr = torch.cat((fr, fr), dim=-1)[None, None]
    ref = torch.einsum("bhid,bhjd->bhij", O._rotary(q, fr), O._rotary(k, fr))
    assert relerr(K.cpu(), ref) < 5e-3  # rotated q, k are rounded to bf16 before the tensor-core product


def test_conv_composition_3x3_then_1x1():
    """conv1x1(conv3x3(x)) as ONE conv with the composed filter (decoder tail of the NS2d autoencoder)"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(61)
    x = torch.randn(2, 64, 16, 16, generator=g)
    w1, b1 = torch.randn(64, 64, 3, 3, generator=g) / 24, torch.randn(64, generator=g) * 0.1
    w2, b2 = torch.randn(64, 64, 1, 1, generator=g) / 8, torch.randn(64, generator=g) * 0.1
    ref = F.conv2d(ref_conv(x, w1, b1, 1, 1, (1, 1, 1, 1), (1, 1)), w2.double(), b2.double())
    h1, h2 = Holder(w1, b1), Holder(w2, b2)
    filt = ops.composed_filter(ops.PackedFilter.of(h1.weight, h1.bias), ops.PackedFilter.of(h2.weight, h2.bias))
    with ops.precision("fp32"):
        y = ops.conv2d(act_from(x), filt, pad=(1, 1, 1, 1), pad_mode=(1, 1), engine=ops.ENGINE_SIMT)
    assert relerr(act_to_nchw(y), ref) < 3e-6


@pytest.mark.parametrize("shape,groups", [((5, 128, 8, 8), 1), ((5, 128, 8, 8), 32), ((3, 64, 7, 15), 32)])
def test_group_norm_act_fused_small(shape, groups):
    """single-kernel statistics + normalise + activation == the two-kernel path, bit for bit"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(71)
    x = torch.randn(*shape, generator=g) * 1.3 + 0.2
    gamma, beta = (torch.rand(shape[1], generator=g) + 0.5).to(DEV), torch.randn(shape[1], generator=g).to(DEV)
    a = act_from(x, torch.bfloat16)
    fused = ops.LazyNorm(a, groups, 1e-5, gamma, beta, None, ops.ACT_GELU).materialize()
    s, t = ops.group_norm_affine(a, groups, 1e-5, gamma, beta)
    two = ops.affine_act(a, s, t, ops.ACT_GELU)
    assert torch.equal(fused.t, two.t)
    ref = F.gelu(F.group_norm(x.bfloat16().double(), groups, gamma.cpu().double(), beta.cpu().double(), 1e-5))
    assert relerr(act_to_nchw(fused), ref) < 4e-3


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_umma_tf32(case):
    """tcgen05.mma.kind::tf32 engine vs fp64 conv of the SAME TF32-rounded operands (fp32 accumulation order only)"""
    ops = ops_mod()
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    x, w, b = _conv_case(case, seed=2)

    def rt(t):  # round to nearest TF32 (10-bit mantissa), ties away from zero like cvt.rna
        i = t.contiguous().view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)
    ref = ref_conv(rt(x), rt(w), b, stride, dil, pad, modes, virt)
    h = Holder(w, b)
    with ops.precision("tf32"):
        a = act_from(rt(x))
        a.tf32 = True
        y = ops.conv2d(a, ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil, pad=pad, pad_mode=modes,
                       virt=virt, engine=ops.ENGINE_UMMA, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 5e-6
    with ops.precision("tf32"):
        y2 = ops.conv2d(a, ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil, pad=pad, pad_mode=modes, virt=virt)
    assert y2.tf32 and y2.t.dtype == torch.float32
    assert torch.equal(act_to_nchw(y2), rt(act_to_nchw(y)))  # stored values are the TF32 roundings


# ---- fp16 precision mode: the same 16-bit kernels with IEEE-half operands (LNS_F16) -------------------------------------
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_umma_f16(case):
    """tcgen05 gather engine with f16 operands vs an fp64 conv of the SAME f16-rounded operands"""
    ops = ops_mod()
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    x, w, b = _conv_case(case, seed=1)
    ref = ref_conv(x.half().float(), w.half().float(), b, stride, dil, pad, modes, virt)
    h = Holder(w, b)
    with ops.precision("fp16"):
        y = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil,
                       pad=pad, pad_mode=modes, virt=virt, engine=ops.ENGINE_UMMA, out_dtype=torch.float32)
        y16 = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil,
                         pad=pad, pad_mode=modes, virt=virt, engine=ops.ENGINE_UMMA)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 5e-6
    assert y16.t.dtype == torch.float16
    assert relerr(act_to_nchw(y16), ref) < 5e-4  # f16 output rounding (2^-12 per element)


@pytest.mark.parametrize("case", HALO_CASES)
def test_conv_halo_f16(case):
    """halo engine with f16 operands (bias folded into the GEMM as f16 hi+lo), full epilogue"""
    ops = ops_mod()
    B, Cout, H, W, dil, modes, virt = case
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(Cout, 64, 3, 3, generator=g) / 24.0
    b = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = virt if virt is not None else (H, W)
    res = torch.randn(B, Cout, Ho, Wo, generator=g)
    ref = ref_conv(x.half().float(), w.half().float(), b, 1, dil, (dil,) * 4, modes, virt)
    ref = F.gelu(ref) + res.half().double()
    h = Holder(w, b)
    with ops.precision("fp16"):
        y = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4,
                       pad_mode=modes, virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.float16),
                       engine=ops.ENGINE_HALO, out_dtype=torch.float32)
        y16 = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4,
                         pad_mode=modes, virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.float16),
                         engine=ops.ENGINE_HALO)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 5e-6
    assert y16.t.dtype == torch.float16
    assert torch.equal(act_to_nchw(y16), act_to_nchw(y).half().float())  # same values, rounded once


def test_f16_saturates_instead_of_overflowing():
    """conversions to LNS_F16 clamp to +-65504 (an inf would poison every norm statistic downstream)"""
    ops = ops_mod()
    x = torch.full((1, 64, 4, 4), 300.0)
    a = act_from(x, torch.float32)
    scale = torch.full((64,), 1000.0, device=DEV)
    shift = torch.zeros(64, device=DEV)
    y = ops.affine_act(a, scale, shift, ops.ACT_NONE, out_dtype=torch.float16)
    torch.cuda.synchronize()
    assert torch.isfinite(y.t).all() and float(y.t.float().max()) == 65504.0


@pytest.mark.parametrize("n_hw", [(8, 8), (12, 24), (7, 15)])
def test_attention_tensor_core_f16(n_hw):
    ops = ops_mod()
    H, W = n_hw
    n, heads, dh = H * W, 8, 64
    g = torch.Generator().manual_seed(17)
    qkv = torch.randn(3, n, 3 * heads * dh, generator=g)
    qr = qkv.half().double()
    q, k, v = [t.view(3, n, heads, dh).transpose(1, 2) for t in qr.split(heads * dh, dim=-1)]
    attn = F.softmax(q @ k.transpose(-1, -2) * dh ** -0.5, dim=-1)
    ref = (attn @ v).transpose(1, 2).reshape(3, n, heads * dh)
    a = ops.Act(qkv.to(DEV).half().reshape(-1), 3, H, W, 3 * heads * dh)
    y = ops.attention(a, heads, dh, dh ** -0.5)
    assert y.t.dtype == torch.float16
    assert relerr(y.as_tokens().float().cpu(), ref) < 8e-4  # P and the output are rounded to f16


@pytest.mark.parametrize("H,W", [(16, 16), (32, 32), (24, 48), (12, 20)])
def test_fablock_fused_f16(H, W):
    """Fused FABlock2D core with f16 storage / operands vs the fp64 oracle of the reference block, and vs the unfused path"""
    ops = ops_mod()
    import lns_oracle as O
    from modules.factorized_attention import FABlock2D
    torch.manual_seed(3)
    blk = FABlock2D(64, 64, 64, 8, 64).to(DEV).eval()
    g = torch.Generator().manual_seed(41)
    x = torch.randn(3, 64, H, W, generator=g)
    sd = {k: v.cpu().double() for k, v in blk.state_dict().items()}
    ref = O.fa_block(x.double(), O.SD(sd))
    a = act_from(x, torch.float16)
    assert ops.fablock_core_supported(a, 64)
    with torch.no_grad(), ops.precision("fp16"):
        fused = act_to_nchw(blk._fwd(a))
        orig_full, orig = ops.fablock_full_supported, ops.fablock_core_supported
        ops.fablock_full_supported = lambda *aa, **kk: False
        try:
            core = act_to_nchw(blk._fwd(a))
            ops.fablock_core_supported = lambda *aa, **kk: False
            unfused = act_to_nchw(blk._fwd(a))
        finally:
            ops.fablock_full_supported, ops.fablock_core_supported = orig_full, orig
    e_f, e_c, e_u = relerr(fused, ref), relerr(core, ref), relerr(unfused, ref)
    print(f"[FABlock2D {H}x{W} f16] whole-block kernel vs fp64 oracle {e_f:.2e}, fused core {e_c:.2e}, unfused {e_u:.2e}")
    assert e_f < 4e-3 and e_c < 4e-3 and e_u < 4e-3


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("H,W,B", [(32, 32, 5), (16, 16, 9), (12, 20, 3), (48, 24, 2)])
def test_fa_axis_kernel_vs_fp64(H, W, B, prec):
    """One-kernel pooled branch (to_in -> PoolingReducer -> to_qk -> rotary -> q k^T) vs the same chain in fp64 torch"""
    ops = ops_mod()
    from modules.factorized_attention import FABlock2D
    torch.manual_seed(5)
    blk = FABlock2D(64, 64, 64, 8, 64).to(DEV).eval()
    g = torch.Generator().manual_seed(43)
    results = []
    with torch.no_grad(), ops.precision(prec):
        mx = ops.Act(torch.randn(B * H * 64, generator=g).to(DEV), B, H, 1, 64)
        my = ops.Act(torch.randn(B * W * 64, generator=g).to(DEV), B, W, 1, 64)
        kk = blk._axis_kernels(mx, my)
        assert kk is not None
        for pooled, red, lrk, K in ((mx, blk.to_x[0], blk.low_rank_kernel_x, kk[0]), (my, blk.to_y[1], blk.low_rank_kernel_y, kk[1])):
            n = pooled.H
            x = pooled.t.view(B, n, 64).double().cpu()
            w = lambda t: t.detach().double().cpu()
            t1 = x @ w(blk.to_in[0].weight).view(64, 64).t() @ w(red.to_in.weight).t()
            t1 = F.layer_norm(t1, (64,), w(red.out_ffn[0].weight), w(red.out_ffn[0].bias), red.out_ffn[0].eps)
            z = F.gelu(t1 @ w(red.out_ffn[1].weight).t()) @ w(red.out_ffn[3].weight).t() + w(red.out_ffn[3].bias)
            qk = z @ w(lrk.to_qk.weight).t()
            q, k = [v.view(B, n, 8, 128).transpose(1, 2) for v in qk.split(1024, dim=-1)]
            cos_t, sin_t = [v.double().cpu() for v in lrk._tables(n, DEV)]
            rot = lambda v: torch.cat([v[..., :64] * cos_t - v[..., 64:] * sin_t, v[..., 64:] * cos_t + v[..., :64] * sin_t], -1)
            ref = rot(q) @ rot(k).transpose(-1, -2)
            assert tuple(K.shape) == (B, 8, n, n)
            results.append(relerr(K.cpu(), ref))
    print(f"[fa_axis {H}x{W} {prec}] K_x {results[0]:.2e} K_y {results[1]:.2e}")
    assert max(results) < (6e-3 if prec == "bf16" else 8e-4)


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("B,dil", [(4, 1), (6, 2), (601, 1), (1187, 2), (5, 1)])
def test_conv_latent_engine(B, dil, prec):
    """latent-grid engine (resident halos of 4 samples, interleaved two-sample MMA tiles, streamed filter) vs an fp64 conv
    of the same 16-bit-rounded operands, full epilogue (bias, GELU, residual); batch sizes that are not multiples of the
    4-sample super tile and larger than one wave of CTAs"""
    ops = ops_mod()
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    g = torch.Generator().manual_seed(100 + B + dil)
    x = torch.randn(B, 128, 8, 8, generator=g)
    w = torch.randn(128, 128, 3, 3, generator=g) / 34.0
    b = torch.randn(128, generator=g) * 0.1
    res = torch.randn(B, 128, 8, 8, generator=g)
    rd = lambda t: t.to(dt).float()
    ref = ref_conv(rd(x), rd(w), b, 1, dil, (dil,) * 4, (1, 1), None)
    ref = F.gelu(ref) + rd(res).double()
    h = Holder(w, b)
    with ops.precision(prec):
        y = ops.conv2d(act_from(x, dt), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4, pad_mode=(1, 1),
                       act=ops.ACT_GELU, residual=act_from(res, dt), engine=ops.ENGINE_LATENT, out_dtype=torch.float32)
        y16 = ops.conv2d(act_from(x, dt), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4, pad_mode=(1, 1),
                         act=ops.ACT_GELU, residual=act_from(res, dt))  # engine chosen by the dispatcher
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 5e-6
    assert y16.t.dtype == dt
    assert torch.equal(act_to_nchw(y16), act_to_nchw(y).to(dt).float())  # same values, rounded once


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("B,H,W", [(6, 8, 8), (3, 7, 15), (5, 12, 24), (300, 8, 8)])
def test_ffn_fused_vs_unfused_and_fp64(B, H, W, prec):
    """propagator FFN in one tcgen05 kernel (GroupNorm apply + 1x1 + GELU + 1x1 + residual) vs the three-launch path and vs
    an fp64 evaluation of the reference block (train_stage2_ns2d.py:44-53); rows straddle samples when H*W does not divide 128"""
    ops = ops_mod()
    from modules.propagator import DilatedResidualBlock, ffn_fwd
    torch.manual_seed(9)
    blk = DilatedResidualBlock(128, dilation=2).to(DEV).eval()
    with torch.no_grad():
        blk.ffn[0].weight.uniform_(0.5, 1.5)
        blk.ffn[0].bias.normal_(0, 0.2)
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    g = torch.Generator().manual_seed(50 + B)
    x = torch.randn(B, 128, H, W, generator=g) * 1.5 + 0.3
    xr = x.to(dt).double()
    gn, f1, _, f2 = blk.ffn
    w = lambda t: t.detach().double().cpu()
    ref = xr + F.conv2d(F.gelu(F.conv2d(F.group_norm(xr, 1, w(gn.weight), w(gn.bias), gn.eps), w(f1.weight))), w(f2.weight))
    a = act_from(x, dt)
    with torch.no_grad(), ops.precision(prec):
        assert ops.ffn_fused_supported(a, ops.PackedFilter.of(f1.weight, None), ops.PackedFilter.of(f2.weight, None))
        fused = act_to_nchw(ffn_fwd(blk.ffn, a))
        orig = ops.ffn_fused_supported
        ops.ffn_fused_supported = lambda *aa, **kk: False
        try:
            unfused = act_to_nchw(ffn_fwd(blk.ffn, a))
        finally:
            ops.ffn_fused_supported = orig
    e_f, e_u = relerr(fused, ref), relerr(unfused, ref)
    print(f"[ffn {B}x{H}x{W} {prec}] fused vs fp64 {e_f:.2e}, unfused {e_u:.2e}")
    tol = 8e-3 if prec == "bf16" else 1e-3
    assert e_f < tol and e_u < tol


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("B,H,W,use_pe", [(5, 8, 8, True), (3, 7, 15, True), (4, 4, 8, False), (301, 8, 8, True), (2, 8, 16, False)])
def test_sablock_fused_vs_unfused_and_fp64(B, H, W, use_pe, prec):
    """whole SABlock in one kernel vs the four-launch path and vs an fp64 evaluation of the reference block
    (modules/basics.py:384-404); one and two samples per CTA, token counts that are not multiples of 16"""
    ops = ops_mod()
    from modules.basics import SABlock
    torch.manual_seed(11)
    blk = SABlock(128, 8, 64, use_pe=use_pe, block_size=128).to(DEV).eval()
    with torch.no_grad():
        for lin in (blk.to_q, blk.to_k, blk.to_v, blk.proj_out):
            lin.weight.normal_(0, 0.08)
        blk.to_v.bias.normal_(0, 0.1)
        blk.proj_out.bias.normal_(0, 0.1)
        blk.ln.weight.uniform_(0.5, 1.5)
        blk.ln.bias.normal_(0, 0.2)
    dt = torch.bfloat16 if prec == "bf16" else torch.float16
    g = torch.Generator().manual_seed(60 + B)
    x = torch.randn(B, 128, H, W, generator=g)
    n = H * W
    w = lambda t: t.detach().double().cpu()
    xt = x.to(dt).double().flatten(2).transpose(1, 2)  # [B, n, 128]
    tn = F.layer_norm(xt, (128,), w(blk.ln.weight), w(blk.ln.bias), blk.ln.eps)
    if use_pe:
        tn = tn + w(blk.pe)[:, :n]
    q, k, v = tn @ w(blk.to_q.weight).t(), tn @ w(blk.to_k.weight).t(), tn @ w(blk.to_v.weight).t() + w(blk.to_v.bias)
    sp = lambda t: t.view(B, n, 8, 64).transpose(1, 2)
    att = F.softmax(sp(q) @ sp(k).transpose(-1, -2) * 64 ** -0.5, dim=-1) @ sp(v)
    ref = (att.transpose(1, 2).reshape(B, n, 512) @ w(blk.proj_out.weight).t() + w(blk.proj_out.bias) + xt)
    ref = ref.transpose(1, 2).reshape(B, 128, H, W)
    a = act_from(x, dt)
    with torch.no_grad(), ops.precision(prec):
        assert ops.sablock_fused_supported(a, 8, 64)
        fused = act_to_nchw(blk._fwd(a))
        orig = ops.sablock_fused_supported
        ops.sablock_fused_supported = lambda *aa, **kk: False
        try:
            unfused = act_to_nchw(blk._fwd(a))
        finally:
            ops.sablock_fused_supported = orig
    e_f, e_u = relerr(fused, ref), relerr(unfused, ref)
    print(f"[sablock {B}x{H}x{W} pe={use_pe} {prec}] fused vs fp64 {e_f:.2e}, unfused {e_u:.2e}")
    tol = 6e-3 if prec == "bf16" else 8e-4
    assert e_f < tol and e_u < tol


@pytest.mark.parametrize("cin,cout,layout,dtype", [(1, 64, "nchw", torch.float32), (3, 64, "nchw", torch.bfloat16),
                                                   (4, 64, "nchw", torch.bfloat16), (16, 128, "nhwc", torch.bfloat16),
                                                   (16, 128, "nhwc", torch.float32), (16, 16, "nhwc", torch.float32)])
def test_channel_lift_v2(cin, cout, layout, dtype):
    """1x1 channel lift (filter in registers, inputs staged once per CTA): the encoder's NCHW fp32 input lift and the
    propagator's / decoder's lift of the fp32 NHWC latent, with bias, Swish and a pixel count that is not a multiple of the
    128-pixel CTA tile; 16-bit and fp32 outputs"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(7 + cin + cout)
    B, H, W = 5, 9, 13
    x = torch.randn(B, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 1, 1, generator=g) / max(1.0, cin ** 0.5)
    b = torch.randn(cout, generator=g) * 0.2
    ref = F.silu(F.conv2d(x.double(), w.double(), b.double()))
    h = Holder(w, b)
    xa = ops.Act.from_nchw(x.to(DEV)) if layout == "nchw" else act_from(x, torch.float32)
    with ops.precision("fp32" if dtype == torch.float32 else "bf16"):
        y = ops.conv2d(xa, ops.PackedFilter.of(h.weight, h.bias), act=ops.ACT_SILU, out_dtype=dtype)
    torch.cuda.synchronize()
    assert y.t.dtype == dtype
    assert relerr(act_to_nchw(y), ref) < (3e-6 if dtype == torch.float32 else 4e-3)  # bf16: output rounding only


def test_pointwise_projection_step_grouped_output():
    """lns_pointwise_proj_steps: the samples are `steps` groups of `B` trajectories (step-major); sample (t, b) must land in
    slot [b][t] of the [B, K, C, H, W] result -- and nowhere else"""
    ops = ops_mod()
    g = torch.Generator().manual_seed(29)
    Bt, S, K, C, H, W, Co = 3, 2, 5, 64, 12, 20, 2
    x = torch.randn(S * Bt, C, H, W, generator=g)
    w = torch.randn(Co, C, 1, 1, generator=g) / 8
    b = torch.randn(Co, generator=g)
    sc, sh = torch.rand(S * Bt, C, generator=g) + 0.5, torch.randn(S * Bt, C, generator=g) * 0.2
    xin = F.silu(x.bfloat16().double() * sc[:, :, None, None].double() + sh[:, :, None, None].double())
    ref = F.conv2d(xin, w.double(), b.double()).view(S, Bt, Co, H, W)
    h = Holder(w, b)
    out = torch.zeros(Bt, K, Co, H, W, device=DEV)
    t0, chw = 2, Co * H * W  # the group covers steps t0, t0 + 1
    dst = ops.Act(out.view(-1)[t0 * chw:], S * Bt, H, W, Co, bstride=K * chw, layout=ops.NCHW, group=Bt, gstride=chw)
    with ops.precision("bf16"):
        ops.conv2d(act_from(x, torch.bfloat16), ops.PackedFilter.of(h.weight, h.bias),
                   pro=(sc.to(DEV).reshape(-1), sh.to(DEV).reshape(-1), ops.ACT_SILU), out=dst, out_layout=ops.NCHW)
    torch.cuda.synchronize()
    got = out.cpu()
    for t in range(S):
        assert relerr(got[:, t0 + t], ref[t]) < 3e-4
    assert float(got[:, :t0].abs().max()) == 0.0 and float(got[:, t0 + S:].abs().max()) == 0.0
    with pytest.raises(ops.LnsError):  # any other conv refuses a step-grouped output
        w3 = Holder(torch.randn(8, C, 1, 1, generator=g), torch.zeros(8))
        bad = ops.Act(torch.zeros(S * Bt * 8 * H * W, device=DEV), S * Bt, H, W, 8, layout=ops.NCHW, group=Bt, gstride=1)
        ops.conv2d(act_from(x, torch.bfloat16), ops.PackedFilter.of(w3.weight, w3.bias), out=bad, out_layout=ops.NCHW)


# ---- split operands (LNS_W_UMMA_F16X2): the building blocks of the 'fp16s' precision mode ------------------------------------
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_umma_split_x3(case):
    """gather engine, fp32 activations split into f16 hi + lo in the producer, split filter, three MMAs per K step:
    fp32-class result vs an fp64 conv of the UNROUNDED operands"""
    ops = ops_mod()
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    x, w, b = _conv_case(case, seed=5)
    ref = ref_conv(x, w, b, stride, dil, pad, modes, virt)
    h = Holder(w, b)
    res = torch.randn(ref.shape, generator=torch.Generator().manual_seed(6))
    with ops.precision("fp16"):
        y = ops.conv2d(act_from(x, torch.float32), ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil, pad=pad,
                       pad_mode=modes, virt=virt, engine=ops.ENGINE_UMMA, split=True, out_dtype=torch.float32,
                       act=ops.ACT_GELU, residual=act_from(res, torch.float32))
    torch.cuda.synchronize()
    assert y.t.dtype == torch.float32
    assert relerr(act_to_nchw(y), F.gelu(ref) + res.double()) < 3e-6


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_umma_split_w2(case):
    """gather engine, f16 activations, split filter, two MMAs per K step: no filter rounding -- vs an fp64 conv of the
    f16-rounded activation and the UNROUNDED filter"""
    ops = ops_mod()
    B, Cin, Cout, H, W, k, stride, dil, pad, modes, virt = case
    x, w, b = _conv_case(case, seed=7)
    ref = ref_conv(x.half().float(), w, b, stride, dil, pad, modes, virt)
    h = Holder(w, b)
    with ops.precision("fp16"):
        y = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), stride=stride, dil=dil, pad=pad,
                       pad_mode=modes, virt=virt, engine=ops.ENGINE_UMMA, split=True, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 3e-6


@pytest.mark.parametrize("case", [c for c in HALO_CASES if c[1] == 64 and c[4] == 1])
def test_conv_halo_split_w2(case):
    """halo engine with BOTH filter planes resident (Cout = 64) and two MMAs per (tap, k)"""
    ops = ops_mod()
    B, Cout, H, W, dil, modes, virt = case
    g = torch.Generator().manual_seed(33)
    x = torch.randn(B, 64, H, W, generator=g)
    w = torch.randn(Cout, 64, 3, 3, generator=g) / 24.0
    b = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = virt if virt is not None else (H, W)
    res = torch.randn(B, Cout, Ho, Wo, generator=g)
    ref = ref_conv(x.half().float(), w, b, 1, dil, (dil,) * 4, modes, virt)
    ref = F.gelu(ref) + res.half().double()
    h = Holder(w, b)
    with ops.precision("fp16"):
        y = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4,
                       pad_mode=modes, virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.float16),
                       engine=ops.ENGINE_HALO, split=True, out_dtype=torch.float32)
        y16 = ops.conv2d(act_from(x, torch.float16), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4,
                         pad_mode=modes, virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.float16),
                         engine=ops.ENGINE_HALO, split=True)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 5e-6
    assert torch.equal(act_to_nchw(y16), act_to_nchw(y).half().float())


# ---- block-halo engine for the coarse levels (LNS_ENGINE_COARSE, csrc/conv_coarse.cu) ---------------------------------------
COARSE_CASES = [
    # B, Cin, Cout, H, W, dil, modes(h,w), virt
    (5, 128, 128, 8, 8, 1, (1, 1), None),        # NS2d latent grid, odd block count (ragged last super tile)
    (4, 128, 128, 8, 8, 2, (1, 1), None),        # dilation 2, circular
    (3, 128, 128, 7, 15, 2, (0, 0), None),       # two-phase latent grid: ragged blocks, zeros
    (2, 128, 128, 12, 24, 3, (0, 1), None),      # shallow-water latent grid: half periodic, dilation 3
    (3, 64, 64, 16, 16, 1, (1, 1), None),
    (2, 128, 64, 16, 16, 1, (1, 1), None),
    (2, 64, 128, 24, 48, 1, (0, 1), None),
    (3, 128, 128, 8, 8, 1, (1, 1), (16, 16)),    # nearest x2 folded into the halo fill
    (2, 64, 64, 14, 30, 1, (0, 0), (28, 60)),
    (2, 64, 64, 28, 60, 1, (0, 0), (61, 121)),   # general nearest to an odd size
    (300, 128, 128, 8, 8, 1, (1, 1), None),      # more super tiles than SMs: persistent loop, ring wrap-around
]


def _coarse_case(case, seed):
    B, Cin, Cout, H, W, dil, modes, virt = case
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g) * 0.1
    Ho, Wo = virt if virt is not None else (H, W)
    res = torch.randn(B, Cout, Ho, Wo, generator=g)
    return x, w, b, res


@pytest.mark.parametrize("case", COARSE_CASES)
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
def test_conv_coarse_16bit(case, prec):
    """16-bit activations (one halo plane per slab) vs an fp64 conv of the same rounded operands; full epilogue"""
    ops = ops_mod()
    B, Cin, Cout, H, W, dil, modes, virt = case
    dt = torch.float16 if prec == "fp16" else torch.bfloat16
    x, w, b, res = _coarse_case(case, 41)
    ref = ref_conv(x.to(dt).float(), w.to(dt).float(), b, 1, dil, (dil,) * 4, modes, virt)
    ref = F.gelu(ref) + res.to(dt).double()
    h = Holder(w, b)
    with ops.precision(prec):
        y = ops.conv2d(act_from(x, dt), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4, pad_mode=modes, virt=virt,
                       act=ops.ACT_GELU, residual=act_from(res, dt), engine=ops.ENGINE_COARSE, out_dtype=torch.float32)
        y16 = ops.conv2d(act_from(x, dt), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4, pad_mode=modes, virt=virt,
                         act=ops.ACT_GELU, residual=act_from(res, dt), engine=ops.ENGINE_COARSE)
    torch.cuda.synchronize()
    assert tuple(act_to_nchw(y).shape) == tuple(ref.shape)
    assert relerr(act_to_nchw(y), ref) < 5e-6
    assert y16.t.dtype == dt
    assert torch.equal(act_to_nchw(y16), act_to_nchw(y).to(dt).float())


@pytest.mark.parametrize("case", COARSE_CASES)
@pytest.mark.parametrize("wsplit", [False, True])
def test_conv_coarse_fp32_split(case, wsplit):
    """fp32 activations split into f16 hi + lo halo planes in the producer.  Plain filter: no activation rounding (vs fp64 conv of
    the exact activation and the f16-rounded filter); split filter: fp32-class result vs the exact conv"""
    ops = ops_mod()
    B, Cin, Cout, H, W, dil, modes, virt = case
    if not ops._coarse_fits(Cin, Cout, dil, True, wsplit):
        pytest.skip("two fp32-split halo planes per slab do not fit in shared memory at this dilation (gather engine instead)")
    x, w, b, res = _coarse_case(case, 43)
    ref = ref_conv(x, w if wsplit else w.half().float(), b, 1, dil, (dil,) * 4, modes, virt)
    ref = F.gelu(ref) + res.double()
    h = Holder(w, b)
    with ops.precision("fp16"):
        y = ops.conv2d(act_from(x, torch.float32), ops.PackedFilter.of(h.weight, h.bias), dil=dil, pad=(dil,) * 4, pad_mode=modes,
                       virt=virt, act=ops.ACT_GELU, residual=act_from(res, torch.float32), engine=ops.ENGINE_COARSE, split=wsplit,
                       out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert relerr(act_to_nchw(y), ref) < 3e-6


# ---- FABlock2D with every contraction on tcgen05 (csrc/fablock_tc.cu) -----------------------------------------------------------
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("n,B", [(16, 5), (32, 3), (32, 300)])
def test_fablock_tc_vs_full_and_oracle(n, B, prec):
    """tcgen05 whole-block kernel (block-diagonal axial GEMMs, MN-major pixel rows, 2-CTA cluster at 32x32) vs the fp64 oracle
    of the reference block and vs the mma.sync whole-block kernel"""
    ops = ops_mod()
    import lns_oracle as O
    from modules.factorized_attention import FABlock2D
    dt = torch.float16 if prec == "fp16" else torch.bfloat16
    torch.manual_seed(3)
    blk = FABlock2D(64, 64, 64, 8, 64).to(DEV).eval()
    g = torch.Generator().manual_seed(45)
    x = torch.randn(B, 64, n, n, generator=g)
    a = act_from(x, dt)
    saved = ops._state.fablock_tc
    ops._state.fablock_tc = True
    try:
        assert ops.fablock_tc_supported(a, 64, 64)
        with torch.no_grad(), ops.precision(prec):
            tc = act_to_nchw(blk._fwd(a))
            ops._state.fablock_tc = False
            full = act_to_nchw(blk._fwd(a))
    finally:
        ops._state.fablock_tc = saved
    nb = min(B, 4)
    sd = {k: v.cpu().double() for k, v in blk.state_dict().items()}
    ref = O.fa_block(x[:nb].double(), O.SD(sd))
    e_tc, e_full = relerr(tc[:nb], ref), relerr(full[:nb], ref)
    print(f"[FABlock2D {n}x{n} {prec} B={B}] tcgen05 kernel vs fp64 oracle {e_tc:.2e}, mma.sync kernel {e_full:.2e}, tc vs full "
          f"{relerr(tc, full):.2e}")
    tol = 4e-3 if prec == "fp16" else 3e-2
    assert e_tc < tol and relerr(tc, full) < tol
    # every sample, not only the first ones: per-sample error against the other kernel
    per = ((tc - full).flatten(1).norm(dim=1) / full.flatten(1).norm(dim=1)).max().item()
    assert per < tol


# ---- FABlock2D whole-block kernel on pre-staged operands with a producer warp (csrc/fablock_full.cu, fablock_full2_kernel) --------
@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("H,W,B", [(16, 16, 5), (32, 32, 3), (32, 32, 300), (16, 16, 700), (16, 32, 4), (32, 16, 4)])
def test_fablock_staged_vs_full_and_oracle(H, W, B, prec):
    """producer-warp whole-block kernel (normalised input staged by the pre-pass, bulk copies, tcgen05 issue off the compute warps)
    vs the fp64 oracle of the reference block (modules/factorized_attention.py:144-159) and vs the in-kernel-staging kernel"""
    ops = ops_mod()
    import lns_oracle as O
    from modules.factorized_attention import FABlock2D
    dt = torch.float16 if prec == "fp16" else torch.bfloat16
    torch.manual_seed(3)
    blk = FABlock2D(64, 64, 64, 8, 64).to(DEV).eval()
    with torch.no_grad():  # a non-trivial GroupNorm affine: the staged path applies it in the pre-pass
        blk.in_norm.weight.uniform_(0.5, 1.5)
        blk.in_norm.bias.uniform_(-0.3, 0.3)
    g = torch.Generator().manual_seed(46)
    x = torch.randn(B, 64, H, W, generator=g) * 1.7 + 0.4
    a = act_from(x, dt)
    saved = ops._state.fablock_staged
    ops._state.fablock_staged = True
    try:
        assert ops.fablock_full_staged_supported(a, 64, 64)
        with torch.no_grad(), ops.precision(prec):
            st = act_to_nchw(blk._fwd(a))
            st2 = act_to_nchw(blk._fwd(a))
            ops._state.fablock_staged = False
            full = act_to_nchw(blk._fwd(a))
    finally:
        ops._state.fablock_staged = saved
    assert torch.equal(st, st2)  # deterministic (fixed-order reductions, no atomics)
    nb = min(B, 4)
    sd = {k: v.cpu().double() for k, v in blk.state_dict().items()}
    ref = O.fa_block(x[:nb].double(), O.SD(sd))
    e_st, e_full = relerr(st[:nb], ref), relerr(full[:nb], ref)
    print(f"[FABlock2D {H}x{W} {prec} B={B}] staged kernel vs fp64 oracle {e_st:.2e}, in-kernel staging {e_full:.2e}, staged vs full "
          f"{relerr(st, full):.2e}")
    tol = 4e-3 if prec == "fp16" else 3e-2
    assert e_st < tol and relerr(st, full) < tol
    per = ((st - full).flatten(1).norm(dim=1) / full.flatten(1).norm(dim=1)).max().item()
    assert per < tol


@pytest.mark.parametrize("prec", ["fp16", "bf16"])
@pytest.mark.parametrize("H,W,B", [(16, 16, 7), (32, 32, 5), (16, 32, 3), (32, 16, 3), (32, 32, 300)])
def test_fablock_prepass_staged_v3_vs_v2_and_torch(H, W, B, prec):
    """bulk-copy-fed pre-pass (sample resident in shared memory) vs the LSU pre-pass and vs torch: GroupNorm(1, 64) affine,
    both pooled tensors (modules/factorized_attention.py:86-94,113), and the staged copy = GN(u) in the whole-block kernel's
    shared-memory order (row s = pixel s ^ ((s >> log2 W) & 7), 16-byte chunk ch at ch ^ (s & 7))"""
    import os
    ops = ops_mod()
    dt = torch.float16 if prec == "fp16" else torch.bfloat16
    g = torch.Generator().manual_seed(47)
    x = torch.randn(B, 64, H, W, generator=g) * 1.3 + 0.7
    gamma = (torch.rand(64, generator=g) + 0.5).to(DEV)
    beta = (torch.randn(64, generator=g) * 0.2).to(DEV)
    a = act_from(x, dt)
    with torch.no_grad(), ops.precision(prec):
        s3, t3, px3, py3, st3 = ops.fablock_prepass(a, 1e-5, gamma, beta, staged=True)
        os.environ["LNS_PREPASS_V2"] = "1"
        try:
            s2, t2, px2, py2, st2 = ops.fablock_prepass(a, 1e-5, gamma, beta, staged=True)
        finally:
            del os.environ["LNS_PREPASS_V2"]
    torch.cuda.synchronize()
    xr = a.t.float().view(B, H, W, 64).double()  # the 16-bit values the kernels read
    mean = xr.mean(dim=(1, 2, 3), keepdim=True)
    var = xr.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    xn = (xr - mean) / torch.sqrt(var + 1e-5) * gamma.double() + beta.double()
    assert relerr(px3.t.view(B, H, 64), xn.mean(dim=2)) < 2e-6 and relerr(py3.t.view(B, W, 64), xn.mean(dim=1)) < 2e-6
    for v3, v2 in ((s3, s2), (t3, t2), (px3.t, px2.t), (py3.t, py2.t)):
        assert relerr(v3, v2) < 1e-6
    # staged copy: undo the permutation + swizzle and compare with GN(u)
    lg = 5 if W == 32 else 4
    sl = torch.arange(H * W, device=DEV)
    src = sl ^ ((sl >> lg) & 7)
    chunk = torch.arange(8, device=DEV)
    pos = (chunk[None, :] ^ (sl[:, None] & 7))  # [HW][8]: where chunk ch of row sl sits
    st = st3.view(B, H * W, 8, 8).float()
    un = torch.empty_like(st)
    un[:, src[:, None].expand(-1, 8), chunk[None, :].expand(H * W, -1)] = st[:, sl[:, None].expand(-1, 8), pos]
    tol = 1.2e-3 if prec == "fp16" else 9e-3  # one rounding to 16 bits
    assert relerr(un.view(B, H, W, 64), xn) < tol
    assert (st3.float() - st2.float()).abs().max() <= 2 ** (-8 if prec == "fp16" else -5)  # at most one 16-bit ulp apart
