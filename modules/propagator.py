"""Latent propagators.

1. ``SimpleCNN`` / ``DilatedResidualBlock`` (+ the conditional variants) -- the propagator that the reference actually
   trains and rolls out.  The reference defines them inside each stage-2 script (train_stage2_ns2d.py:25-87,
   train_stage2_SW.py:25-88, train_stage2_twophase.py:25-88, train_stage2_twophase_conditional.py:25-121); they are
   provided here as importable classes with identical parameter names, and ``lns_b200.rollout`` also recognises the
   scripts' own instances structurally.
2. ``SimpleMLP`` / ``SimpleResNet`` / ``ConditionalResNet`` -- the classes of the reference's ``modules/propagator.py``.
   Only ``SimpleMLP`` is constructible there (the two ResNets call ``ResidualBlock`` without the required
   ``num_dimensions``, modules/propagator.py:22-24,87); that defect is mirrored, not fixed.
"""
import torch
import torch.nn as nn

from lns_b200 import ops

from ._base import LnsModule, LnsError, conv_layer, filt_of, norm_affine, lazy_norm
from .basics import GroupNorm, ResidualBlock, Swish
from .cond_utils import zero_module, fourier_embedding


def _make_conv(dim_in, dim_out, k, padding, dilation, padding_mode, periodic_direction, bias=True):
    if periodic_direction is not None:
        from .autoencoder2d_half_periodic import HalfPeriodicConv2d
        return HalfPeriodicConv2d(dim_in, dim_out, kernel_size=k, stride=1, padding=padding, dilation=dilation,
                                  bias=bias, periodic_direction=periodic_direction)
    return nn.Conv2d(dim_in, dim_out, kernel_size=k, stride=1, padding=padding, dilation=dilation,
                     padding_mode=padding_mode, bias=bias)


class DilatedResidualBlock(LnsModule):
    """x += Conv3(GELU(Conv3_dil(GELU(Conv3(GN1(x))))));  x += Conv1(GELU(Conv1(GN1(x))))
    (reference: train_stage2_ns2d.py:25-53; half-periodic convs in train_stage2_SW.py:25-53)."""

    def __init__(self, dim, dilation=1, padding_mode="circular", periodic_direction=None):
        super().__init__()
        self.dim, self.dilation, self.padding_mode = dim, dilation, padding_mode
        mk = lambda pad, dil: _make_conv(dim, dim, 3, pad, dil, padding_mode, periodic_direction)  # noqa: E731
        self.conv = nn.Sequential(nn.GroupNorm(1, dim), mk(1, 1), nn.GELU(), mk(dilation, dilation), nn.GELU(), mk(1, 1))
        self.ffn = nn.Sequential(nn.GroupNorm(1, dim), nn.Conv2d(dim, dim, 1, 1, 0, bias=False), nn.GELU(),
                                 nn.Conv2d(dim, dim, 1, 1, 0, bias=False))

    def _fwd(self, x):
        return dilated_block_fwd(self, x)


def dilated_block_fwd(blk, x):
    """Works on this module's class and on the reference scripts' structurally identical one."""
    gn, c1, _, c2, _, c3 = blk.conv
    h = conv_layer(x, c1, pro=lazy_norm(x, gn), act=ops.ACT_GELU)
    h = conv_layer(h, c2, act=ops.ACT_GELU)
    x = conv_layer(h, c3, residual=x)
    return ffn_fwd(blk.ffn, x)


def ffn_fwd(ffn, x, prescale=None):
    """x + Conv1(GELU(Conv1(GN1(x * prescale)))): one tcgen05 kernel (statistics kernel + lns_ffn_fused) on the 16-bit paths
    when the block is 128 -> 128 -> 128 with bias-free 1x1 convs, else three launches."""
    gn2, f1, _, f2 = ffn
    if (isinstance(f1, nn.Conv2d) and isinstance(f2, nn.Conv2d) and f1.kernel_size == (1, 1) and f2.kernel_size == (1, 1)
            and f1.stride == (1, 1) and f2.stride == (1, 1) and ops.ffn_fused_supported(x, filt_of(f1), filt_of(f2))):
        s, t = norm_affine(x, gn2, prescale=prescale)
        return ops.ffn_fused(x, s, t, filt_of(f1), filt_of(f2))
    h = conv_layer(x, f1, pro=lazy_norm(x, gn2, prescale=prescale), act=ops.ACT_GELU)
    return conv_layer(h, f2, residual=x)


def simple_cnn_fwd(net, z, out=None, out_dtype=None):
    """in_proj 1x1 -> blocks -> GroupNorm(32) -> 1x1 (reference: train_stage2_ns2d.py:82-87)."""
    z = conv_layer(z, net.in_proj)
    for blk in net.net:
        z = dilated_block_fwd(blk, z)
    return conv_layer(z, net.out_proj[1], pro=lazy_norm(z, net.out_proj[0]), out=out,
                      out_dtype=out_dtype if out is None else None)


class SimpleCNN(LnsModule):
    """The unconditional latent propagator (reference: train_stage2_ns2d.py:56-87).  ``padding_mode`` /
    ``periodic_direction`` select the NS2d (circular), two-phase (zeros) and shallow-water (half-periodic) variants,
    which differ only in that argument in the reference scripts."""

    def __init__(self, latent_dim, prop_n_block, prop_n_embd, dilation=2, padding_mode="circular",
                 periodic_direction=None):
        super().__init__()
        self.latent_dim, self.prop_n_block, self.prop_n_embd = latent_dim, prop_n_block, prop_n_embd
        self.in_proj = nn.Conv2d(latent_dim, prop_n_embd, 1, 1, 0)
        self.net = nn.Sequential(*[DilatedResidualBlock(prop_n_embd, dilation=dilation, padding_mode=padding_mode,
                                                        periodic_direction=periodic_direction)
                                   for _ in range(prop_n_block)])
        self.out_proj = nn.Sequential(GroupNorm(prop_n_embd), nn.Conv2d(prop_n_embd, latent_dim, 1, 1, 0))

    def _fwd(self, z):
        return simple_cnn_fwd(self, z, out_dtype=torch.float32)

    def forward(self, z):
        return self._fwd(ops.nchw_to_act(z, torch.float32)).to_nchw()


class CondDilatedResidualBlock(LnsModule):
    """Conditional block (reference: train_stage2_twophase_conditional.py:25-75):
        h = Conv3_dil(GELU(Conv3(GN1(x)))) + Linear(emb);  x = x + Conv3_zero(GELU(GN1(h)))
        x = x + FFN(x * (1 + cond_conv2(emb_out)))   with cond_conv2 = GN1 -> 1x1 -> GELU -> 1x1_zero on [B,C,1,1]"""

    def __init__(self, dim, cond_emb_dim, dilation=1, padding_mode="circular"):
        super().__init__()
        self.dim, self.dilation, self.padding_mode = dim, dilation, padding_mode
        self.cond_emb = nn.Linear(cond_emb_dim, dim)
        self.conv1 = nn.Sequential(
            nn.GroupNorm(1, dim),
            nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, padding_mode=padding_mode),
            nn.GELU(),
            nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=dilation, dilation=dilation, padding_mode=padding_mode))
        self.cond_conv1 = nn.Sequential(
            nn.GroupNorm(1, dim), nn.GELU(),
            zero_module(nn.Conv2d(dim, dim, kernel_size=(3, 3), padding=(1, 1), padding_mode=padding_mode)))
        self.cond_conv2 = nn.Sequential(nn.GroupNorm(1, dim), nn.Conv2d(dim, dim, 1), nn.GELU(),
                                        zero_module(nn.Conv2d(dim, dim, 1)))
        self.ffn = nn.Sequential(nn.GroupNorm(1, dim), nn.Conv2d(dim, dim, 1, 1, 0, bias=False), nn.GELU(),
                                 nn.Conv2d(dim, dim, 1, 1, 0, bias=False))

    def _fwd(self, x, cond):
        return cond_block_fwd(self, x, cond_block_prepare(self, cond))


def cond_block_prepare(blk, cond_rows):
    """Step-invariant conditioning vectors of one block from cond_emb [B,1,1,Ce] (fp32 Act):
    (shift [B,C] = Linear(emb), 1 + gate [B,C] with gate = cond_conv2(shift))  -- hoisted out of the K-step loop."""
    f32 = torch.float32
    shift = ops.conv2d(cond_rows, filt_of(blk.cond_emb), out_dtype=f32)          # [B,1,1,C]
    gn, c1, _, c2 = blk.cond_conv2
    s, t = norm_affine(shift, gn)                                               # GN(1,C) over the C values of a sample
    g = conv_layer(shift, c1, pro=(s, t, ops.ACT_NONE), act=ops.ACT_GELU, out_dtype=f32)
    gate = conv_layer(g, c2, out_dtype=f32)
    ones = torch.ones_like(gate.t)
    one_plus = ops.affine_act(gate, ones, ones, ops.ACT_NONE)                    # 1 + gate, [B,C]
    return shift.t, one_plus.t


def cond_block_fwd(blk, x, prepared):
    shift, one_plus_gate = prepared
    gn, c1, _, c2 = blk.conv1
    h = conv_layer(x, c1, pro=lazy_norm(x, gn), act=ops.ACT_GELU)
    h = conv_layer(h, c2, sample_bias=shift)
    gn1, _, cz = blk.cond_conv1
    x = conv_layer(h, cz, pro=lazy_norm(h, gn1, ops.ACT_GELU), residual=x)
    # GN1(x * (1 + gate)): per-channel statistics of x rescale exactly, so the gate only enters the finalize kernel
    return ffn_fwd(blk.ffn, x, prescale=one_plus_gate)


class CondSimpleCNN(LnsModule):
    """Conditional propagator (reference: train_stage2_twophase_conditional.py:78-121; there it is also called
    ``SimpleCNN``): cond_emb = MLP(fourier_embedding(param)); in_proj; blocks(z, cond_emb); GN32 -> 1x1."""

    def __init__(self, latent_dim, cond_emb_dim, prop_n_block, prop_n_embd, dilation=2):
        super().__init__()
        self.latent_dim, self.cond_emb_dim = latent_dim, cond_emb_dim
        self.prop_n_block, self.prop_n_embd = prop_n_block, prop_n_embd
        self.in_proj = nn.Conv2d(latent_dim, prop_n_embd, 1, 1, 0)
        self.cond_emb_proj = nn.Sequential(nn.Linear(cond_emb_dim, cond_emb_dim), nn.GELU(),
                                           nn.Linear(cond_emb_dim, cond_emb_dim))
        self.net = nn.ModuleList([CondDilatedResidualBlock(prop_n_embd, cond_emb_dim=cond_emb_dim, dilation=dilation,
                                                           padding_mode="zeros") for _ in range(prop_n_block)])
        self.out_proj = nn.Sequential(GroupNorm(prop_n_embd), nn.Conv2d(prop_n_embd, latent_dim, 1, 1, 0))

    def _fwd(self, z, param):
        return cond_cnn_fwd(self, z, cond_cnn_prepare(self, param), out_dtype=torch.float32)

    def forward(self, z, param):
        return self._fwd(ops.nchw_to_act(z, torch.float32), param).to_nchw()


def cond_cnn_prepare(net, param):
    """Everything that depends only on `param` (computed once per rollout, not per step)."""
    emb = ops.rows_act(fourier_embedding(param, dim=net.cond_emb_dim))
    l1, _, l2 = net.cond_emb_proj
    f32 = torch.float32
    h = ops.conv2d(emb, filt_of(l1), act=ops.ACT_GELU, out_dtype=f32)
    cond = ops.conv2d(h, filt_of(l2), out_dtype=f32)
    return [cond_block_prepare(blk, cond) for blk in net.net]


def cond_cnn_fwd(net, z, prepared, out=None, out_dtype=None):
    z = conv_layer(z, net.in_proj)
    for blk, prep in zip(net.net, prepared):
        z = cond_block_fwd(blk, z, prep)
    return conv_layer(z, net.out_proj[1], pro=lazy_norm(z, net.out_proj[0]), out=out,
                      out_dtype=out_dtype if out is None else None)


# ---- the classes of the reference's modules/propagator.py --------------------------------------------------------
class SimpleMLP(LnsModule):
    """x + MLP(flatten(x)) with flatten order (h w c) (reference: modules/propagator.py:34-51)."""

    def __init__(self, args):
        super().__init__()
        d = args.latent_dim * args.latent_resolution ** 2
        self.net = nn.Sequential(nn.Linear(d, args.propagator_dim), Swish(),
                                 nn.Linear(args.propagator_dim, args.propagator_dim), Swish(),
                                 nn.Linear(args.propagator_dim, d))

    def _fwd(self, x):
        # NHWC memory order of one sample IS the reference's '(h w c)' flattening
        f32 = torch.float32
        rows = ops.Act(x.t, x.B, 1, 1, x.H * x.W * x.C, bstride=x.bstride)
        h = ops.conv2d(rows, filt_of(self.net[0]), act=ops.ACT_SILU, out_dtype=f32)
        h = ops.conv2d(h, filt_of(self.net[2]), act=ops.ACT_SILU, out_dtype=f32)
        y = ops.conv2d(h, filt_of(self.net[4]), residual=rows, out_dtype=f32)
        return ops.Act(y.t, x.B, x.H, x.W, x.C)

    def forward(self, x):
        return self._fwd(ops.nchw_to_act(x, torch.float32)).to_nchw()


class SimpleResNet(nn.Module):
    """Unconstructible in the reference (ResidualBlock called without num_dimensions, modules/propagator.py:22-24);
    mirrored: the same TypeError is raised here."""

    def __init__(self, args):
        super().__init__()
        padding_mode = "circular" if args.is_periodic else "zeros"
        self.net = nn.Sequential(
            nn.Conv2d(args.latent_dim, args.propagator_dim, 1, 1, 0), Swish(),
            nn.Conv2d(args.propagator_dim, args.propagator_dim, 3, 1, 1, padding_mode=padding_mode),
            GroupNorm(args.propagator_dim),
            ResidualBlock(args.propagator_dim, args.propagator_dim, padding_mode=padding_mode),
            ResidualBlock(args.propagator_dim, args.propagator_dim, padding_mode=padding_mode),
            ResidualBlock(args.propagator_dim, args.propagator_dim, padding_mode=padding_mode),
            GroupNorm(args.propagator_dim), Swish(),
            nn.Conv2d(args.propagator_dim, args.latent_dim, 1, 1, 0))


class ConditionalResNet(nn.Module):
    """Unconstructible in the reference as well (modules/propagator.py:87 -- same missing argument; it additionally
    needs ``CABlock``, which has an inverted ``channel_last`` branch, modules/basics.py:526).  Not on any shipped path."""

    def __init__(self, args):
        super().__init__()
        raise TypeError("ConditionalResNet cannot be constructed in the reference either "
                        "(ResidualBlock.__init__() missing 1 required positional argument: 'num_dimensions')")
