"""Shallow-water autoencoder, periodic along one axis and zero-padded along the other -- drop-in for the reference's
``modules/autoencoder2d_half_periodic.py`` (``HalfPeriodicConv2d`` :26-52, ``UpSampleBlock2D`` :55-65,
``DownSampleBlock2d`` :68-74, ``HalfPeriodicResBlock2d`` :77-103, ``Encoder`` :106-144, ``Decoder`` :147-230,
``SimpleAutoencoder`` :233-259)."""
import math

import torch
import torch.nn as nn

from lns_b200 import ops

from ._base import LnsModule, run_layers, conv_layer, norm_affine, lazy_norm, latent_dtype, encoder_hi_px, decoder_hi_px
from .basics import GroupNorm, Swish, FourierBasicBlock, SABlock
from .factorized_attention import FABlock2D


class NormSwish(LnsModule):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm_act = nn.Sequential(GroupNorm(in_channels), Swish())

    def _affine(self, x):
        return lazy_norm(x, self.norm_act[0], ops.ACT_SILU)

    def _fwd(self, x):
        return self._affine(x).materialize()


class HalfPeriodicConv2d(nn.Conv2d, LnsModule):
    """nn.Conv2d whose own padding is 0; `padding` is applied by hand, circular along `periodic_direction` and zeros
    along the other axis (reference :26-52).  Here both paddings are part of the conv kernel's index map."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 periodic_direction="x"):
        nn.Conv2d.__init__(self, in_channels, out_channels, kernel_size, stride, 0, dilation, groups, bias)
        if groups != 1:
            raise NotImplementedError("grouped convolutions are not on the rollout path")
        if periodic_direction not in ("x", "y"):
            raise ValueError("periodic_direction must be x or y")
        self.padding_ = padding
        self.periodic_direction = periodic_direction

    def _fwd(self, x, **kw):
        return conv_layer(x, self, **kw)

    def forward(self, x):
        return LnsModule.forward(self, x)


class UpSampleBlock2D(LnsModule):
    def __init__(self, channels, periodic_direction="x"):
        super().__init__()
        self.conv_layer = HalfPeriodicConv2d(channels, channels, 3, 1, 1, periodic_direction=periodic_direction)

    def _fwd(self, x):
        return conv_layer(x, self.conv_layer, virt=(2 * x.H, 2 * x.W))


class DownSampleBlock2d(LnsModule):
    def __init__(self, channels, periodic_direction="x"):
        super().__init__()
        self.conv_layer = HalfPeriodicConv2d(channels, channels, 3, 2, 1, periodic_direction=periodic_direction)

    def _fwd(self, x):
        return conv_layer(x, self.conv_layer)


class HalfPeriodicResBlock2d(LnsModule):
    """conv2(NormSwish(conv1(NormSwish(x)))) + [channel_up](x)  (reference :77-103)."""

    def __init__(self, in_channels, out_channels, periodic_direction="x"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm_act1 = NormSwish(in_channels)
        self.norm_act2 = NormSwish(out_channels)
        self.periodic_direction = periodic_direction
        self.conv1 = HalfPeriodicConv2d(in_channels, out_channels, 3, 1, 1, periodic_direction=periodic_direction)
        self.conv2 = HalfPeriodicConv2d(out_channels, out_channels, 3, 1, 1, periodic_direction=periodic_direction)
        if in_channels != out_channels:
            self.channel_up = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def _fwd(self, x):
        skip = conv_layer(x, self.channel_up) if hasattr(self, "channel_up") else x
        h = conv_layer(x, self.conv1, pro=self.norm_act1._affine(x))
        return conv_layer(h, self.conv2, pro=self.norm_act2._affine(h), residual=skip)


class Encoder(LnsModule):
    def __init__(self, args):
        super().__init__()
        ch = args.encoder_channels
        height = args.resolutions[0]
        assert (len(ch) - 2) == int(math.log2(height // args.latent_resolution))
        pd = args.periodic_direction
        layers = [nn.Conv2d(args.in_channels, ch[0], 1, 1, 0), Swish(),
                  HalfPeriodicResBlock2d(ch[0], ch[0], periodic_direction=pd)]
        for i in range(len(ch) - 1):
            cin, cout = ch[i], ch[i + 1]
            for _ in range(args.encoder_res_blocks):
                layers.append(HalfPeriodicResBlock2d(cin, cout, periodic_direction=pd))
                cin = cout
            if i != len(ch) - 2:
                layers.append(DownSampleBlock2d(ch[i + 1], periodic_direction=pd))
                height //= 2
        layers += [HalfPeriodicResBlock2d(ch[-1], ch[-1], periodic_direction=pd), GroupNorm(ch[-1]), Swish(),
                   nn.Conv2d(ch[-1], args.latent_dim, 1, 1, 0)]
        self.model = nn.Sequential(*layers)

    def _fwd(self, x):
        return run_layers(self.model, x, final_dtype=latent_dtype(self.model[-1].out_channels))

    def forward(self, x):
        with ops.device_of(x):
            return self._fwd(ops.Act.from_nchw(x)).to_nchw()


class Decoder(LnsModule):
    def __init__(self, args):
        super().__init__()
        ch = args.decoder_channels
        pd = args.periodic_direction
        heads, dim = args.decoder_attn_heads, args.decoder_attn_dim
        cin = ch[0]
        height = args.latent_resolution
        hw_ratio = args.resolutions[1] / args.resolutions[0]
        no_coarse = bool(args.disable_coarse_attn) if args.disable_coarse_attn is not None else False
        rb = lambda a, b: HalfPeriodicResBlock2d(a, b, periodic_direction=pd)  # noqa: E731

        def attn(c, h):
            if args.use_fa:
                return FABlock2D(c, dim, dim, heads, c, use_rope=True, kernel_multiplier=2)
            return SABlock(c, heads, dim, use_pe=False, block_size=h * int(h * (hw_ratio + 0.5)))

        layers = [HalfPeriodicConv2d(args.latent_dim, cin, 3, 1, 1, periodic_direction=pd)]
        if not no_coarse:
            layers += [SABlock(cin, heads, dim, use_pe=False, block_size=height * int(height * (hw_ratio + 0.5))),
                       rb(cin, cin)]
        else:
            layers += [rb(cin, cin), rb(cin, cin)]
        for i in range(len(ch)):
            cout = ch[i]
            for _ in range(args.decoder_res_blocks):
                layers.append(rb(cin, cout))
                cin = cout
                if height in args.attn_resolutions:
                    layers.append(attn(cin, height))
            if i != 0 and i != len(ch) - 1:
                layers.append(UpSampleBlock2D(cin, periodic_direction=pd))
                height *= 2
        layers.append(nn.Upsample(size=(args.Ly, args.Lx), mode="nearest"))
        height = args.Ly
        layers.append(HalfPeriodicConv2d(cin, cin, 3, 1, 1, periodic_direction=pd))
        if args.final_smoothing:
            layers.append(FourierBasicBlock(cin, cin, modes=[16, int(16 * hw_ratio)]))
        else:
            if height in args.attn_resolutions:
                layers.append(attn(cin, height))
            layers.append(HalfPeriodicConv2d(cin, cin, 3, 1, 1, periodic_direction=pd))
        layers += [GroupNorm(cin), Swish(), nn.Conv2d(cin, args.in_channels, 1, 1, 0)]
        self.model = nn.Sequential(*layers)

    def _fwd(self, x, out=None):
        return run_layers(self.model, x, final_out=out, final_layout=ops.NCHW)

    def forward(self, x):
        with ops.device_of(x):
            return self._fwd(ops.nchw_to_act(x)).to_nchw()


class SimpleAutoencoder(LnsModule):
    def __init__(self, args):
        super().__init__()
        self.encoder = Encoder(args)
        self.decoder = Decoder(args)
        self.quant_conv = nn.Conv2d(args.latent_dim, args.latent_dim, 1)
        self.post_quant_conv = nn.Conv2d(args.latent_dim, args.latent_dim, 1)

    def _encode(self, x_nchw_act, out=None):
        with ops.hi_region(encoder_hi_px(self.encoder.model, x_nchw_act)), ops.wsplit_region("enc"):
            h = self.encoder._fwd(x_nchw_act)
            return conv_layer(h, self.quant_conv, out=out, out_dtype=torch.float32 if out is None else None)

    def _decode(self, z, out=None):
        with ops.hi_region(decoder_hi_px(z)), ops.wsplit_region("dec"):
            h = conv_layer(z, self.post_quant_conv, out_dtype=latent_dtype(z.C))
            return self.decoder._fwd(h, out=out)

    def forward(self, x):
        return self.decode(self.encode(x))

    def encode(self, x):
        with ops.device_of(x):
            return self._encode(ops.Act.from_nchw(x)).to_nchw()

    def decode(self, z):
        with ops.device_of(z):
            return self._decode(ops.nchw_to_act(z, torch.float32)).to_nchw()

    def load_checkpoint(self, path, device=None):
        self.load_state_dict(torch.load(path, map_location=device), strict=True)
