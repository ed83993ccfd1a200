"""Two-phase (non-square 61x121, zero padding) autoencoder -- drop-in for the reference's
``modules/autoencoder2d_nonsquared.py`` (``Encoder`` :17-68, ``Decoder`` :148-247, ``SimpleAutoencoder`` :250-276).
``CondEncoder`` / ``ConditionalSimpleAutoencoder`` (:71-145, :279-305) have no caller in the reference and are out of
scope."""
import math

import torch
import torch.nn as nn

from lns_b200 import ops

from ._base import LnsModule, run_layers, conv_layer, latent_dtype, encoder_hi_px, decoder_hi_px
from .basics import ResidualBlock, SABlock, DownSampleBlock, UpSampleBlock, GroupNorm, Swish, FourierBasicBlock
from .factorized_attention import FABlock2D


class Encoder(LnsModule):
    def __init__(self, args):
        super().__init__()
        ch = args.encoder_channels
        height = args.resolutions[0]
        hw_ratio = args.hw_ratio
        assert (len(ch) - 2) == int(math.log2(height // args.latent_resolution))
        pm = "circular" if args.is_periodic else "zeros"
        layers = [nn.Conv2d(args.in_channels, ch[0], 1, 1, 0), Swish(), nn.Conv2d(ch[0], ch[0], 3, 1, 1, padding_mode=pm)]
        for i in range(len(ch) - 1):
            cin, cout = ch[i], ch[i + 1]
            for _ in range(args.encoder_res_blocks):
                layers.append(ResidualBlock(cin, cout, num_dimensions=2, padding_mode=pm))
                cin = cout
                if height in args.fourier_resolutions:
                    base = 6 if height <= 32 else 10
                    layers.append(FourierBasicBlock(cin, cout, modes=[base, int(base * hw_ratio)]))
            if i != len(ch) - 2:
                layers.append(DownSampleBlock(ch[i + 1], num_dimensions=2, padding_mode=pm))
                height //= 2
        layers += [ResidualBlock(ch[-1], ch[-1], num_dimensions=2, padding_mode=pm), GroupNorm(ch[-1]), Swish(),
                   nn.Conv2d(ch[-1], args.latent_dim, 1, 1, 0, padding_mode=pm)]
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
        pm = "circular" if args.is_periodic else "zeros"
        heads, dim = args.decoder_attn_heads, args.decoder_attn_dim
        cin = ch[0]
        height = args.latent_resolution
        hw_ratio = args.resolutions[1] / args.resolutions[0]
        no_coarse = bool(args.disable_coarse_attn) if args.disable_coarse_attn is not None else False
        rb = lambda a, b: ResidualBlock(a, b, num_dimensions=2, padding_mode=pm)  # noqa: E731

        def attn(c, h):
            if args.use_fa:
                return FABlock2D(c, dim, dim, heads, c, use_rope=True, kernel_multiplier=2)
            return SABlock(c, heads, dim, use_pe=True, block_size=h * int(h * (hw_ratio + 0.5)))

        layers = [nn.Conv2d(args.latent_dim, cin, 3, 1, 1, padding_mode=pm), rb(cin, cin)]
        if not no_coarse:
            layers.append(SABlock(cin, heads, dim, use_pe=True, block_size=height * int(height * (hw_ratio + 0.5))))
        layers.append(rb(cin, cin))
        for i in range(len(ch)):
            cout = ch[i]
            for _ in range(args.decoder_res_blocks):
                layers.append(rb(cin, cout))
                cin = cout
                if height in args.attn_resolutions:
                    layers.append(attn(cin, height))
            if i != 0 and i != len(ch) - 1:
                layers.append(UpSampleBlock(cin, num_dimensions=2, padding_mode=pm))
                height *= 2
        layers.append(nn.Upsample(size=(args.Ly, args.Lx), mode="nearest"))
        height = args.Ly
        layers.append(nn.Conv2d(cin, cin, 3, 1, 1, padding_mode=pm))
        if args.final_smoothing:
            layers.append(FourierBasicBlock(cin, cin, modes=[16, int(16 * hw_ratio)]))
        else:
            if height in args.attn_resolutions:
                layers.append(attn(cin, height))
            layers.append(nn.Conv2d(cin, cin, 3, 1, 1, padding_mode=pm))
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
