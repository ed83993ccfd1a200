"""NS2d (square, fully periodic) autoencoder -- drop-in for the reference's ``modules/autoencoder2d.py``
(``Encoder`` :16-72, ``Decoder`` :75-156, ``SimpleAutoencoder`` :160-186).  Layer order, and therefore every
``nn.Sequential`` index in the state_dict, follows the reference constructors."""
import math

import torch
import torch.nn as nn

from lns_b200 import ops

from ._base import LnsModule, run_layers, conv_layer, latent_dtype, encoder_hi_px, decoder_hi_px
from .basics import ResidualBlock, SABlock, DownSampleBlock, UpSampleBlock, GroupNorm, Swish, FourierBasicBlock
from .factorized_attention import FABlock2D


def _attention_block(args, channels, resolution, heads, dim):
    if not args.use_fa:
        return SABlock(channels, heads, dim, use_pe=True, block_size=resolution ** 2)
    return FABlock2D(channels, dim, dim, heads, channels)


class Encoder(LnsModule):
    def __init__(self, args):
        super().__init__()
        ch = args.encoder_channels
        res = args.resolution
        assert (len(ch) - 2) == int(math.log2(res // args.latent_resolution))
        # the reference reads an undefined name `padding_mode` here (modules/autoencoder2d.py:32); the Decoder derives
        # it from args.is_periodic (:84-87) and so do we
        pm = "circular" if args.is_periodic else "zeros"
        layers = [nn.Conv2d(args.in_channels, ch[0], 1, 1, 0), Swish(), nn.Conv2d(ch[0], ch[0], 3, 1, 1, padding_mode=pm)]
        for i in range(len(ch) - 1):
            cin, cout = ch[i], ch[i + 1]
            for _ in range(args.encoder_res_blocks):
                layers.append(ResidualBlock(cin, cout, num_dimensions=2, padding_mode=pm))
                cin = cout
            if res in args.attn_resolutions and args.use_attn_enc:
                layers.append(_attention_block(args, cin, res, args.attn_heads, args.attn_dim))
            if res in args.fourier_resolutions:
                layers.append(FourierBasicBlock(cin, cout, modes=[6, 6] if res <= 32 else [10, 10]))
            if i != len(ch) - 2:
                layers.append(DownSampleBlock(ch[i + 1], num_dimensions=2, padding_mode=pm))
                res //= 2
        layers += [nn.Conv2d(ch[-1], ch[-1], 3, 1, 1, padding_mode=pm), GroupNorm(ch[-1]), Swish(),
                   nn.Conv2d(ch[-1], args.latent_dim, 1, 1, 0)]
        self.model = nn.Sequential(*layers)

    def _fwd(self, x):
        return run_layers(self.model, x, final_dtype=torch.float32)  # the latent stays fp32

    def forward(self, x):
        with ops.device_of(x):
            return self._fwd(ops.Act.from_nchw(x)).to_nchw()


class Decoder(LnsModule):
    def __init__(self, args):
        super().__init__()
        ch = args.decoder_channels
        pm = "circular" if args.is_periodic else "zeros"
        heads, dim = args.attn_heads, args.attn_dim
        no_coarse = bool(args.disable_coarse_attn) if args.disable_coarse_attn is not None else False
        cin = ch[0]
        res = args.latent_resolution
        rb = lambda a, b: ResidualBlock(a, b, num_dimensions=2, padding_mode=pm)  # noqa: E731
        layers = [nn.Conv2d(args.latent_dim, cin, 1, 1, 0), rb(cin, cin)]
        if not no_coarse:
            layers.append(SABlock(cin, heads, dim, use_pe=True, block_size=res ** 2))
        layers.append(rb(cin, cin))
        for i in range(len(ch)):
            cout = ch[i]
            for _ in range(args.decoder_res_blocks):
                layers.append(rb(cin, cout))
                cin = cout
            if res in args.attn_resolutions:
                layers.append(_attention_block(args, cin, res, heads, dim))
            if i != 0 and i != len(ch) - 1:
                layers.append(UpSampleBlock(cin, num_dimensions=2, padding_mode=pm))
                res *= 2
        layers.append(nn.Upsample(size=(args.Ly, args.Lx), mode="nearest"))
        res = args.Ly
        layers.append(nn.Conv2d(cin, cin, 3, 1, 1, padding_mode=pm))
        if args.final_smoothing:
            layers.append(FourierBasicBlock(cin, cin, modes=[16, 16]))
        else:
            if res in args.attn_resolutions:
                layers.append(_attention_block(args, cin, res, heads, dim))
            layers.append(nn.Conv2d(cin, cin, 1, 1, 0, padding_mode=pm))
        layers += [nn.GroupNorm(8, cin), Swish(), nn.Conv2d(cin, args.in_channels, 1, 1, 0)]
        self.model = nn.Sequential(*layers)

    def _fwd(self, x, out=None):
        return run_layers(self.model, x, final_out=out, final_layout=ops.NCHW)

    def forward(self, x):
        with ops.device_of(x):
            return self._fwd(ops.nchw_to_act(x)).to_nchw()


class SimpleAutoencoder(LnsModule):
    """encode = quant_conv(encoder(x)), decode = decoder(post_quant_conv(z)) (reference :160-186).
    Latents are kept in fp32 in both precision modes (they are tiny and autoregressively re-used)."""

    def __init__(self, args):
        super().__init__()
        self.encoder = Encoder(args)
        self.decoder = Decoder(args)
        self.quant_conv = nn.Conv2d(args.latent_dim, args.latent_dim, 1)
        self.post_quant_conv = nn.Conv2d(args.latent_dim, args.latent_dim, 1)

    # -- Act-level entry points used by lns_b200.rollout --
    def _encode(self, x_nchw_act, out=None):
        with ops.hi_region(encoder_hi_px(self.encoder.model, x_nchw_act)), ops.wsplit_region("enc"):
            h = self.encoder._fwd(x_nchw_act)
            return conv_layer(h, self.quant_conv, out=out, out_dtype=torch.float32 if out is None else None)

    def _decode(self, z, out=None):
        with ops.hi_region(decoder_hi_px(z)), ops.wsplit_region("dec"):
            h = conv_layer(z, self.post_quant_conv, out_dtype=latent_dtype(z.C))
            return self.decoder._fwd(h, out=out)

    # -- the reference's tensor API --
    def forward(self, x):
        return self.decode(self.encode(x))

    def encode(self, x):
        with ops.device_of(x):
            return self._encode(ops.Act.from_nchw(x)).to_nchw()

    def decode(self, z):
        with ops.device_of(z):
            return self._decode(ops.nchw_to_act(z, torch.float32)).to_nchw()

    def load_checkpoint(self, path):
        self.load_state_dict(torch.load(path), strict=True)
