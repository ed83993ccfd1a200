"""Building blocks of the LNS autoencoders -- drop-in for the reference's ``modules/basics.py``.

Same class names, constructor signatures and parameter names as the reference (cited per class); the arithmetic runs
in liblns_b200.so.  Only what the rollout path constructs is provided (``LABlock``, ``CABlock``,
the 1-D/3-D variants and the FNO mixer blocks have no caller and are out of scope).
"""
import math
from typing import List

import torch
import torch.nn as nn

from lns_b200 import ops

from ._base import LnsModule, LnsError, conv_layer, norm_affine, lazy_norm, pad_modes, filt_of, cache_of
from .cond_utils import zero_module, ConditionedBlock  # noqa: F401  (re-exported like the reference does)


class GroupNorm(LnsModule):
    """32-group GroupNorm, eps 1e-6 (reference: modules/basics.py:18-24)."""

    def __init__(self, channels):
        super().__init__()
        self.gn = nn.GroupNorm(num_groups=32, num_channels=channels, eps=1e-6, affine=True)

    def _fwd(self, x):
        scale, shift = norm_affine(x, self)
        return ops.affine_act(x, scale, shift, ops.ACT_NONE)


class Swish(LnsModule):
    """x * sigmoid(x) (reference: modules/basics.py:27-29)."""

    def _fwd(self, x):
        return ops.affine_act(x, None, None, ops.ACT_SILU)


def _only_2d(num_dimensions, what):
    if num_dimensions != 2:
        raise NotImplementedError(f"{what}: only num_dimensions=2 is on the rollout path (got {num_dimensions})")


class ResidualBlock(LnsModule):
    """[channel_up 1x1](x) + Conv3x3(Swish(GN(Conv3x3(Swish(GN(x))))))  (reference: modules/basics.py:224-276).

    Two statistics kernels + two implicit-GEMM convs: each GroupNorm+Swish is folded into the following conv's
    prologue, the skip connection into the second conv's epilogue."""

    def __init__(self, in_channels, out_channels, num_dimensions, padding_mode="zeros"):
        super().__init__()
        _only_2d(num_dimensions, "ResidualBlock")
        self.in_channels, self.out_channels, self.num_dimensions = in_channels, out_channels, num_dimensions
        self.block = nn.Sequential(
            GroupNorm(in_channels), Swish(),
            nn.Conv2d(in_channels, out_channels, 3, 1, 1, padding_mode=padding_mode),
            GroupNorm(out_channels), Swish(),
            nn.Conv2d(out_channels, out_channels, 3, 1, 1, padding_mode=padding_mode))
        if in_channels != out_channels:
            self.channel_up = nn.Conv2d(in_channels, out_channels, 1, 1, 0)

    def _fwd(self, x):
        gn1, _, conv1, gn2, _, conv2 = self.block
        h = conv_layer(x, conv1, pro=lazy_norm(x, gn1, ops.ACT_SILU))
        skip = conv_layer(x, self.channel_up) if self.in_channels != self.out_channels else x
        return conv_layer(h, conv2, pro=lazy_norm(h, gn2, ops.ACT_SILU), residual=skip)


class UpSampleBlock(LnsModule):
    """nearest x2 then Conv3x3 (reference: modules/basics.py:279-299); the replicate is folded into the conv's
    gather, the 4x larger tensor never exists."""

    def __init__(self, channels, num_dimensions, padding_mode="zeros"):
        super().__init__()
        _only_2d(num_dimensions, "UpSampleBlock")
        self.num_dimensions = num_dimensions
        self.conv_layer = nn.Conv2d(channels, channels, 3, 1, 1, padding_mode=padding_mode)

    def _fwd(self, x):
        return conv_layer(x, self.conv_layer, virt=(2 * x.H, 2 * x.W))


class DownSampleBlock(LnsModule):
    """F.pad then Conv3x3 stride 2 pad 0 (reference: modules/basics.py:302-328): circular mode pads (1,1,1,1)
    circularly, zeros mode pads (0,1,0,1) with zeros; the conv itself always has zero-mode weights."""

    def __init__(self, channels, num_dimensions, padding_mode="zeros"):
        super().__init__()
        _only_2d(num_dimensions, "DownSampleBlock")
        self.num_dimensions = num_dimensions
        self.conv_layer = nn.Conv2d(channels, channels, 3, 2, 0)
        self.padding_mode = padding_mode if padding_mode != "zeros" else "constant"
        self.pad = []
        for _ in range(num_dimensions):
            self.pad.extend((1, 1) if self.padding_mode == "circular" else (0, 1))

    def _fwd(self, x):
        left, right, top, bottom = self.pad
        modes = pad_modes("circular" if self.padding_mode == "circular" else "zeros")
        return ops.conv2d(x, filt_of(self.conv_layer), stride=2, pad=(top, bottom, left, right), pad_mode=modes)


class SABlock(LnsModule):
    """Pre-norm multi-head self-attention over the pixels of the coarse latent grid
    (reference: modules/basics.py:331-404): LN -> +pe (after the norm) -> q,k (no bias), v -> softmax(q k^T / sqrt(dh)) v
    -> proj -> + input.  q|k|v are one GEMM; the residual add sits in the projection's epilogue."""

    def __init__(self, dim, heads, dim_head, use_pe=False, block_size=512):
        super().__init__()
        self.dim, self.heads, self.dim_head = dim, heads, dim_head
        self.ln = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, heads * dim_head, bias=False)
        self.to_k = nn.Linear(dim, heads * dim_head, bias=False)
        self.to_v = nn.Linear(dim, heads * dim_head)
        self.proj_out = nn.Linear(heads * dim_head, dim)
        self.pe = nn.Parameter(torch.randn(1, block_size, dim) * 0.02, requires_grad=True) if use_pe else None
        self.init_params()

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def init_params(self):
        for m in self.modules():
            self._init_weights(m)

    def _qkv_filter(self):
        c = cache_of(self)
        srcs = (self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_v.bias)
        ent = c.get("qkv")
        if ent is None or any(a is not b for a, b in zip(ent[0], srcs)):
            ent = (srcs, ops.PackedFilter.concat([self.to_q.weight, self.to_k.weight, self.to_v.weight],
                                                 [None, None, self.to_v.bias]))
            c["qkv"] = ent
        return ent[1]

    def _fused_operands(self, dt16):
        """16-bit filter copies / fp32 vectors of lns_sablock_fused, cached until a source parameter changes."""
        srcs = [self.ln.weight, self.ln.bias, self.to_q.weight, self.to_k.weight, self.to_v.weight, self.to_v.bias,
                self.proj_out.weight, self.proj_out.bias] + ([self.pe] if self.pe is not None else [])
        key = (dt16,) + tuple(None if t is None else (t.data_ptr(), t._version, str(t.device)) for t in srcs)
        ent = cache_of(self).get("sa")
        if ent is None or ent[0] != key:
            with torch.no_grad():
                f = lambda t: None if t is None else t.detach().float().contiguous()  # noqa: E731
                ent = (key, dict(
                    ln_g=f(self.ln.weight), ln_b=f(self.ln.bias),
                    wqkv=torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0).detach().to(dt16).contiguous(),
                    bv=f(self.to_v.bias), wproj=self.proj_out.weight.detach().to(dt16).contiguous(), bproj=f(self.proj_out.bias),
                    pe=None if self.pe is None else self.pe.detach().float().reshape(-1, self.dim).contiguous()))
            cache_of(self)["sa"] = ent
        return ent[1]

    def _fwd(self, x):
        x = ops.as_h16(x)
        n = x.H * x.W
        pe = None
        if self.pe is not None:
            if n > self.pe.shape[1]:
                raise LnsError(f"SABlock: {n} tokens exceed the positional table ({self.pe.shape[1]})")
            pe = self.pe.detach()
        if ops.sablock_fused_supported(x, self.heads, self.dim_head) and self.to_q.bias is None and self.to_k.bias is None:
            o = self._fused_operands(x.t.dtype)
            return ops.sablock_fused(x, self.heads, o["ln_g"], o["ln_b"], self.ln.eps,
                                     None if pe is None else o["pe"], o["wqkv"], o["bv"], o["wproj"], o["bproj"],
                                     self.dim_head ** (-0.5))
        t = ops.layernorm(x, self.ln.weight, self.ln.bias, self.ln.eps, pe=pe)
        # on the 16-bit paths q|k|v stay 16-bit whatever the surrounding stage stores (the split-operand mode keeps the coarse
        # stage in fp32): the tensor-core attention kernel takes 16-bit rows, and the block contributes < 1 % of the 16-bit error
        # (tools/precision_study.py) -- an fp32 q|k|v would drop to the CUDA-core attention kernel (30x slower at 288 tokens)
        qkv = ops.conv2d(t, self._qkv_filter(), out_dtype=ops.act_dtype() if ops.fast16() else None)
        o = ops.attention(qkv, self.heads, self.dim_head, self.dim_head ** (-0.5))
        return ops.conv2d(o, filt_of(self.proj_out), residual=x)

    def forward(self, x, channel_last=False):
        if channel_last:  # [b, n, c] tokens
            b, n, c = x.shape
            a = ops.nchw_to_act(x.transpose(1, 2).reshape(b, c, n, 1))
            return self._fwd(a).to_nchw().reshape(b, c, n).transpose(1, 2)
        return super().forward(x)


class SpectralConv2d(LnsModule):
    """2-D Fourier layer: rfft2 -> two corner blocks x complex weights -> irfft2
    (reference: modules/basics.py:99-149).  Runs as truncated DFTs with the mode multiply fused (csrc/spectral.cu)."""

    def __init__(self, in_channels: int, out_channels: int, modes1: int, modes2: int):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1 / (in_channels * out_channels)
        self.weights1 = nn.Parameter(self.scale * torch.rand(in_channels, out_channels, modes1, modes2, 2,
                                                             dtype=torch.float32))
        self.weights2 = nn.Parameter(self.scale * torch.rand(in_channels, out_channels, modes1, modes2, 2,
                                                             dtype=torch.float32))

    def _mode_weights(self):
        """[Ci,Co,m1,m2,2] x2 -> [2,m1,m2,Ci,Co,2] (device-side re-layout with torch, cached per parameter version)."""
        key = (self.weights1.data_ptr(), self.weights1._version, self.weights2.data_ptr(), self.weights2._version)
        ent = cache_of(self).get("wm")
        if ent is None or ent[0] != key:
            w = torch.stack([self.weights1.detach(), self.weights2.detach()], 0)  # [2,Ci,Co,m1,m2,2]
            ent = (key, w.permute(0, 3, 4, 1, 2, 5).contiguous().float())
            cache_of(self)["wm"] = ent
        return ent[1]

    def _fwd(self, x, emb=None):
        return ops.spectral_conv2d(x, self._mode_weights(), self.modes1, self.modes2, self.out_channels, emb=emb)

    def forward(self, x, x_dim=None, y_dim=None):
        return super().forward(x)


class FourierBasicBlock(LnsModule):
    """x + GELU(SpectralConv2d(x) + Conv1x1(x))  (reference: modules/basics.py:531-583).  The sum, the activation and
    the skip are the 1x1 conv's epilogue."""

    expansion: int = 1

    def __init__(self, in_planes: int, planes: int, modes: List[int], activation: str = "gelu", residual: bool = True):
        super().__init__()
        self.modes = modes
        self.num_dimensions = len(modes)
        self.residual = residual
        if self.num_dimensions != 2:
            raise NotImplementedError("FourierBasicBlock: only 2-D is on the rollout path")
        self.fourier = SpectralConv2d(in_planes, planes, modes[0], modes[1])
        self.conv = nn.Conv2d(in_planes, planes, kernel_size=1, stride=1, padding=0)
        acts = {"gelu": ops.ACT_GELU, "silu": ops.ACT_SILU}
        if activation not in acts:
            raise NotImplementedError(f"Activation {activation} not implemented")
        self.activation = nn.GELU() if activation == "gelu" else nn.SiLU()
        self._act = acts[activation]

    def _fwd(self, x):
        spec = self.fourier._fwd(x)
        return conv_layer(x, self.conv, pre_add=spec, act=self._act, residual=x if self.residual else None)
