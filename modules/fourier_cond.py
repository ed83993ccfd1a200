"""Frequency-conditioned Fourier block -- drop-in for the reference's ``modules/fourier_cond.py``
(``FreqLinear`` :16-29, ``SpectralConv2d`` :32-81, ``CondFourierBasicBlock`` :84-118).  No reference script
instantiates it; it is kept API-stable and benchmarked stand-alone."""
import torch
from torch import nn

from lns_b200 import ops

from ._base import LnsModule, conv_layer, filt_of, cache_of
from .cond_utils import ConditionedBlock


class FreqLinear(LnsModule):
    """emb [B, Ccond] -> per-sample complex scale of every retained mode, [B, m1, m2, 2(block)] complex
    (reference :16-29: einsum + bias, reshape (B,m1,m2,2,2), view_as_complex)."""

    def __init__(self, in_channel, modes1, modes2):
        super().__init__()
        self.modes1, self.modes2 = modes1, modes2
        scale = 1 / (in_channel + 4 * modes1 * modes2)
        self.weights = nn.Parameter(scale * torch.randn(in_channel, 4 * modes1 * modes2, dtype=torch.float32))
        self.bias = nn.Parameter(torch.zeros(1, 4 * modes1 * modes2, dtype=torch.float32))

    def _filter(self):
        f = cache_of(self).get("f")
        if f is None:
            w, b = self.weights, self.bias

            def wfn():  # [Ccond, 4 m1 m2] -> OIHW [4 m1 m2, Ccond, 1, 1]
                return w.detach().t()[:, :, None, None]

            def bfn():
                return b.detach().reshape(-1)
            f = ops.PackedFilter(wfn, bfn, lambda: (ops.PackedFilter._fp(w), ops.PackedFilter._fp(b)),
                                 (w.shape[1], w.shape[0], 1, 1))
            cache_of(self)["f"] = f
        return f

    def _fwd(self, emb_rows):
        """emb_rows: Act [B,1,1,Ccond] fp32 -> torch fp32 [B, m1, m2, 2, 2] (.., block, re/im)"""
        h = ops.conv2d(emb_rows, self._filter(), out_dtype=torch.float32)
        return h.t.view(emb_rows.B, self.modes1, self.modes2, 2, 2)

    def forward(self, x):
        return torch.view_as_complex(self._fwd(ops.rows_act(x.float())))


class SpectralConv2d(LnsModule):
    """rfft2 -> (modes * per-sample emb) x weights -> irfft2 (reference :32-81)."""

    def __init__(self, in_channels, out_channels, cond_channels, modes1, modes2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.modes1, self.modes2 = modes1, modes2
        self.scale = 1 / (in_channels * out_channels)
        self.weights1 = nn.Parameter(self.scale * torch.rand(in_channels, out_channels, modes1, modes2, 2,
                                                             dtype=torch.float32))
        self.weights2 = nn.Parameter(self.scale * torch.rand(in_channels, out_channels, modes1, modes2, 2,
                                                             dtype=torch.float32))
        self.cond_emb = FreqLinear(cond_channels, modes1, modes2)

    def _mode_weights(self):
        key = (self.weights1.data_ptr(), self.weights1._version, self.weights2.data_ptr(), self.weights2._version)
        ent = cache_of(self).get("wm")
        if ent is None or ent[0] != key:
            w = torch.stack([self.weights1.detach(), self.weights2.detach()], 0)
            ent = (key, w.permute(0, 3, 4, 1, 2, 5).contiguous().float())
            cache_of(self)["wm"] = ent
        return ent[1]

    def _fwd(self, x, emb_rows):
        emb = self.cond_emb._fwd(emb_rows)
        return ops.spectral_conv2d(x, self._mode_weights(), self.modes1, self.modes2, self.out_channels, emb=emb)

    def forward(self, x, emb):
        return self._fwd(ops.nchw_to_act(x), ops.rows_act(emb.float())).to_nchw()


class CondFourierBasicBlock(ConditionedBlock, LnsModule):
    """x + GELU(fourier(x, emb) + Conv1x1(x) + Linear(emb)[..., None, None])  (reference :84-118)."""

    expansion: int = 1

    def __init__(self, in_planes, planes, modes, residual=True):
        super().__init__()
        self.modes = modes
        self.num_dimensions = len(modes)
        assert self.num_dimensions == 2
        self.residual = residual
        self.fourier = SpectralConv2d(in_planes, planes, in_planes, modes[0], modes[1])
        self.conv = nn.Conv2d(in_planes, planes, kernel_size=1, stride=1, padding=0)
        self.cond_emb = nn.Linear(in_planes, planes)

    def _fwd(self, x, emb_rows):
        spec = self.fourier._fwd(x, emb_rows)
        shift = ops.conv2d(emb_rows, filt_of(self.cond_emb), out_dtype=torch.float32)  # [B, planes]
        return conv_layer(x, self.conv, pre_add=spec, sample_bias=shift.t, act=ops.ACT_GELU,
                          residual=x if self.residual else None)

    def forward(self, x: torch.Tensor, cond_emb: torch.Tensor) -> torch.Tensor:
        return self._fwd(ops.nchw_to_act(x), ops.rows_act(cond_emb.float())).to_nchw()
