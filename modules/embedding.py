"""Rotary position embedding used inside FABlock2D -- drop-in for the rotary part of the reference's
``modules/embedding.py`` (``RotaryEmbedding`` :163-176, ``rotate_half`` :179-182, ``apply_rotary_pos_emb`` :185-186).
The Siren / 2-D / 3-D embedding classes of that file are dead code in the reference and are not provided."""
import torch
from torch import nn


class RotaryEmbedding(nn.Module):
    """Holds the ``inv_freq`` buffer (it is part of the state_dict).  ``angles`` returns the table the
    lns_lowrank_kernel consumes: angle[i][f] = coord[i] * (scale / min_freq) * inv_freq[f]."""

    def __init__(self, dim, min_freq=1 / 64, scale=1.):
        super().__init__()
        inv_freq = 1. / (10000 ** (torch.arange(0, dim, 2).float() / dim))
        self.min_freq = min_freq
        self.scale = scale
        self.register_buffer("inv_freq", inv_freq)

    def angles(self, coordinates):
        """coordinates [n] -> angles [n, dim/2] in float64 (rounded to fp32 only after cos / sin)."""
        t = coordinates.to(self.inv_freq.device).double() * (self.scale / self.min_freq)
        return t[:, None] * self.inv_freq.double()[None, :]

    def forward(self, coordinates, device):
        t = coordinates.to(device).type_as(self.inv_freq) * (self.scale / self.min_freq)
        freqs = torch.einsum("... i , j -> ... i j", t, self.inv_freq)
        return torch.cat((freqs, freqs), dim=-1)


def rotate_half(x):
    x1, x2 = x.chunk(2, dim=-1)
    return torch.cat((-x2, x1), dim=-1)


def apply_rotary_pos_emb(t, freqs):
    return (t * freqs.cos()) + (rotate_half(t) * freqs.sin())
