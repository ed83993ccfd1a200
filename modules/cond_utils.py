"""Conditioning helpers -- drop-in for the part of the reference's ``modules/cond_utils.py`` that the rollout path uses
(``zero_module``, ``fourier_embedding``, ``ConditionedBlock``; the PDEArena ``CondResidualBlock`` is only reachable
from the unused ``CondEncoder`` and is out of scope)."""
import math
from abc import abstractmethod

import torch
from torch import nn


def zero_module(module):
    """Zero all parameters of `module` and return it (reference: modules/cond_utils.py:12-16)."""
    for p in module.parameters():
        p.detach().zero_()
    return module


def fourier_embedding(timesteps: torch.Tensor, dim, max_period=10000):
    """Sinusoidal embedding cat(cos(t f), sin(t f)), f_i = exp(-ln(max_period) i / (dim//2))
    (reference: modules/cond_utils.py:19-38).  CUDA input -> lns_fourier_embedding kernel."""
    from lns_b200 import ops
    if not timesteps.is_cuda:
        raise ops.LnsError("fourier_embedding: CUDA tensors only (no CPU fallback)")
    return ops.fourier_embedding(timesteps, dim, float(max_period))


class ConditionedBlock(nn.Module):
    @abstractmethod
    def forward(self, x, emb):
        """Apply the module to `x` given the embedding `emb`."""


class EmbedSequential(nn.Sequential, ConditionedBlock):
    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, ConditionedBlock) else layer(x)
        return x
