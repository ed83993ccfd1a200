"""Validation metric of the stage-2 scripts, fused behind the rollout.

``train_stage2_*.py::validate_loop`` de-normalises prediction and target and calls
``relative_lp_loss(y_hat, y, reduce_dim=(3, 4), p=2)`` (frame-wise) and ``reduce_dim=(1, 3, 4)`` (sequence-wise)
(train_stage2_ns2d.py:254-257, training_utils.py:9-23).  Here ONE kernel reads both tensors once and leaves three sums per
frame on the device; the de-normalisation ``x*std + mean`` (dataset/ns2d_fno_stage2_simpleae.py:78-79) is applied to the
sums, so neither de-normalised copy exists and nothing but [B, K, C] numbers has to be gathered across GPUs."""
import ctypes

import torch

from . import _C
from ._C import LnsError, check


def frame_sums(pred, target):
    """pred, target: CUDA fp32 [B, K, C, Ly, Lx] -> fp32 [B, K, C, 3] = (sum (p-t)^2, sum t^2, sum t) per frame."""
    if not (pred.is_cuda and target.is_cuda):
        raise LnsError("frame_sums: CUDA tensors only (there is no CPU fallback)")
    if pred.shape != target.shape or pred.dim() != 5 or pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise LnsError("frame_sums: expected two fp32 tensors of the same shape [B, K, C, Ly, Lx]")
    pred, target = pred.contiguous(), target.contiguous()
    B, K, C, Ly, Lx = pred.shape
    out = torch.empty(B, K, C, 3, dtype=torch.float32, device=pred.device)
    rc = _C.lib().lns_frame_sums(ctypes.c_void_p(pred.data_ptr()), ctypes.c_void_p(target.data_ptr()), B * K * C, Ly * Lx,
                                 ctypes.c_void_p(out.data_ptr()),
                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "lns_frame_sums")
    return out


def relative_l2(pred, target, mean=0.0, std=1.0, eps=1e-8):
    """(frame_wise [B, K, C], seq_wise [B, C]) = relative_lp_loss(denorm(pred), denorm(target), reduce_dim=(3,4) / (1,3,4))
    with denorm(x) = x*std + mean (scalars or tensors broadcastable to [B, K, C])."""
    s = frame_sums(pred, target).double()
    P = pred.shape[-1] * pred.shape[-2]
    std = torch.as_tensor(std, dtype=torch.float64, device=s.device)
    mean = torch.as_tensor(mean, dtype=torch.float64, device=s.device)
    diff = std * std * s[..., 0]
    gt = std * std * s[..., 1] + 2.0 * std * mean * s[..., 2] + P * mean * mean
    frame = (diff / gt.clamp_min(eps)).sqrt()
    seq = (diff.sum(1) / gt.sum(1).clamp_min(eps)).sqrt()
    return frame.float(), seq.float()
