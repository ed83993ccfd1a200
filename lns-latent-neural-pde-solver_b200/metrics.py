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


def relative_l2_twophase(pred, target, vel_mean, vel_std, prs_mean, prs_std, eps=1e-8):
    """relative_lp_loss of the de-normalised two-phase fields (vx, vy, p, vof), frame-wise [B, K, C] and sequence-wise [B, C], with
    the reference's ``denormalize`` (dataset/twophase_flow_stage2.py:369-389: velocity / pressure statistics, the velocity zeroed
    on the four closed walls, vof clamped to [0, 1 + 1e-8]) fused into the ONE read of both tensors (lns_frame_sums_denorm)."""
    if not (pred.is_cuda and target.is_cuda):
        raise LnsError("relative_l2_twophase: CUDA tensors only (there is no CPU fallback)")
    if pred.shape != target.shape or pred.dim() != 5 or pred.shape[2] != 4 or pred.dtype != torch.float32 or target.dtype != torch.float32:
        raise LnsError("relative_l2_twophase: expected two fp32 tensors [B, K, 4, Ly, Lx] (vx, vy, p, vof)")
    pred, target = pred.contiguous(), target.contiguous()
    B, K, C, Ly, Lx = pred.shape
    dev = pred.device
    scale = torch.tensor([vel_std, vel_std, prs_std, 1.0], dtype=torch.float32, device=dev)
    shift = torch.tensor([vel_mean, vel_mean, prs_mean, 0.0], dtype=torch.float32, device=dev)
    flags = torch.tensor([1, 1, 0, 2], dtype=torch.int32, device=dev)
    out = torch.empty(B, K, C, 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = _C.lib().lns_frame_sums_denorm(ctypes.c_void_p(pred.data_ptr()), ctypes.c_void_p(target.data_ptr()), B * K * C, Ly, Lx, C,
                                            ctypes.c_void_p(scale.data_ptr()), ctypes.c_void_p(shift.data_ptr()),
                                            ctypes.c_void_p(flags.data_ptr()), 0.0, 1.0 + 1e-8, ctypes.c_void_p(out.data_ptr()),
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    check(rc, "lns_frame_sums_denorm")
    s = out.double()
    frame = (s[..., 0] / s[..., 1].clamp_min(eps)).sqrt()
    seq = (s[..., 0].sum(1) / s[..., 1].sum(1).clamp_min(eps)).sqrt()
    return frame.float(), seq.float()
