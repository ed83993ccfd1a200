"""Bulk pre-encoding of a dataset with the frozen autoencoder: the second caller of ``encode``.

The reference walks its cases one at a time, normalises on the host and calls ``vq_ae.encode`` per case
(dataset/ns2d_fno_stage2_simpleae.py:81-93, dataset/Stage2_SW.py:80-105, dataset/twophase_flow_stage2.py:317-337).  Here the
frames stream through fixed-size chunks: pinned host staging, the host->device copy of chunk i+1 and the device->host copy
of chunk i-1 overlap the encoder of chunk i (three streams, events), and the encoder itself is the same kernel sequence as the
rollout's first stage."""
import numpy as np
import torch

from ._C import LnsError
from . import ops


@torch.no_grad()
def encode_frames(ae, frames, chunk=1024, mean=0.0, std=1.0, eps=1e-8, precision=None, device=None):
    """frames: numpy or CPU torch array [N, C, H, W] (raw values) -> numpy fp32 [N, Cz, h, w] = ae.encode((x-mean)/(std+eps)).
    The normalisation is applied on the device (the reference does it with numpy per case)."""
    x = torch.as_tensor(np.ascontiguousarray(frames)) if not torch.is_tensor(frames) else frames
    if x.dim() != 4:
        raise LnsError("encode_frames: expected [N, C, H, W]")
    device = torch.device(device or next(ae.parameters()).device)
    if device.type != "cuda":
        raise LnsError("encode_frames: the autoencoder must live on a CUDA device (there is no CPU fallback)")
    N = x.shape[0]
    chunk = max(1, min(chunk, N))
    copy_in, copy_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)
    stage_in = [torch.empty((chunk,) + tuple(x.shape[1:]), dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_in = [torch.empty((chunk,) + tuple(x.shape[1:]), dtype=torch.float32, device=device) for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_host = None
    scale, shift = 1.0 / (float(std) + eps), -float(mean) / (float(std) + eps)
    nchunks = (N + chunk - 1) // chunk
    norm_scale = torch.full((chunk * 4,), scale, dtype=torch.float32, device=device)
    norm_shift = torch.full((chunk * 4,), shift, dtype=torch.float32, device=device)

    def stage(i):
        s, n = i & 1, min(chunk, N - i * chunk)
        # the H2D copy of chunk i-2 read this pinned buffer asynchronously (it is queued behind the encoder of chunk i-4): the
        # host may only re-pack the buffer once that copy has completed (never recorded on first use: returns at once)
        in_ready[s].synchronize()
        stage_in[s][:n].copy_(x[i * chunk:i * chunk + n])           # host pack into pinned memory
        with torch.cuda.stream(copy_in):
            copy_in.wait_event(in_free[s])                           # the encoder has consumed this slot's previous chunk
            dev_in[s][:n].copy_(stage_in[s][:n], non_blocking=True)
            in_ready[s].record(copy_in)

    for s in range(2):
        in_free[s].record(compute)
    stage(0)
    with ops.precision(precision or ops.get_precision()):
        for i in range(nchunks):
            s, n = i & 1, min(chunk, N - i * chunk)
            compute.wait_event(in_ready[s])
            xin = dev_in[s][:n]
            if scale != 1.0 or shift != 0.0:
                # (x - mean) / (std + eps) on the device with the library's affine kernel: a contiguous NCHW sample is viewed as
                # rows of 4 elements (C*H*W % 4 == 0 on every configuration), scale / shift are per-(sample, lane) constants
                per = xin[0].numel()
                if per % 4 != 0:
                    raise LnsError("encode_frames: C*H*W must be a multiple of 4")
                a = ops.Act(xin.reshape(-1), n, 1, per // 4, 4)
                xin = ops.affine_act(a, norm_scale[:n * 4], norm_shift[:n * 4], ops.ACT_NONE, out_dtype=torch.float32).t.view(xin.shape)
            z = ae.encode(xin)                                       # [n, Cz, h, w] fp32
            in_free[s].record(compute)
            done = torch.cuda.Event()
            done.record(compute)
            if i + 1 < nchunks:
                stage(i + 1)                                         # overlaps the encoder of chunk i
            if out_host is None:
                out_host = torch.empty((N,) + tuple(z.shape[1:]), dtype=torch.float32).pin_memory()
            with torch.cuda.stream(copy_out):
                copy_out.wait_event(done)
                out_host[i * chunk:i * chunk + n].copy_(z, non_blocking=True)
            z.record_stream(copy_out)                                # the allocator may not recycle z before its copy ran
    copy_out.synchronize()
    return out_host.numpy().copy()


def encode_dataset(ae, data, mean, std, chunk=1024, **kw):
    """NS2d layout of the reference: data [T, H, W, cases] (dataset/ns2d_fno_stage2_simpleae.py:81-93) -> list over cases of
    numpy [T, Cz, h, w], like ``self.encoded_data``."""
    data = np.asarray(data)
    T, H, W, cases = data.shape
    frames = np.ascontiguousarray(np.transpose(data, (3, 0, 1, 2)).reshape(cases * T, 1, H, W)).astype(np.float32)
    z = encode_frames(ae, frames, chunk=chunk, mean=mean, std=std, **kw)
    return [z[c * T:(c + 1) * T] for c in range(cases)]
