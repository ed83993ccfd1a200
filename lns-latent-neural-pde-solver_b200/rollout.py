"""The latent rollout engine: encode -> K autoregressive propagator steps -> decode, as ONE CUDA graph.

Reference semantics: ``LatentDynamics.predict(x, steps, to_x)`` (train_stage2_ns2d.py:143-158; conditional variant
train_stage2_twophase_conditional.py:177-193).  Differences in execution, not in results:
  * the K latent states are written by the propagator's output projection directly into a resident
    [B, K, h, w, Cz] fp32 stack (batch-strided output, no torch.stack copy);
  * decode never feeds back into the loop, so the K decodes are batched: the latent stack is step-major and every S steps
    their B*S samples are decoded as one launch group ON A SECOND STREAM, overlapping the rest of the propagator loop (whose
    launches are small and latency-bound: ~1 ms per step for ~0.25 ms of tensor work); the decoder's output projection
    writes NCHW fp32 straight into slot [b][t] of the [B, K, C, Ly, Lx] result (lns_pointwise_proj_steps).
    LNS_ROLLOUT_PIPELINE=0 restores the serial order (all steps, then the decode in chunks);
  * everything that depends only on the conditioning parameter is computed once, not per step;
  * the whole sequence is captured in a CUDA graph and replayed (no Python / launch overhead per step).
"""
import os

import torch

from . import ops
from .ops import Act, LnsError


def _find_parts(model):
    ae = getattr(model, "vq_ae", None) or getattr(model, "ae", None)
    prop = getattr(model, "propagator", None)
    if ae is None or prop is None:
        raise LnsError("Rollout: model must have .vq_ae (or .ae) and .propagator like the reference's LatentDynamics")
    if not hasattr(ae, "_encode"):
        raise LnsError("Rollout: the autoencoder must be one of this repository's modules.* SimpleAutoencoder classes "
                       "(load the reference state_dict into them; the layouts are identical)")
    return ae, prop


class Rollout:
    """Compiled rollout for a fixed (batch, steps).  Call with x [B, C, Ly, Lx] (fp32, CUDA) [and param [B]];
    returns the engine's static output buffer [B, steps, C, Ly, Lx] (to_x) or [B, steps, Cz, h, w] -- clone it if it
    must survive the next call."""

    def __init__(self, model, batch, steps, to_x=True, precision=None, use_graph=True, decode_chunk=None,
                 device=None, pipeline=None):
        from modules.propagator import simple_cnn_fwd, cond_cnn_fwd, cond_cnn_prepare  # drop-in package at repo root
        self._simple_cnn_fwd, self._cond_cnn_fwd, self._cond_cnn_prepare = simple_cnn_fwd, cond_cnn_fwd, cond_cnn_prepare
        self.model = model
        self.ae, self.prop = _find_parts(model)
        self.conditional = hasattr(self.prop, "cond_emb_proj")
        self.B, self.K, self.to_x = int(batch), int(steps), bool(to_x)
        self.precision = precision or ops.get_precision()
        p = next(self.ae.parameters())
        self.device = torch.device(device) if device is not None else p.device
        if self.device.type != "cuda":
            raise LnsError("Rollout: the model must live on a CUDA device (no CPU fallback)")
        enc0 = self.ae.encoder.model[0]
        dec_last = self.ae.decoder.model[-1]
        up = [m for m in self.ae.decoder.model if isinstance(m, torch.nn.Upsample)][0]
        self.C = dec_last.out_channels
        self.Cin = enc0.in_channels
        self.Ly, self.Lx = up.size
        self.Cz = self.ae.quant_conv.out_channels
        self.x_static = torch.zeros(self.B, self.Cin, self.Ly, self.Lx, dtype=torch.float32, device=self.device)
        self.param_static = torch.zeros(self.B, dtype=torch.float32, device=self.device) if self.conditional else None
        self.decode_chunk = decode_chunk
        self.pipeline = (os.environ.get("LNS_ROLLOUT_PIPELINE", "1") != "0") if pipeline is None else bool(pipeline)
        self.steps_per_group = 0
        self._side = None
        self.use_graph = use_graph
        self.graph = None
        self.out = None
        self.zs = None
        self.launches_per_call = None
        self._built = False
        self._fp = None

    def _fingerprint(self):
        """(data_ptr, version) of every parameter and buffer the captured launch sequence depends on.  The graph bakes in the
        addresses of the packed filter images (built from these tensors during the eager warm-up): after load_state_dict, an
        optimizer step or .to() a replay would keep using stale -- or freed -- packed weights, so __call__ re-captures."""
        return tuple((t.data_ptr(), t._version) for t in list(self.model.parameters()) + list(self.model.buffers()))

    # ---- the sequence of kernel launches -------------------------------------------------------------------------
    def _run(self):
        B, K = self.B, self.K
        with torch.cuda.device(self.device), ops.precision(self.precision):
            z0 = self.ae._encode(Act.from_nchw(self.x_static))            # [B,h,w,Cz] fp32
            h, w, Cz = z0.H, z0.W, z0.C
            hwc = h * w * Cz
            n = B * K
            chw = self.C * self.Ly * self.Lx
            if self.zs is None:
                self.zs = torch.empty(B * K * hwc, dtype=torch.float32, device=self.device)
                self.h, self.w = h, w
                # pipelined decode (to_x): the latent stack is STEP-major [K][B][h][w][Cz]; as soon as S steps exist their
                # B*S samples are decoded on a second stream while the (latency-bound, small-launch) propagator goes on
                self.steps_per_group = self._steps_per_group(n) if self.to_x and self.pipeline else 0
                if self.to_x:
                    self.out = torch.empty(B, K, self.C, self.Ly, self.Lx, dtype=torch.float32, device=self.device)
            S = self.steps_per_group
            prepared = self._cond_cnn_prepare(self.prop, self.param_static) if self.conditional else None
            main = torch.cuda.current_stream(self.device)
            if S:
                if self._side is None:
                    self._side = torch.cuda.Stream(device=self.device)
                self._side.wait_stream(main)  # fork (inside a capture this makes the side stream part of the graph)
                out_flat = self.out.view(-1)
            z_in = z0
            for t in range(K):
                if S:
                    z_out = Act(self.zs[t * B * hwc:], B, h, w, Cz)                 # step t of every trajectory, contiguous
                else:
                    z_out = Act(self.zs[t * hwc:], B, h, w, Cz, bstride=K * hwc)    # slot t of every trajectory
                if self.conditional:
                    self._cond_cnn_fwd(self.prop, z_in, prepared, out=z_out)
                else:
                    self._simple_cnn_fwd(self.prop, z_in, out=z_out)
                z_in = z_out
                if S and ((t + 1) % S == 0 or t == K - 1):
                    t0 = (t // S) * S
                    ns = t + 1 - t0
                    done = torch.cuda.Event()
                    done.record(main)
                    self._side.wait_event(done)
                    with torch.cuda.stream(self._side):
                        zin = Act(self.zs[t0 * B * hwc:], ns * B, h, w, Cz)
                        # sample (step t0 + j, trajectory b) of the group -> slot [b][t0 + j] of the [B, K, C, Ly, Lx] result
                        dst = Act(out_flat[t0 * chw:], ns * B, self.Ly, self.Lx, self.C, bstride=K * chw, layout=ops.NCHW,
                                  group=B, gstride=chw)
                        self.ae._decode(zin, out=dst)
            if S:
                main.wait_stream(self._side)  # join
            elif self.to_x:
                out_flat = self.out.view(-1)
                chunk = self.decode_chunk or self._default_chunk(n)
                for c0 in range(0, n, chunk):
                    m = min(chunk, n - c0)
                    zin = Act(self.zs[c0 * hwc:], m, h, w, Cz)
                    dst = Act(out_flat[c0 * chw:], m, self.Ly, self.Lx, self.C, layout=ops.NCHW)
                    self.ae._decode(zin, out=dst)
            else:
                if self.out is None:
                    self.out = torch.empty(B, K, Cz, h, w, dtype=torch.float32, device=self.device)
                rc = ops._C.lib().lns_nhwc_to_nchw(ops._ptr(self.zs), ops.F32, n, h, w, Cz, hwc, ops._ptr(self.out),
                                                    hwc, ops._stream())
                ops.check(rc, "lns_nhwc_to_nchw")
                ops._state.launches += 1

    def _steps_per_group(self, n):
        """Rollout steps per pipelined decode group: about one default decode chunk, but at least four groups per rollout so
        that the decode of the early steps overlaps the propagator of the later ones."""
        chunk = self.decode_chunk or self._default_chunk(n)
        return max(1, min(int(round(chunk / self.B)), -(-self.K // 4)))

    def _default_chunk(self, n):
        """Samples per decode launch group: about 16 M output pixels per channel, split evenly over the B*K samples and
        rounded up to a multiple of the SM count -- the per-sample kernels (FABlock2D, SABlock: one CTA per sample) then
        run whole waves (measured on B200, NS2d 1184 x 20: 4096 -> 217.6k, 4736 = 32 x 148 -> 220.3k trajectory-steps/s)."""
        target = max(8, (16 << 20) // (self.Ly * self.Lx))
        nchunks = max(1, -(-n // int(1.2 * target)))
        chunk = -(-n // nchunks)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.device.type == "cuda" else 148
        if chunk >= sms:
            chunk = -(-chunk // sms) * sms
        return max(1, min(n, chunk))

    def build(self):
        if self._built:
            return self
        with torch.cuda.device(self.device):
            self._build()
        self._fp = self._fingerprint()
        self._built = True
        return self

    def _build(self):
        # eager warm-up: packs the filters, allocates the resident buffers, counts the launches of one rollout
        for p in self.ae.parameters():
            if p.device != self.device:
                raise LnsError("Rollout: all parameters must be on the rollout device")
        self._run()                      # first call also packs weights (extra launches)
        before = ops.launch_count()
        self._run()
        self.launches_per_call = ops.launch_count() - before
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._run()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._run()

    def __call__(self, x, param=None):
        if tuple(x.shape) != tuple(self.x_static.shape):
            raise LnsError(f"Rollout: expected x of shape {tuple(self.x_static.shape)}, got {tuple(x.shape)}")
        if self.conditional and param is None:
            raise LnsError("Rollout: this propagator is conditional, pass `param` [B]")
        if self._built and self._fingerprint() != self._fp:
            # a parameter was replaced or modified in place since the capture: re-pack and re-capture (buffers are kept)
            self._built, self.graph = False, None
        self.build()
        with torch.cuda.device(self.device):
            self.x_static.copy_(x, non_blocking=True)
            if self.conditional:
                self.param_static.copy_(param.reshape(-1), non_blocking=True)
            if self.graph is not None:
                self.graph.replay()
            else:
                self._run()
        return self.out

    def latents(self):
        """The resident latent stack as [B, K, h, w, Cz] fp32 (channel-last)."""
        if self.steps_per_group:
            return self.zs.view(self.K, self.B, self.h, self.w, self.Cz).transpose(0, 1)
        return self.zs.view(self.B, self.K, self.h, self.w, self.Cz)
