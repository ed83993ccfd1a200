"""The latent-space TRAINING rollout:

    LatentDynamics.forward(z_in, z_out, loss_fn)          train_stage2_ns2d.py:126-141, train_stage2_SW.py:126-142,
        z_pred = stack_t propagator^t(z_in);  loss_fn(z_pred, z_out)      train_stage2_twophase.py:126-142

i.e. back-propagation through t_out autoregressive steps of ``SimpleCNN`` (train_stage2_ns2d.py:25-87).  The whole rollout is
ONE autograd node: its forward runs the steps on the library's conv / norm kernels and keeps, per step, only what the backward
needs (block inputs, pre-activations and their GELU outputs, the folded GroupNorm affines); its backward walks the steps in reverse on the kernels of
csrc/backward.cu -- filter gradients (lns_conv2d_wgrad, accumulating over steps into one buffer per parameter), GroupNorm /
GELU backward -- and computes every DATA gradient as a forward convolution with the flipped, transposed filter on the same
engines as the forward pass.  The loss itself is the caller's torch function of z_pred (as in the reference).  No torch
arithmetic happens between the two ends; there is no fallback.

Precision: activations and gradients are fp32.  In the 'fp32' mode every GEMM runs on the CUDA-core engine (the validation path,
gradients within 2e-5 of fp64 autograd); in 'fp16s' the forward convolutions run on tcgen05 with split operands (three MMAs
per K step on fp32 storage, the mechanism of ops.hi_region), and so does the backward: data gradients on the same split-operand
tcgen05 engines, filter gradients on TF32 mma.sync with hi + lo split operands (lns_conv2d_wgrad, tensor_core = 1).
Gradients are 1e-4 ... 1e-7 in magnitude -- below the IEEE-half normal range, where the hi + lo split of an UNSCALED gradient
loses its low bits (measured: 1.6e-4 per layer instead of 3e-6) -- so that backward pass is LOSS-SCALED: the incoming gradient
is multiplied by a power of two S chosen from its max |.| (lns_absmax; target 64, three decades of headroom below the half
maximum), every gradient in flight carries S, and the accumulating kernels multiply by 1/S (exact).
The precision mode is thread-local and autograd runs backward on its own thread: the mode of the forward call is recorded
in the node and re-installed there."""
import math
import ctypes

import torch

from . import _C, ops
from .ops import Act, LnsError


def _mods():
    from modules import _base
    return _base


class _Tape:
    __slots__ = ("z", "blocks", "last")

    def __init__(self):
        self.blocks = []


def _gn(norm):
    return getattr(norm, "gn", norm)


def _affine(x, norm):
    g = _gn(norm)
    return ops.group_norm_affine(x, g.num_groups, g.eps, g.weight, g.bias)


_F32 = torch.float32


def _region():
    """fp32-stored activations on the tensor-core engines in the split-operand mode (every layer is a 'high-precision layer')."""
    return ops.hi_region(1 << 30)


def step_fwd(net, z, tape):
    """One SimpleCNN step on fp32 NHWC activations, recording the tape (train_stage2_ns2d.py:82-87, :50-53)."""
    B_ = _mods()
    cl = B_.conv_layer
    a = cl(z, net.in_proj, out_dtype=_F32)
    tape.z = z
    for blk in net.net:
        gn1, c1, _, c2, _, c3 = blk.conv
        s1, t1 = _affine(a, gn1)
        p1 = cl(a, c1, pro=(s1, t1, ops.ACT_NONE), out_dtype=_F32)
        # the GELU outputs are materialised once and kept: the filter gradient reads them nine times (once per tap), and
        # recomputing the erf there was a quarter of the wgrad kernel (ncu)
        h1 = ops.affine_act(p1, None, None, ops.ACT_GELU, out_dtype=_F32)
        p2 = cl(h1, c2, out_dtype=_F32)
        h2 = ops.affine_act(p2, None, None, ops.ACT_GELU, out_dtype=_F32)
        x2 = cl(h2, c3, residual=a, out_dtype=_F32)
        gn2, f1, _, f2 = blk.ffn
        s2, t2 = _affine(x2, gn2)
        q1 = cl(x2, f1, pro=(s2, t2, ops.ACT_NONE), out_dtype=_F32)
        hq = ops.affine_act(q1, None, None, ops.ACT_GELU, out_dtype=_F32)
        x3 = cl(hq, f2, residual=x2, out_dtype=_F32)
        tape.blocks.append((a, s1, t1, p1, h1, p2, h2, x2, s2, t2, q1, hq))
        a = x3
    so, to = _affine(a, net.out_proj[0])
    tape.last = (a, so, to)
    return cl(a, net.out_proj[1], pro=(so, to, ops.ACT_NONE), out_dtype=_F32)


def _dgrad_filter(conv):
    """PackedFilter of the adjoint convolution: W'[i][o][ky][kx] = W[o][i][KH-1-ky][KW-1-kx] (data movement only)."""
    B_ = _mods()
    c = B_.cache_of(conv)
    f = c.get("dgrad_filter")
    if f is None or c.get("dgrad_src") is not conv.weight:
        w = conv.weight
        Cout, Cin, KH, KW = w.shape
        f = ops.PackedFilter(lambda: w.detach().flip(2, 3).transpose(0, 1).contiguous(), None,
                             lambda: (ops.PackedFilter._fp(w),), (Cin, Cout, KH, KW))
        c["dgrad_filter"], c["dgrad_src"] = f, w
    return f


class _Bw:
    """Per-backward-pass settings: gradient buffers, tensor-core engines on / off, 1 / loss scale."""
    __slots__ = ("G", "tc", "inv")

    def __init__(self, G, tc, inv):
        self.G, self.tc, self.inv = G, tc, inv


def _dgrad(conv, dy, bw, residual=None):
    geo = _mods().conv_geometry(conv)
    # tensor-core mode: engine chosen by ops.conv2d inside the hi region (fp32 storage, split operands); else the exact engine
    return ops.conv2d(dy, _dgrad_filter(conv), use_bias=False, residual=residual, out_dtype=_F32,
                      engine=None if bw.tc else ops.ENGINE_SIMT, **geo)


def _wgrad(conv, x, pro, dy, bw):
    geo = _mods().conv_geometry(conv)
    gw = bw.G.get(id(conv.weight))
    if gw is not None:
        kh, kw = conv.kernel_size
        ops.conv2d_wgrad(x, dy, gw, KH=kh, KW=kw, dil=geo["dil"], pad=geo["pad"], pad_mode=geo["pad_mode"], pro=pro,
                         tensor_core=bw.tc, out_scale=bw.inv)
    if conv.bias is not None and id(conv.bias) in bw.G:
        ops.chan_sum_accum(dy, bw.G[id(conv.bias)], out_scale=bw.inv)


def _gn_bwd(norm, x, dy, dskip, bw):
    g = _gn(norm)
    return ops.group_norm_bwd(x, dy, g.num_groups, g.eps, g.weight, dskip=dskip,
                              dgamma=bw.G.get(id(g.weight)) if g.weight is not None else None,
                              dbeta=bw.G.get(id(g.bias)) if g.bias is not None else None, out_scale=bw.inv)


def step_bwd(net, tape, dzo, bw, extra=None):
    """Backward of step_fwd: dzo = gradient w.r.t. the step's output; parameter gradients are ACCUMULATED into bw.G[id(param)]
    (times bw.inv: dzo carries the loss scale);
    returns the gradient w.r.t. the step's input (+ extra, the loss gradient arriving at that latent directly)."""
    a, so, to = tape.last
    _wgrad(net.out_proj[1], a, (so, to, ops.ACT_NONE), dzo, bw)
    da = _gn_bwd(net.out_proj[0], a, _dgrad(net.out_proj[1], dzo, bw), None, bw)
    for blk, tp in zip(reversed(list(net.net)), reversed(tape.blocks)):
        x, s1, t1, p1, h1, p2, h2, x2, s2, t2, q1, hq = tp
        gn1, c1, _, c2, _, c3 = blk.conv
        gn2, f1, _, f2 = blk.ffn
        # x3 = x2 + f2(gelu(q1)),  q1 = f1(GN2(x2))
        _wgrad(f2, hq, None, da, bw)
        dq1 = ops.act_bwd(_dgrad(f2, da, bw), q1, ops.ACT_GELU)
        _wgrad(f1, x2, (s2, t2, ops.ACT_NONE), dq1, bw)
        dx2 = _gn_bwd(gn2, x2, _dgrad(f1, dq1, bw), da, bw)
        # x2 = x + c3(gelu(p2)),  p2 = c2(gelu(p1)),  p1 = c1(GN1(x))
        _wgrad(c3, h2, None, dx2, bw)
        dp2 = ops.act_bwd(_dgrad(c3, dx2, bw), p2, ops.ACT_GELU)
        _wgrad(c2, h1, None, dp2, bw)
        dp1 = ops.act_bwd(_dgrad(c2, dp2, bw), p1, ops.ACT_GELU)
        _wgrad(c1, x, (s1, t1, ops.ACT_NONE), dp1, bw)
        da = _gn_bwd(gn1, x, _dgrad(c1, dp1, bw), dx2, bw)
    _wgrad(net.in_proj, tape.z, None, da, bw)
    return _dgrad(net.in_proj, da, bw, residual=extra)


def _check_net(net):
    ok = (hasattr(net, "in_proj") and hasattr(net, "net") and hasattr(net, "out_proj") and not hasattr(net, "cond_emb_proj")
          and all(hasattr(b, "conv") and hasattr(b, "ffn") and len(b.conv) == 6 and len(b.ffn) == 4 for b in net.net))
    if not ok:
        raise NotImplementedError("the training rollout is implemented for the unconditional SimpleCNN propagator "
                                  "(train_stage2_ns2d.py / _SW.py / _twophase.py); the conditional variant has no backward yet")


class _RolloutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, z0, t_out, *params):
        ops._need_cuda(z0, "rollout_train")
        B, C, h, w = z0.shape
        z_pred = torch.empty(B, t_out, C, h, w, dtype=_F32, device=z0.device)
        tapes = []
        with ops.device_of(z0), _region():
            z = ops.nchw_to_act(z0.detach(), _F32)
            for t in range(t_out):
                tape = _Tape()
                z = step_fwd(net, z, tape)
                tapes.append(tape)
                dst = ctypes.c_void_p(z_pred.data_ptr() + t * C * h * w * 4)
                rc = _C.lib().lns_nhwc_to_nchw(ops._ptr(z.t), z.dtype, B, h, w, C, z.bstride, dst, t_out * C * h * w, ops._stream())
                _C.check(rc, "lns_nhwc_to_nchw")
                ops._state.launches += 1
        ctx.net, ctx.tapes, ctx.params, ctx.shape = net, tapes, params, (B, t_out, C, h, w)
        ctx.need_z0 = z0.requires_grad
        ctx.precision = ops.get_precision()
        return z_pred

    @staticmethod
    def backward(ctx, dz_pred):
        net, tapes, params = ctx.net, ctx.tapes, ctx.params
        B, T, C, h, w = ctx.shape
        dz_pred = dz_pred.contiguous().float()
        G = {id(p): torch.zeros_like(p, dtype=_F32, memory_format=torch.contiguous_format) for p in params if p.requires_grad}
        with ops.device_of(dz_pred), ops.precision(ctx.precision), _region():
            tc = ops.split16()
            S = 1.0
            if tc:
                amax = ops.absmax(dz_pred)
                if amax > 0.0 and math.isfinite(amax):
                    S = 2.0 ** max(-60, min(60, math.floor(math.log2(64.0 / amax))))
            bw = _Bw(G, tc, 1.0 / S)
            sc = torch.full((B * C,), S, dtype=_F32, device=dz_pred.device) if S != 1.0 else None
            zs = torch.zeros_like(sc) if sc is not None else None

            def loss_grad(t):
                out = Act.empty(B, h, w, C, _F32, dz_pred.device)
                src = ctypes.c_void_p(dz_pred.data_ptr() + t * C * h * w * 4)
                rc = _C.lib().lns_nchw_to_nhwc(src, B, C, h, w, T * C * h * w, ops._ptr(out.t), out.dtype, out.bstride, ops._stream())
                _C.check(rc, "lns_nchw_to_nhwc")
                ops._state.launches += 1
                return ops.affine_act(out, sc, zs, ops.ACT_NONE, out_dtype=_F32) if sc is not None else out

            dz = loss_grad(T - 1)
            for t in range(T - 1, -1, -1):
                dz = step_bwd(net, tapes[t], dz, bw, extra=loss_grad(t - 1) if t > 0 else None)
            dz0 = None
            if ctx.need_z0:
                if sc is not None:
                    sc.fill_(1.0 / S)
                    dz = ops.affine_act(dz, sc, zs, ops.ACT_NONE, out_dtype=_F32)
                dz0 = dz.to_nchw()
        ctx.tapes = None
        return (None, dz0, None) + tuple(G.get(id(p)) for p in params)


def rollout_train(net, z0, t_out):
    """z_pred [B, t_out, C, h, w] = the t_out autoregressive steps of `net` (a SimpleCNN) from z0 [B, C, h, w], differentiable
    w.r.t. the propagator's parameters (and z0)."""
    _check_net(net)
    if z0.dim() != 4:
        raise LnsError("rollout_train: z0 must be [B, C, h, w]")
    return _RolloutFn.apply(net, z0.contiguous().float(), int(t_out), *list(net.parameters()))
