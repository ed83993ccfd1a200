"""The latent-space TRAINING rollout:

    LatentDynamics.forward(z_in, z_out, loss_fn)          train_stage2_ns2d.py:126-141, train_stage2_SW.py:126-142,
        z_pred = stack_t propagator^t(z_in);  loss_fn(z_pred, z_out)      train_stage2_twophase.py:126-142

i.e. back-propagation through t_out autoregressive steps of ``SimpleCNN`` (train_stage2_ns2d.py:25-87), and the conditional
variant ``forward(z_in, z_out, param, loss_fn)`` (train_stage2_twophase_conditional.py:160-175, blocks :25-75): there the
step-invariant conditioning network (embedding MLP, per-block shift Linear and gate cond_conv2) runs once per rollout, the
gradients that reach its outputs are summed over the steps by lns_pixel_dot and back-propagated once.  The whole rollout is
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
is multiplied by a power of two S chosen ON THE DEVICE from its max |.| (lns_absmax + lns_loss_scale; target 64, three decades
of headroom below the half maximum), every gradient in flight carries S, and the accumulating kernels multiply by 1/S (exact).
No step of either pass synchronises with the host, so a whole training step can be captured in a CUDA graph (GraphedTrainStep).
The precision mode is thread-local and autograd runs backward on its own thread: the mode of the forward call is recorded
in the node and re-installed there."""
import math
import ctypes

import torch
import torch.nn as nn

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


def _geo(mod):
    """stride / dilation / padding of an nn.Conv2d holder; an nn.Linear is a 1x1 convolution on [B, 1, 1, C] rows"""
    if isinstance(mod, nn.Linear):
        return dict(stride=1, dil=1, pad=(0, 0, 0, 0), pad_mode=(ops.PAD_ZEROS, ops.PAD_ZEROS))
    return _mods().conv_geometry(mod)


def _fwd(x, mod, **kw):
    geo = _geo(mod)
    geo.update(kw)
    return ops.conv2d(x, _mods().filt_of(mod), out_dtype=_F32, **geo)


def _dgrad_filter(conv):
    """PackedFilter of the adjoint convolution: W'[i][o][ky][kx] = W[o][i][KH-1-ky][KW-1-kx] (data movement only)."""
    B_ = _mods()
    c = B_.cache_of(conv)
    f = c.get("dgrad_filter")
    if f is None or c.get("dgrad_src") is not conv.weight:
        w = conv.weight
        if w.dim() == 2:
            Cout, Cin = w.shape
            f = ops.PackedFilter(lambda: w.detach().t()[:, :, None, None].contiguous(), None,
                                 lambda: (ops.PackedFilter._fp(w),), (Cin, Cout, 1, 1))
        else:
            Cout, Cin, KH, KW = w.shape
            f = ops.PackedFilter(lambda: w.detach().flip(2, 3).transpose(0, 1).contiguous(), None,
                                 lambda: (ops.PackedFilter._fp(w),), (Cin, Cout, KH, KW))
        c["dgrad_filter"], c["dgrad_src"] = f, w
    return f


class _Bw:
    """Per-backward-pass settings: gradient buffers, tensor-core engines on / off, 1 / loss scale (a one-element device tensor
    or None: the scale is chosen on the device, so the pass never synchronises with the host)."""
    __slots__ = ("G", "tc", "inv")

    def __init__(self, G, tc, inv):
        self.G, self.tc, self.inv = G, tc, inv


def _dgrad(conv, dy, bw, residual=None):
    geo = _geo(conv)
    # tensor-core mode: engine chosen by ops.conv2d inside the hi region (fp32 storage, split operands); else the exact engine
    return ops.conv2d(dy, _dgrad_filter(conv), use_bias=False, residual=residual, out_dtype=_F32,
                      engine=None if bw.tc else ops.ENGINE_SIMT, **geo)


def _wgrad(conv, x, pro, dy, bw):
    geo = _geo(conv)
    gw = bw.G.get(id(conv.weight))
    if gw is not None:
        kh, kw = (1, 1) if isinstance(conv, nn.Linear) else conv.kernel_size
        ops.conv2d_wgrad(x, dy, gw.view(gw.shape[0], gw.shape[1], kh, kw), KH=kh, KW=kw, dil=geo["dil"], pad=geo["pad"],
                         pad_mode=geo["pad_mode"], pro=pro, tensor_core=bw.tc, out_scale_dev=bw.inv)
    if conv.bias is not None and id(conv.bias) in bw.G:
        ops.chan_sum_accum(dy, bw.G[id(conv.bias)], out_scale_dev=bw.inv)


def _gn_bwd(norm, x, dy, dskip, bw):
    g = _gn(norm)
    return ops.group_norm_bwd(x, dy, g.num_groups, g.eps, g.weight, dskip=dskip,
                              dgamma=bw.G.get(id(g.weight)) if g.weight is not None else None,
                              dbeta=bw.G.get(id(g.bias)) if g.bias is not None else None, out_scale_dev=bw.inv)


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



# ---- conditional propagator (train_stage2_twophase_conditional.py:25-121) ---------------------------------------------------
def _gelu(x):
    return ops.affine_act(x, None, None, ops.ACT_GELU, out_dtype=_F32)


class _CondTape:
    """Step-invariant conditioning network of one rollout: cond = MLP(fourier_embedding(param)); per block shift =
    Linear(cond) and gate = cond_conv2(shift) (reference :66-75, :114-116), with what its backward needs."""
    __slots__ = ("emb", "e1pre", "e1", "cond", "blocks")


def cond_prepare_fwd(net, param):
    from modules.cond_utils import fourier_embedding
    ct = _CondTape()
    ct.emb = ops.rows_act(fourier_embedding(param, dim=net.cond_emb_dim))
    l1, _, l2 = net.cond_emb_proj
    ct.e1pre = _fwd(ct.emb, l1)
    ct.e1 = _gelu(ct.e1pre)
    ct.cond = _fwd(ct.e1, l2)
    ct.blocks = []
    for blk in net.net:
        shift = _fwd(ct.cond, blk.cond_emb)                                     # [B,1,1,C]
        gn, c1, _, c2 = blk.cond_conv2
        sg, tg = _affine(shift, gn)
        g1pre = _fwd(shift, c1, pro=(sg, tg, ops.ACT_NONE))
        g1 = _gelu(g1pre)
        gate = _fwd(g1, c2)
        ones = torch.ones_like(gate.t)
        onep = ops.affine_act(gate, ones, ones, ops.ACT_NONE, out_dtype=_F32)   # 1 + gate
        ct.blocks.append((shift, sg, tg, g1pre, g1, onep))
    return ct


def cond_step_fwd(net, z, ct, tape):
    """One conditional SimpleCNN step (reference :66-75 per block, :117-121), recording the tape."""
    a = _fwd(z, net.in_proj)
    tape.z = z
    for blk, cb in zip(net.net, ct.blocks):
        shift, onep = cb[0], cb[5]
        gn1, c1, _, c2 = blk.conv1
        s1, t1 = _affine(a, gn1)
        p1 = _fwd(a, c1, pro=(s1, t1, ops.ACT_NONE))
        h1 = _gelu(p1)
        p2 = _fwd(h1, c2, sample_bias=shift.t)
        gnb, _, cz = blk.cond_conv1
        sb, tb = _affine(p2, gnb)
        nb = ops.affine_act(p2, sb, tb, ops.ACT_NONE, out_dtype=_F32)
        hb = _gelu(nb)
        x2 = _fwd(hb, cz, residual=a)
        hin = ops.scale_add(x2, scale=onep.t)                                   # x2 * (1 + gate)
        gn2, f1, _, f2 = blk.ffn
        s2, t2 = _affine(hin, gn2)
        q1 = _fwd(hin, f1, pro=(s2, t2, ops.ACT_NONE))
        hq = _gelu(q1)
        x3 = _fwd(hq, f2, residual=x2)
        tape.blocks.append((a, s1, t1, p1, h1, p2, nb, hb, x2, hin, s2, t2, q1, hq))
        a = x3
    so, to = _affine(a, net.out_proj[0])
    tape.last = (a, so, to)
    return _fwd(a, net.out_proj[1], pro=(so, to, ops.ACT_NONE))


def cond_step_bwd(net, ct, tape, dzo, bw, acc, extra=None):
    """Backward of cond_step_fwd; acc[i] = (dshift [B*C], dgate [B*C]) accumulate the gradients that reach block i's
    conditioning vectors (summed over the rollout steps, back-propagated once by cond_prepare_bwd)."""
    a, so, to = tape.last
    _wgrad(net.out_proj[1], a, (so, to, ops.ACT_NONE), dzo, bw)
    da = _gn_bwd(net.out_proj[0], a, _dgrad(net.out_proj[1], dzo, bw), None, bw)
    for i in range(len(net.net) - 1, -1, -1):
        blk, cb = net.net[i], ct.blocks[i]
        onep = cb[5]
        x, s1, t1, p1, h1, p2, nb, hb, x2, hin, s2, t2, q1, hq = tape.blocks[i]
        gn1, c1, _, c2 = blk.conv1
        gnb, _, cz = blk.cond_conv1
        gn2, f1, _, f2 = blk.ffn
        # x3 = x2 + f2(gelu(q1)),  q1 = f1(GN2(x2 * (1 + gate)))
        _wgrad(f2, hq, None, da, bw)
        dq1 = ops.act_bwd(_dgrad(f2, da, bw), q1, ops.ACT_GELU)
        _wgrad(f1, hin, (s2, t2, ops.ACT_NONE), dq1, bw)
        dhin = _gn_bwd(gn2, hin, _dgrad(f1, dq1, bw), None, bw)
        ops.pixel_dot(dhin, x2, acc[i][1])                                      # d (1 + gate)
        dx2 = ops.scale_add(dhin, scale=onep.t, skip=da)
        # x2 = x + cz(gelu(GNb(p2))),  p2 = c2(gelu(p1)) + shift,  p1 = c1(GN1(x))
        _wgrad(cz, hb, None, dx2, bw)
        dnb = ops.act_bwd(_dgrad(cz, dx2, bw), nb, ops.ACT_GELU)
        dp2 = _gn_bwd(gnb, p2, dnb, None, bw)
        ops.pixel_dot(dp2, None, acc[i][0])                                     # d shift
        _wgrad(c2, h1, None, dp2, bw)
        dp1 = ops.act_bwd(_dgrad(c2, dp2, bw), p1, ops.ACT_GELU)
        _wgrad(c1, x, (s1, t1, ops.ACT_NONE), dp1, bw)
        da = _gn_bwd(gn1, x, _dgrad(c1, dp1, bw), dx2, bw)
    _wgrad(net.in_proj, tape.z, None, da, bw)
    return _dgrad(net.in_proj, da, bw, residual=extra)


def cond_prepare_bwd(net, ct, acc, bw):
    """Back-propagate the accumulated shift / gate gradients through cond_conv2, the per-block Linear and the embedding MLP."""
    B = ct.cond.B
    dcond = None
    for i in range(len(net.net) - 1, -1, -1):
        blk = net.net[i]
        shift, sg, tg, g1pre, g1, _ = ct.blocks[i]
        C = shift.C
        dshift_acc = Act(acc[i][0], B, 1, 1, C)
        dgate = Act(acc[i][1], B, 1, 1, C)
        gn, c1, _, c2 = blk.cond_conv2
        _wgrad(c2, g1, None, dgate, bw)
        dg1pre = ops.act_bwd(_dgrad(c2, dgate, bw), g1pre, ops.ACT_GELU)
        _wgrad(c1, shift, (sg, tg, ops.ACT_NONE), dg1pre, bw)
        dshift = _gn_bwd(gn, shift, _dgrad(c1, dg1pre, bw), dshift_acc, bw)
        _wgrad(blk.cond_emb, ct.cond, None, dshift, bw)
        dcond = _dgrad(blk.cond_emb, dshift, bw, residual=dcond)
    l1, _, l2 = net.cond_emb_proj
    _wgrad(l2, ct.e1, None, dcond, bw)
    de1pre = ops.act_bwd(_dgrad(l2, dcond, bw), ct.e1pre, ops.ACT_GELU)
    _wgrad(l1, ct.emb, None, de1pre, bw)


def _check_net(net):
    """-> True for the conditional propagator, False for the unconditional one; raises for anything else."""
    base = hasattr(net, "in_proj") and hasattr(net, "net") and hasattr(net, "out_proj")
    if base and hasattr(net, "cond_emb_proj"):
        if all(hasattr(b, "conv1") and hasattr(b, "cond_conv1") and hasattr(b, "cond_conv2") and hasattr(b, "cond_emb")
               and hasattr(b, "ffn") for b in net.net):
            return True
    elif base and all(hasattr(b, "conv") and hasattr(b, "ffn") and len(b.conv) == 6 and len(b.ffn) == 4 for b in net.net):
        return False
    raise NotImplementedError("the training rollout is implemented for the SimpleCNN propagators of the stage-2 scripts "
                              "(train_stage2_ns2d.py / _SW.py / _twophase.py and the conditional _twophase_conditional.py)")


class _RolloutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, z0, t_out, param, *params):
        ops._need_cuda(z0, "rollout_train")
        B, C, h, w = z0.shape
        z_pred = torch.empty(B, t_out, C, h, w, dtype=_F32, device=z0.device)
        tapes = []
        with ops.device_of(z0), _region():
            z = ops.nchw_to_act(z0.detach(), _F32)
            ct = cond_prepare_fwd(net, param.detach().float().contiguous()) if param is not None else None
            for t in range(t_out):
                tape = _Tape()
                z = cond_step_fwd(net, z, ct, tape) if ct is not None else step_fwd(net, z, tape)
                tapes.append(tape)
                dst = ctypes.c_void_p(z_pred.data_ptr() + t * C * h * w * 4)
                rc = _C.lib().lns_nhwc_to_nchw(ops._ptr(z.t), z.dtype, B, h, w, C, z.bstride, dst, t_out * C * h * w, ops._stream())
                _C.check(rc, "lns_nhwc_to_nchw")
                ops._state.launches += 1
        ctx.net, ctx.tapes, ctx.params, ctx.shape, ctx.ct = net, tapes, params, (B, t_out, C, h, w), ct
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
            s2 = ops.loss_scale(dz_pred) if tc else None     # device tensor (S, 1 / S): no host round trip, graph-capturable
            bw = _Bw(G, tc, s2[1:] if tc else None)

            def loss_grad(t):
                out = Act.empty(B, h, w, C, _F32, dz_pred.device)
                src = ctypes.c_void_p(dz_pred.data_ptr() + t * C * h * w * 4)
                rc = _C.lib().lns_nchw_to_nhwc(src, B, C, h, w, T * C * h * w, ops._ptr(out.t), out.dtype, out.bstride, ops._stream())
                _C.check(rc, "lns_nchw_to_nhwc")
                ops._state.launches += 1
                return ops.scale_by(out, s2[0:1]) if tc else out

            ct = ctx.ct
            acc = None
            if ct is not None:
                Cp = ct.blocks[0][0].C
                acc = [(torch.zeros(B * Cp, dtype=_F32, device=dz_pred.device), torch.zeros(B * Cp, dtype=_F32, device=dz_pred.device))
                       for _ in net.net]
            dz = loss_grad(T - 1)
            for t in range(T - 1, -1, -1):
                ex = loss_grad(t - 1) if t > 0 else None
                dz = cond_step_bwd(net, ct, tapes[t], dz, bw, acc, extra=ex) if ct is not None else step_bwd(net, tapes[t], dz, bw, extra=ex)
            if ct is not None:
                cond_prepare_bwd(net, ct, acc, bw)
            dz0 = None
            if ctx.need_z0:
                dz0 = (ops.scale_by(dz, s2[1:]) if tc else dz).to_nchw()
        ctx.tapes = ctx.ct = None
        return (None, dz0, None, None) + tuple(G.get(id(p)) for p in params)


def rollout_train(net, z0, t_out, param=None):
    """z_pred [B, t_out, C, h, w] = the t_out autoregressive steps of `net` (a SimpleCNN of the stage-2 scripts; the conditional
    one takes `param` [B]) from z0 [B, C, h, w], differentiable w.r.t. the propagator's parameters (and z0)."""
    cond = _check_net(net)
    if cond != (param is not None):
        raise LnsError("rollout_train: `param` goes with the conditional propagator (and only with it)")
    if z0.dim() != 4:
        raise LnsError("rollout_train: z0 must be [B, C, h, w]")
    return _RolloutFn.apply(net, z0.contiguous().float(), int(t_out), param, *list(net.parameters()))


class GraphedTrainStep:
    """One training step -- LatentDynamics.forward(z_in, z_out[, param], loss_fn) + backward [+ optimizer.step()] -- captured once
    as a CUDA graph and replayed: at the reference's training shape (configs/ns2d_stage2_prop.yml: batch 32, out_tw 2) the eager
    step is ~400 launches of 5-40 us kernels, i.e. launch bound.  Inputs are copied into static buffers; `.grad` tensors are
    static (zeroed inside the graph).  `optimizer`: a capturable torch optimizer (e.g. AdamW(..., capturable=True)) to put
    `step()` inside the graph, or None to leave it to the caller."""

    def __init__(self, model, z_in, z_out, loss_fn, param=None, optimizer=None, precision=None, warmup=3):
        self.model, self.loss_fn, self.optimizer = model, loss_fn, optimizer
        self.precision = precision or ops.get_precision()
        self.z_in, self.z_out = z_in.clone(), z_out.clone()
        self.param = param.clone() if param is not None else None
        self.params = [p for p in model.propagator.parameters() if p.requires_grad]
        side = torch.cuda.Stream(z_in.device)
        side.wait_stream(torch.cuda.current_stream(z_in.device))
        with torch.cuda.stream(side):
            for _ in range(warmup):       # warm-up off the capture stream: packs filters, sizes the allocator pools, creates .grad
                self._one()
        torch.cuda.current_stream(z_in.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._one()
        # the filter images cached during capture live in the graph's pool and hold data only while a replay is running:
        # eager callers (predict, an eager training step) must not find them
        ops.invalidate_packed()

    def _one(self):
        ops.invalidate_packed()   # the filter re-layout kernels must be part of the captured sequence: weights change between replays
        for p in self.params:
            if p.grad is not None:
                p.grad.zero_()
        with ops.precision(self.precision):
            args = (self.z_in, self.z_out) + ((self.param,) if self.param is not None else ()) + (self.loss_fn,)
            loss = self.model(*args)
            loss.backward()
        if self.optimizer is not None:
            self.optimizer.step()
        return loss

    def __call__(self, z_in, z_out, param=None):
        self.z_in.copy_(z_in, non_blocking=True)
        self.z_out.copy_(z_out, non_blocking=True)
        if self.param is not None:
            self.param.copy_(param, non_blocking=True)
        self.graph.replay()
        return self.loss
