"""Common plumbing of the drop-in modules."""
import weakref

import torch
import torch.nn as nn

from lns_b200 import ops
from lns_b200.ops import Act, LnsError


def pad_modes(padding_mode):
    """torch padding_mode string -> (mode_h, mode_w)"""
    if padding_mode == "circular":
        return (ops.PAD_CIRCULAR, ops.PAD_CIRCULAR)
    if padding_mode == "zeros":
        return (ops.PAD_ZEROS, ops.PAD_ZEROS)
    raise LnsError(f"unsupported padding_mode {padding_mode!r}")


_CACHES = weakref.WeakKeyDictionary()


def cache_of(mod):
    """Per-module dictionary of device-side derived data (packed filter images, 16-bit operand copies, rotary tables).  It lives
    OUTSIDE the module: copy.deepcopy / pickling / torch.save of a model never drag packed images -- or closures over the
    ORIGINAL module's parameters -- into the copy, which starts with an empty cache of its own."""
    c = _CACHES.get(mod)
    if c is None:
        c = {}
        _CACHES[mod] = c
    return c


def filt_of(mod):
    """PackedFilter of an nn.Conv2d / nn.Linear parameter holder (cached per module object and per parameter object)."""
    c = cache_of(mod)
    f = c.get("filter")
    if f is None or c.get("filter_src") is not mod.weight:  # (the parameter object itself was replaced: new PackedFilter)
        f = ops.PackedFilter.of(mod.weight, mod.bias)
        c["filter"], c["filter_src"] = f, mod.weight
    return f


class LnsModule(nn.Module):
    """nn.Module whose forward takes / returns the reference's NCHW fp32 tensors and whose ``_fwd`` maps an NHWC
    ``Act`` to an ``Act`` with library kernels.  Containers call children's ``_fwd`` directly, so a whole network
    stays channel-last between its two ends."""

    def _fwd(self, x, *args):  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, x, *args):
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            raise LnsError("lns_b200 modules are inference-only (no backward kernels yet): call under torch.no_grad()")
        with ops.device_of(x):
            y = self._fwd(ops.nchw_to_act(x), *args)
            return y.to_nchw()


def run_sequential(layers, x):
    for layer in layers:
        x = layer._fwd(x)
    return x


# ---- helpers shared by the block implementations ---------------------------------------------------------------
def conv_geometry(conv):
    """stride / dilation / padding of an nn.Conv2d holder, including HalfPeriodicConv2d
    (reference: modules/autoencoder2d_half_periodic.py:26-52 -- its nn.Conv2d padding is always 0, the real padding
    `padding_` is applied by hand: circular along the periodic direction, zeros along the other)."""
    k = conv.kernel_size[0]
    if hasattr(conv, "periodic_direction"):
        p = conv.padding_
        if conv.periodic_direction == "x":
            modes = (ops.PAD_ZEROS, ops.PAD_CIRCULAR)
        elif conv.periodic_direction == "y":
            modes = (ops.PAD_CIRCULAR, ops.PAD_ZEROS)
        else:
            raise ValueError("periodic_direction must be x or y")
        pad = (p, p, p, p)
    else:
        p = conv.padding if isinstance(conv.padding, tuple) else (conv.padding, conv.padding)
        pad = (p[0], p[0], p[1], p[1])
        modes = pad_modes(conv.padding_mode) if k > 1 else (ops.PAD_ZEROS, ops.PAD_ZEROS)
    return dict(stride=conv.stride[0], dil=conv.dilation[0], pad=pad, pad_mode=modes)


def conv_layer(x, conv, **kw):
    geo = conv_geometry(conv)
    geo.update(kw)
    return ops.conv2d(x, filt_of(conv), **geo)


def lazy_norm(x, norm, act=ops.ACT_NONE, prescale=None):
    """Pending GroupNorm-family normalisation (+ activation) of x, to be consumed by the next conv (ops.LazyNorm)."""
    gn = getattr(norm, "gn", norm)
    if isinstance(gn, nn.GroupNorm):
        return ops.LazyNorm(x, gn.num_groups, gn.eps, gn.weight, gn.bias, prescale, act)
    if isinstance(gn, nn.InstanceNorm2d):
        return ops.LazyNorm(x, x.C, gn.eps, gn.weight, gn.bias, prescale, act)
    raise LnsError(f"unsupported norm {type(norm).__name__}")


def norm_affine(x, norm, prescale=None):
    """(scale, shift) of a GroupNorm-family module applied to x: basics.GroupNorm (wrapper with .gn), nn.GroupNorm,
    nn.InstanceNorm2d (no affine, per-channel groups)."""
    gn = getattr(norm, "gn", norm)
    if isinstance(gn, nn.GroupNorm):
        return ops.group_norm_affine(x, gn.num_groups, gn.eps, gn.weight, gn.bias, prescale)
    if isinstance(gn, nn.InstanceNorm2d):
        return ops.group_norm_affine(x, x.C, gn.eps, gn.weight, gn.bias, prescale)
    raise LnsError(f"unsupported norm {type(norm).__name__}")


_ACT_OF = {"Swish": ops.ACT_SILU, "SiLU": ops.ACT_SILU, "GELU": ops.ACT_GELU}


def act_code(layer):
    name = type(layer).__name__
    if name == "GELU" and getattr(layer, "approximate", "none") != "none":
        raise LnsError("only exact-erf GELU is implemented")
    return _ACT_OF.get(name)


def _is_norm(layer):
    return isinstance(getattr(layer, "gn", layer), (nn.GroupNorm, nn.InstanceNorm2d))


def run_layers(layers, x, final_out=None, final_layout=ops.NHWC, final_dtype=None):
    """Execute an nn.Sequential-style list with peephole fusion:
       norm [+ act] -> conv      : statistics kernel + affine folded into the conv's gather prologue
       conv -> act               : activation in the conv epilogue
       nn.Upsample(size) -> conv : nearest resize folded into the conv's index map (never materialised)
    Blocks (anything with ``_fwd``) are called as they are.  ``final_out`` / ``final_layout`` apply to the last conv
    (lets the decoder's output projection write straight into the caller's [B,K,C,H,W] buffer)."""
    layers = list(layers)
    pro = None    # pending (scale, shift, act)
    virt = None   # pending nearest-resize target
    i, n = 0, len(layers)
    while i < n:
        layer = layers[i]
        nxt = layers[i + 1] if i + 1 < n else None
        if isinstance(layer, nn.Conv2d):
            kw = {}
            # conv kxk directly followed by a conv 1x1 (no activation between): one conv with the composed filter
            # (bf16 path only; the fp32 validation path keeps the reference's two steps)
            if (ops.fast16() and isinstance(nxt, nn.Conv2d) and nxt.kernel_size == (1, 1)
                    and layer.kernel_size[0] > 1 and nxt.stride == (1, 1) and i + 1 < n - 1
                    and not hasattr(nxt, "periodic_direction")):
                c = cache_of(layer)
                f1, f2 = filt_of(layer), filt_of(nxt)
                ent = c.get("composed")
                if ent is None or ent[0] is not f1 or ent[1] is not f2:
                    ent = (f1, f2, ops.composed_filter(f1, f2))
                    c["composed"] = ent
                filt = ent[2]
                geo = conv_geometry(layer)
                x = ops.conv2d(x, filt, pro=pro, virt=virt, **geo)
                pro, virt = None, None
                i += 2
                continue
            a = act_code(nxt) if nxt is not None else None
            if a is not None:
                kw["act"] = a
                i += 1
            if i == n - 1:
                if final_out is not None:
                    kw["out"] = final_out
                elif final_dtype is not None:
                    kw["out_dtype"] = final_dtype
                kw["out_layout"] = final_layout
            x = conv_layer(x, layer, pro=pro, virt=virt, **kw)
            pro, virt = None, None
        elif _is_norm(layer):
            if pro is not None or virt is not None:
                raise LnsError("run_layers: norm after a pending norm/resize is not supported")
            a = act_code(nxt) if nxt is not None else None
            if a is not None:
                i += 1
            pro = lazy_norm(x, layer, a or ops.ACT_NONE)
        elif isinstance(layer, nn.Upsample):
            if layer.mode != "nearest" or layer.size is None:
                raise LnsError("only nn.Upsample(size=..., mode='nearest') is implemented")
            if pro is not None:
                x = _flush(x, pro)
                pro = None
            virt = tuple(layer.size)
        elif act_code(layer) is not None:
            if virt is not None:
                raise LnsError("run_layers: activation after a pending resize is not supported")
            if pro is not None:
                x = _flush(x, pro)
                pro = None
            pro = (None, None, act_code(layer))
        elif hasattr(layer, "_fwd"):
            if virt is not None:
                raise LnsError("run_layers: resize must be followed by a conv")
            if pro is not None:
                x = _flush(x, pro)
                pro = None
            x = layer._fwd(x)
        elif isinstance(layer, nn.Identity):
            pass
        else:
            raise LnsError(f"run_layers: unsupported layer {type(layer).__name__}")
        i += 1
    if virt is not None:
        raise LnsError("run_layers: trailing resize")
    if pro is not None:
        x = _flush(x, pro)
    return x


def _flush(x, pro):
    return pro.materialize() if isinstance(pro, ops.LazyNorm) else ops.affine_act(x, *pro)


def latent_dtype(channels):
    """Storage type of latent-sized tensors: fp32 unless they can feed the tensor-core engine directly (C % 64 == 0);
    always fp32 in the split-operand mode, whose coarse stages keep fp32 storage."""
    if ops.split16():
        return torch.float32
    return ops.act_dtype() if channels % 64 == 0 else torch.float32


def decoder_hi_px(z):
    """'fp16s': the decoder's high-precision region = its two coarsest levels (the latent grid and the first x2 level), where
    tools/precision_study.py / tools/gpu_precision_attrib.py locate 70-80 % of the 16-bit error of a decode."""
    return 4 * z.H * z.W


def encoder_hi_px(layers, x):
    """Same region seen from the encoder: the grids after its last two down-sampling blocks."""
    n = sum(1 for m in layers if type(m).__name__.startswith("DownSample"))
    k = max(n - 1, 0)
    return (x.H >> k) * (x.W >> k)
