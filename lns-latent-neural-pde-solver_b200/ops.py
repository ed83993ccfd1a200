"""Thin Python wrappers over the C-ABI kernels: activation handles (NHWC), weight packing caches and one function
per library entry point.  PyTorch is used for device memory and streams only; all arithmetic on the path happens in
liblns_b200.so.  Nothing here falls back to torch ops: a CPU tensor or a missing library raises."""
import ctypes
import math
import os
import threading
from contextlib import contextmanager

import torch

from . import _C
from ._C import ConvDesc, LnsError, check

F32, BF16, TF32_CODE, F16 = 0, 1, 2, 3
TF32 = "tf32"  # storage sentinel: fp32 words holding TF32-rounded values (LNS_TF32); accepted wherever a dtype is
NHWC, NCHW = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU = 0, 1, 2
PAD_ZEROS, PAD_CIRCULAR = 0, 1
W_SIMT_F32, W_UMMA_BF16, W_UMMA_TF32, W_UMMA_F16, W_UMMA_F16X2 = 0, 1, 2, 3, 4
ENGINE_SIMT, ENGINE_UMMA, ENGINE_HALO, ENGINE_LATENT, ENGINE_COARSE = 0, 1, 2, 3, 4

_TORCH_DT = {F32: torch.float32, BF16: torch.bfloat16, F16: torch.float16}
H16_DTYPES = (torch.bfloat16, torch.float16)  # the 16-bit storage types of the tensor-core paths


def dt_code(dtype):
    if dtype == torch.float32:
        return F32
    if dtype == torch.bfloat16:
        return BF16
    if dtype == torch.float16:
        return F16
    raise LnsError(f"unsupported dtype {dtype}")


# ---- precision policy --------------------------------------------------------------------------------
class _State(threading.local):
    def __init__(self):
        self.precision = "fp16s"  # the 16-bit tensor-core mode that meets the 2e-3 per-stage bound (set_precision below)
        self.launches = 0
        self.timeline = None  # list of (label, start_event, end_event) when profiling (tools/timeline.py)
        self.hi_px = 0        # 'fp16s': convs touching a grid of <= hi_px pixels run with split operands on fp32 storage
        self.hi_exact = int(os.environ.get("LNS_HI_EXACT", "0"))  # experiment: those convs on the fp32 CUDA-core engine instead
        self.wsplit = False   # 'fp16s': split filters (2 MMAs per K step) on the tcgen05 convs outside the hi region too
        self.wsplit_policy = os.environ.get("LNS_WSPLIT", "enc")  # where ops.wsplit_region turns that on: enc | all | none
        self.hi_scale = float(os.environ.get("LNS_HI_SCALE", "1"))  # experiment: 4 = one more resolution level in the hi region
        self.hi_wsplit = os.environ.get("LNS_HI_WSPLIT", "1") != "0"  # hi layers also split the filter (3 MMAs instead of 2)
        # experiment: split the filter only on the COARSEST level of the region (grids of <= hi_px / 4 pixels)
        self.hi_wsplit_coarsest = os.environ.get("LNS_HI_WSPLIT", "1") == "coarsest"
        self.coarse = os.environ.get("LNS_COARSE", "1") != "0"        # use the block-halo engine (conv_coarse.cu) where it applies
        self.coarse_pro = os.environ.get("LNS_COARSE_PRO", "1") != "0"  # ... and let its fp32 producer apply the pending norm + act
        self.coarse_stats = os.environ.get("LNS_COARSE_STATS", "1") != "0"  # ... and its epilogue emit the next GroupNorm's statistics
        # FABlock2D with every contraction on tcgen05 (fablock_tc.cu): correct and tested, but its per-head chain of eight
        # barrier-separated MMA / drain stages is latency bound at one CTA per SM -- 7.7 ms vs 5.2 ms for the mma.sync kernel at
        # 4736 x 32x32 (profiles/r02_fablock_tc.md) -- so it is opt-in until the stages are software-pipelined
        self.fablock_tc = os.environ.get("LNS_FABLOCK_TC", "0") != "0"
        # FABlock2D whole-block kernel on pre-staged operands with a producer warp (fablock_full2_kernel); LNS_FABLOCK_STAGED=0
        # falls back to the in-kernel staging version (fablock_full_kernel)
        self.fablock_staged = os.environ.get("LNS_FABLOCK_STAGED", "1") != "0"


def _mark(label, flops=0.0, nbytes=0.0):
    """Profiling hook: returns a closer that records a CUDA-event pair around one library call (eager mode only), together with
    the call's ALGORITHMIC flops (2 x MACs of the reference layer) and bytes (compulsory reads + writes) -- bench.py's roofline."""
    tl = _state.timeline
    if tl is None:
        return None
    e0 = torch.cuda.Event(enable_timing=True)
    e0.record()
    return (label, e0, float(flops), float(nbytes))


def _done(tok):
    if tok is not None:
        e1 = torch.cuda.Event(enable_timing=True)
        e1.record()
        _state.timeline.append((tok[0], tok[1], e1, tok[2], tok[3]))


def _abytes(*acts):
    """bytes of the given activations (None entries are skipped)"""
    return float(sum(a.B * a.H * a.W * a.C * a.t.element_size() for a in acts if a is not None))


_state = _State()


def get_precision():
    return _state.precision


def set_precision(p):
    """'bf16': bf16 activations, tcgen05 tensor-core GEMMs with fp32 accumulation (the fast path).
    'fp16': IEEE-half activations and GEMM operands through the SAME kernels, bytes and tensor-core rate as 'bf16'
            (tcgen05.mma.kind::f16 takes either format) with an 11-bit significand: TF32-class rounding error, i.e. the
            16-bit path that meets the 2e-3 per-step bound.  Conversions saturate at +-65504 (the residual stream of a
            trained model must stay inside the half range; norm statistics, softmax and latents are fp32 as in 'bf16').
    'fp16s': the fp16 path with SPLIT operands where the rounding error is made (tools/precision_study.py): filters as
            hi + lo IEEE-half pairs everywhere (two MMAs per K step, no filter rounding), and the coarse stages of the
            autoencoder (ops.hi_region: grids of at most 4x the latent's pixels) on fp32 storage with the activation split
            into hi + lo halves inside the conv's producer (three MMAs: fp32-class layers on the f16 tensor path).  This is
            the 16-bit tensor-core mode that meets <= 2e-3 per stage on all four configurations.
    'tf32': fp32 storage rounded to TF32 at every write, tcgen05.mma.kind::tf32 GEMMs with fp32 accumulation (the
            tensor-core path that meets the 2e-3 per-step bound; half the MMA rate and twice the bytes of bf16).
    'fp32': fp32 activations and CUDA-core fp32 FMA GEMMs (the validation path, <=1e-5 vs the reference)."""
    if p not in ("bf16", "fp16", "fp16s", "fp32", "tf32"):
        raise ValueError("precision must be 'bf16', 'fp16', 'fp16s', 'tf32' or 'fp32'")
    _state.precision = p


@contextmanager
def precision(p):
    old = _state.precision
    set_precision(p)
    try:
        yield
    finally:
        _state.precision = old


def act_dtype():
    if _state.precision == "bf16":
        return torch.bfloat16
    if _state.precision in ("fp16", "fp16s"):
        return torch.float16
    return TF32 if _state.precision == "tf32" else torch.float32


def fast16():
    """True on the 16-bit tensor-core paths ('bf16' / 'fp16' / 'fp16s'), which share every kernel and fusion decision."""
    return _state.precision in ("bf16", "fp16", "fp16s")


def split16():
    """True in the split-operand mode 'fp16s'."""
    return _state.precision == "fp16s"


@contextmanager
def wsplit_region(kind):
    """Inside: ('fp16s' only) every tcgen05 conv uses the split filter.  kind = 'enc' (the encoder: it runs once per rollout and
    its error is carried by the fine levels, where only the filter split is affordable) or 'dec'."""
    old = _state.wsplit
    pol = _state.wsplit_policy
    _state.wsplit = split16() and (pol == "all" or (pol == "enc" and kind == "enc"))
    try:
        yield
    finally:
        _state.wsplit = old


@contextmanager
def hi_region(px):
    """Inside: ('fp16s' only) every conv whose input or output grid has at most `px` pixels is a high-precision layer --
    fp32 activation storage, operands split into hi + lo halves (lns_conv2d with LNS_W_UMMA_F16X2)."""
    old = _state.hi_px
    _state.hi_px = int(px * _state.hi_scale) if split16() else 0
    try:
        yield
    finally:
        _state.hi_px = old


def launch_count():
    """Number of liblns_b200 kernel launches issued by this thread so far (bench.py's gpu_launches)."""
    return _state.launches


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def device_of(t):
    """Context manager: make the device of CUDA tensor `t` current (single-process multi-GPU callers); no-op for CPU tensors,
    which the ops reject themselves."""
    if torch.is_tensor(t) and t.is_cuda:
        return torch.cuda.device(t.device)
    from contextlib import nullcontext
    return nullcontext()


def _need_cuda(t, what):
    if not t.is_cuda:
        raise LnsError(f"{what}: lns_b200 runs on CUDA tensors only (there is no CPU fallback); got device {t.device}")


# ---- activation handle ---------------------------------------------------------------------------------
class Act:
    """A [B,H,W,C] activation living in `t` (element (0,0,0,0) at t.data_ptr()); channel-last unless layout=NCHW.
    `bstride` is the distance between samples in elements (lets the K latent states of a rollout interleave as
    [B,K,...] without copies)."""
    __slots__ = ("t", "B", "H", "W", "C", "bstride", "layout", "tf32", "group", "gstride", "stats")

    def __init__(self, t, B, H, W, C, bstride=None, layout=NHWC, tf32=False, group=None, gstride=0):
        if t.is_cuda and t.device.index != torch.cuda.current_device():
            # kernels are enqueued on the CURRENT device's current stream: a tensor of another GPU would be read through the
            # wrong context.  The module / Rollout entry points install the guard themselves (ops.device_of).
            raise LnsError(f"activation on {t.device} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                           f"wrap the call in `with torch.cuda.device({t.device.index})`")
        self.t, self.B, self.H, self.W, self.C = t, B, H, W, C
        self.bstride = H * W * C if bstride is None else bstride
        self.layout = layout
        self.tf32 = tf32  # fp32 words holding TF32-rounded values: kernels round to nearest TF32 when they write it
        # two-level sample index (output of the decoder's projection only): sample s lives at
        # (s % group) * bstride + (s // group) * gstride -- `B // group` rollout steps of `group` trajectories, step-major
        self.group, self.gstride = group, gstride
        # (partial [B][nchunk][C][2] fp32, nchunk): per-channel sums of this activation emitted by the kernel that produced it
        # (conv_coarse.cu epilogue) -- a following GroupNorm only runs lns_norm_finalize
        self.stats = None

    @property
    def dtype(self):
        """LNS_* storage code passed to the library"""
        return TF32_CODE if self.tf32 else dt_code(self.t.dtype)

    @property
    def adtype(self):
        """what to pass as `dtype=` to get another Act of the same storage type"""
        return TF32 if self.tf32 else self.t.dtype

    @property
    def contiguous(self):
        return self.bstride == self.H * self.W * self.C

    @staticmethod
    def empty(B, H, W, C, dtype, device):
        tf = isinstance(dtype, str) and dtype == TF32
        return Act(torch.empty(B * H * W * C, dtype=torch.float32 if tf else dtype, device=device), B, H, W, C, tf32=tf)

    def like(self, C=None, dtype=None, H=None, W=None):
        return Act.empty(self.B, H or self.H, W or self.W, C or self.C, dtype if dtype is not None else self.adtype,
                         self.t.device)

    @staticmethod
    def from_nchw(x):
        """Wrap (no copy) an NCHW fp32 torch tensor."""
        _need_cuda(x, "Act.from_nchw")
        if x.dtype != torch.float32:
            raise LnsError("NCHW inputs must be fp32")
        x = x.contiguous()
        B, C, H, W = x.shape
        return Act(x, B, H, W, C, layout=NCHW)

    def to_nchw(self):
        """-> NCHW fp32 torch tensor [B,C,H,W]."""
        if self.layout == NCHW:
            return self.t.view(self.B, self.C, self.H, self.W)
        out = torch.empty(self.B, self.C, self.H, self.W, dtype=torch.float32, device=self.t.device)
        rc = _C.lib().lns_nhwc_to_nchw(_ptr(self.t), self.dtype, self.B, self.H, self.W, self.C, self.bstride,
                                       _ptr(out), self.C * self.H * self.W, _stream())
        check(rc, "lns_nhwc_to_nchw")
        _state.launches += 1
        return out

    def as_tokens(self):
        """[B, H*W, C] torch view (contiguous NHWC only) -- for tests."""
        assert self.layout == NHWC and self.contiguous
        return self.t.view(self.B, self.H * self.W, self.C)

    def to_torch_nhwc(self):
        assert self.layout == NHWC and self.contiguous
        return self.t.view(self.B, self.H, self.W, self.C)


def nchw_to_act(x, dtype=None):
    """NCHW fp32 torch tensor -> NHWC Act of `dtype` (one transposing kernel)."""
    _need_cuda(x, "nchw_to_act")
    x = x.contiguous().float()
    B, C, H, W = x.shape
    out = Act.empty(B, H, W, C, dtype or act_dtype(), x.device)
    rc = _C.lib().lns_nchw_to_nhwc(_ptr(x), B, C, H, W, C * H * W, _ptr(out.t), out.dtype, out.bstride, _stream())
    check(rc, "lns_nchw_to_nhwc")
    _state.launches += 1
    return out


# ---- weights ---------------------------------------------------------------------------------------------
_PACK_EPOCH = [0]  # process-wide (not thread-local: autograd's backward thread must see the same value)


def invalidate_packed():
    """Make every PackedFilter re-lay its filter at its next use.  Parameter updates are normally detected through the tensors'
    version counters; a CAPTURED training step needs the packing kernels inside the graph (the weights change between replays
    while the captured launch sequence does not), so it invalidates before every forward (lns_b200.train.GraphedTrainStep)."""
    _PACK_EPOCH[0] += 1


class PackedFilter:
    """Device-side re-laid copies of one conv / linear filter (OIHW fp32 source), built lazily per format and
    rebuilt when a source parameter changes (load_state_dict, .to(), optimizer step).  `key_fn` is a cheap fingerprint
    of the source parameters (data_ptr, version, device); the source tensors are only touched on a cache miss."""

    def __init__(self, weight_fn, bias_fn=None, key_fn=None, shape=None):
        self._weight_fn = weight_fn  # () -> fp32 OIHW tensor [Cout,Cin,KH,KW]
        self._bias_fn = bias_fn      # () -> fp32 [Cout] tensor or None
        self._key_fn = key_fn        # () -> hashable fingerprint of the sources
        self._cache = {}
        self._bias = None
        self._bias_key = None
        self._shape = shape

    @staticmethod
    def _fp(t):
        return None if t is None else (t.data_ptr(), t._version, str(t.device))

    @staticmethod
    def of(weight, bias=None):
        """weight: nn.Parameter [Cout,Cin,KH,KW] or [Cout,Cin] (nn.Linear); bias: nn.Parameter [Cout] or None"""
        def wfn():
            w = weight.detach()
            return w[:, :, None, None] if w.dim() == 2 else w

        def bfn():
            return None if bias is None else bias.detach()
        shape = tuple(weight.shape) if weight.dim() == 4 else (weight.shape[0], weight.shape[1], 1, 1)
        return PackedFilter(wfn, bfn, lambda: (PackedFilter._fp(weight), PackedFilter._fp(bias)), shape)

    @staticmethod
    def concat(weights, biases):
        """Stack several [Cout_i, Cin] linears along Cout (q|k|v in one GEMM); a None bias contributes zeros."""
        def wfn():
            ws = [w.detach() for w in weights]
            ws = [w[:, :, None, None] if w.dim() == 2 else w for w in ws]
            return torch.cat(ws, 0)

        def bfn():
            if all(b is None for b in biases):
                return None
            parts = [b.detach().float() if b is not None else torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
                     for w, b in zip(weights, biases)]
            return torch.cat(parts, 0)
        w0 = weights[0]
        shape = (sum(w.shape[0] for w in weights), w0.shape[1]) + ((1, 1) if w0.dim() == 2 else tuple(w0.shape[2:]))
        return PackedFilter(wfn, bfn, lambda: tuple(PackedFilter._fp(t) for t in list(weights) + list(biases)), shape)

    def key(self):
        return (self._key_fn(), _PACK_EPOCH[0]) if self._key_fn is not None else None

    def dims(self):
        if self._shape is None:
            self._shape = tuple(self._weight_fn().shape)
        return self._shape

    def get(self, fmt):
        key = self.key()
        ent = self._cache.get(fmt)
        if ent is not None and ent[0] == key:
            return ent[1]
        w = self._weight_fn()
        _need_cuda(w, "PackedFilter")
        Cout, Cin, KH, KW = w.shape
        nbytes = _C.lib().lns_packed_weight_bytes(Cout, Cin, KH, KW, fmt)
        if nbytes < 0:
            raise LnsError(f"filter {tuple(w.shape)} cannot be packed in format {fmt}")
        src = w.contiguous().float()
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
        rc = _C.lib().lns_pack_conv_weight(_ptr(src), Cout, Cin, KH, KW, fmt, _ptr(out), _stream())
        check(rc, "lns_pack_conv_weight")
        _state.launches += 1
        self._cache[fmt] = (key, out)
        return out

    def bias(self):
        if self._bias_fn is None:
            return None
        key = self.key()
        if self._bias_key != key or self._bias_key is None:
            b = self._bias_fn()
            self._bias = None if b is None else b.contiguous().float().clone()
            self._bias_key = key if key is not None else object()
        return self._bias


def composed_filter(first, second):
    """PackedFilter of  second(1x1) o first(kxk)  -- two convolutions with no non-linearity between them are one
    convolution: W'[o][i][tap] = sum_m W2[o][m] W1[m][i][tap],  b' = W2 b1 + b2  (composed in fp64, stored fp32)."""
    def wfn():
        w1, w2 = first._weight_fn(), second._weight_fn()
        return torch.einsum("om,mikl->oikl", w2[:, :, 0, 0].double(), w1.double()).float()

    def bfn():
        w2 = second._weight_fn()
        b1 = first._bias_fn() if first._bias_fn else None
        b2 = second._bias_fn() if second._bias_fn else None
        b = torch.zeros(w2.shape[0], dtype=torch.float64, device=w2.device)
        if b1 is not None:
            b = b + w2[:, :, 0, 0].double() @ b1.double()
        if b2 is not None:
            b = b + b2.double()
        return b.float()
    d1, d2 = first.dims(), second.dims()
    return PackedFilter(wfn, bfn, lambda: (first.key(), second.key()), (d2[0], d1[1], d1[2], d1[3]))


# ---- conv ---------------------------------------------------------------------------------------------------
def _umma_ok(x, Cin, Cout, y_layout):
    if x.layout != NHWC or y_layout != NHWC or Cout % 16 != 0 or x.bstride % 8 != 0:
        return False
    if fast16():
        return x.t.dtype == act_dtype() and Cin % 64 == 0
    if _state.precision == "tf32":  # fp32 words are read as TF32 by tcgen05.mma.kind::tf32
        return x.t.dtype == torch.float32 and Cin % 32 == 0
    return False


def _coarse_fits(Cin, Cout, dil, f32in, wsplit):
    """shared-memory budget of conv_coarse.cu: the halo planes of one MMA tile + a three-stage filter ring"""
    hwd = 8 + 2 * dil
    tile = (Cin // 64) * (2 if f32in else 1) * hwd * 2 * hwd * 128
    return tile + 3 * Cout * 128 * (2 if wsplit else 1) + 1536 <= 227 * 1024


def conv2d(x, filt, *, stride=1, dil=1, pad=(0, 0, 0, 0), pad_mode=(PAD_ZEROS, PAD_ZEROS), virt=None, use_bias=True,
           sample_bias=None, pro=None, act=ACT_NONE, pre_add=None, residual=None, out=None, out_dtype=None,
           out_layout=NHWC, engine=None, split=None):
    """y = act(conv(pro(resize(x))) + bias + sample_bias + pre_add) + residual.   (lns_conv2d)

    split=True (with engine=ENGINE_UMMA | ENGINE_HALO): use the split filter format LNS_W_UMMA_F16X2 -- f16 activations ->
    two MMAs per K step, fp32 activations (gather engine only) -> split in the producer, three MMAs.  The 'fp16s' precision
    mode chooses this per layer by itself.

    pad = (top, bottom, left, right) on the virtual input; pad_mode = (mode_h, mode_w); virt = (Hv, Wv) nearest-resize
    target (None: no resize); pro = (scale[B,C], shift[B,C], act) per-(sample,channel) affine applied to x first."""
    Cout, Cin, KH, KW = filt.dims()
    if Cin != x.C:
        raise LnsError(f"conv2d: filter expects Cin={Cin}, activation has C={x.C}")
    Hv, Wv = virt if virt is not None else (x.H, x.W)
    pt, pb, pl, pr = pad
    Hout = (Hv + pt + pb - dil * (KH - 1) - 1) // stride + 1
    Wout = (Wv + pl + pr - dil * (KW - 1) - 1) // stride + 1
    if (engine is None and KH == 1 and KW == 1 and Cout <= 4 and out_layout == NCHW and x.layout == NHWC and Cin % 4 == 0
            and Cin <= 512 and stride == 1 and virt is None and act == ACT_NONE and pre_add is None and residual is None
            and sample_bias is None and (out is None or out.t.dtype == torch.float32) and x.B <= 65535):
        return _pointwise_proj(x, filt, use_bias, pro, out)
    # 'fp16s': high-precision layers (ops.hi_region) keep fp32 storage and run with split operands
    hi = engine is None and split16() and _state.hi_px > 0 and min(x.H * x.W, Hout * Wout) <= _state.hi_px
    if hi and x.t.dtype == torch.float16 and Hout * Wout > _state.hi_px and not isinstance(pro, LazyNorm):
        hi = False  # a 16-bit activation leaving the region: nothing left to preserve, the ordinary engines are faster
    split_fmt = bool(split) and engine in (ENGINE_UMMA, ENGINE_HALO, ENGINE_COARSE)
    if hi:
        if out is None and out_dtype is None and out_layout == NHWC and Hout * Wout <= _state.hi_px:
            out_dtype = torch.float32
        coarse_ok = (KH == 3 and KW == 3 and stride == 1 and Cin in (64, 128) and Cout in (64, 128) and 1 <= dil <= 3
                     and pt == pb == pl == pr == dil and dil <= min(Hv, Wv) and sample_bias is None and pre_add is None
                     and _state.coarse and x.layout == NHWC and out_layout == NHWC and not x.tf32 and not _state.hi_exact
                     and _coarse_fits(Cin, Cout, dil, x.t.dtype == torch.float32, _state.hi_wsplit))
        hi_w = _state.hi_wsplit and (not _state.hi_wsplit_coarsest or 4 * min(x.H * x.W, Hout * Wout) <= _state.hi_px)
        # the block-halo engine's fp32 producer applies the pending per-sample affine + activation itself (conv_coarse.cu): the
        # GroupNorm-applied tensor is never written -- only its statistics kernel runs
        fuse_pro = (coarse_ok and pro is not None and x.t.dtype == torch.float32 and x.bstride % 4 == 0 and _state.coarse_pro)
        if fuse_pro:
            if isinstance(pro, LazyNorm):
                if pro.x is not x:
                    raise LnsError("conv2d: the pending normalisation belongs to a different activation")
                pro = pro.as_tuple()
        elif isinstance(pro, LazyNorm) and x.layout == NHWC:
            if pro.x is not x:
                raise LnsError("conv2d: the pending normalisation belongs to a different activation")
            x = pro.materialize(out_dtype=torch.float32)  # the normalised activation is not rounded to 16 bits
            pro = None
        elif pro is not None and x.layout == NHWC:
            x = affine_act(x, pro[0], pro[1], pro[2], out_dtype=torch.float32)
            pro = None
        coarse_ok = coarse_ok and (pro is None or fuse_pro)
        if (x.layout == NHWC and out_layout == NHWC and Cin % 64 == 0 and Cout % 16 == 0 and not x.tf32 and not _state.hi_exact
                and ((x.t.dtype == torch.float32 and x.bstride % 4 == 0) or (x.t.dtype == torch.float16 and x.bstride % 8 == 0))):
            if coarse_ok:
                # block-halo engine: the fp32 activation is read once per 8x8 block and split into hi + lo halo planes
                engine, split_fmt = ENGINE_COARSE, hi_w
            else:
                engine, split_fmt = ENGINE_UMMA, True   # gather engine: fp32 input 3 MMAs (x3), f16 input 2 MMAs (w2)
        else:
            engine = ENGINE_SIMT                        # tiny-channel layers: exact CUDA-core path
    fused_pro = engine == ENGINE_COARSE and pro is not None and x.t.dtype == torch.float32 and not isinstance(pro, LazyNorm)
    if engine is None:
        engine = ENGINE_UMMA if _umma_ok(x, Cin, Cout, out_layout) else ENGINE_SIMT
        if (engine == ENGINE_UMMA and x.t.dtype in H16_DTYPES and KH == 3 and KW == 3 and stride == 1 and Cin == 64
                and Cout in (64, 128)
                and pt == pb == pl == pr == dil and 1 <= dil <= 3 and Hout >= 16 and Wout >= 8):
            engine = ENGINE_HALO  # full-resolution layers: shared-memory halo + resident filter
        elif (engine == ENGINE_UMMA and x.t.dtype in H16_DTYPES and KH == 3 and KW == 3 and stride == 1 and Cin == 128
                and Cout == 128 and virt is None and (x.H, x.W) == (8, 8) and pt == pb == pl == pr == dil and dil in (1, 2)
                and tuple(pad_mode) == (PAD_CIRCULAR, PAD_CIRCULAR) and sample_bias is None and pre_add is None
                and x.B >= 4):
            engine = ENGINE_LATENT  # the propagator's latent grid: resident halos of 4 samples, streamed filter
        if split16() and _state.wsplit and x.t.dtype == torch.float16 and engine in (ENGINE_UMMA, ENGINE_HALO):
            # split filter, two MMAs per K step (the halo engine keeps both planes resident only for Cout = 64, dilation 1)
            split_fmt = engine == ENGINE_UMMA or (Cout == 64 and dil == 1)
    if isinstance(pro, LazyNorm):
        if pro.x is not x:
            raise LnsError("conv2d: the pending normalisation belongs to a different activation")
        if engine in (ENGINE_UMMA, ENGINE_HALO, ENGINE_LATENT, ENGINE_COARSE):
            x = pro.materialize()  # these engines gather with cp.async (no transform in flight)
            pro = None
        else:
            pro = pro.as_tuple()
    if engine in (ENGINE_UMMA, ENGINE_HALO, ENGINE_LATENT, ENGINE_COARSE) and pro is not None and not fused_pro:
        x = affine_act(x, pro[0], pro[1], pro[2])
        pro = None
    if out is None:
        if out_dtype is None:
            out_dtype = torch.float32 if out_layout == NCHW else act_dtype()
        if out_layout == NCHW:
            out = Act(torch.empty(x.B * Cout * Hout * Wout, dtype=torch.float32, device=x.t.device), x.B, Hout, Wout,
                      Cout, layout=NCHW)
        else:
            out = Act.empty(x.B, Hout, Wout, Cout, out_dtype, x.t.device)
    else:
        if (out.B, out.H, out.W, out.C) != (x.B, Hout, Wout, Cout):
            raise LnsError(f"conv2d: out shape {(out.B, out.H, out.W, out.C)} != {(x.B, Hout, Wout, Cout)}")
        if out.group is not None:
            raise LnsError("conv2d: a step-grouped output is only supported by the output projection (lns_pointwise_proj_steps)")
    d = ConvDesc()
    d.x, d.x_dtype, d.x_layout = x.t.data_ptr(), x.dtype, x.layout
    d.B, d.Hin, d.Win, d.Cin, d.x_bstride = x.B, x.H, x.W, x.C, x.bstride
    d.Hv, d.Wv = Hv, Wv
    d.KH, d.KW, d.stride, d.dil, d.pad_t, d.pad_l = KH, KW, stride, dil, pt, pl
    d.pad_mode_h, d.pad_mode_w = pad_mode
    fmt = W_SIMT_F32
    if engine in (ENGINE_UMMA, ENGINE_HALO, ENGINE_LATENT, ENGINE_COARSE):
        fmt = {torch.bfloat16: W_UMMA_BF16, torch.float16: W_UMMA_F16}.get(x.t.dtype, W_UMMA_TF32)
        if engine == ENGINE_COARSE and x.t.dtype == torch.float32:  # operands of the in-kernel hi/lo split
            fmt = W_UMMA_BF16 if _state.precision == "bf16" else W_UMMA_F16
        if split_fmt:
            fmt = W_UMMA_F16X2
    wbuf = filt.get(fmt)
    d.w, d.w_format, d.engine = wbuf.data_ptr(), fmt, engine
    bias = filt.bias() if use_bias else None
    d.bias = bias.data_ptr() if bias is not None else None
    d.sample_bias = sample_bias.data_ptr() if sample_bias is not None else None
    if pro is not None:
        d.pro_scale = pro[0].data_ptr() if pro[0] is not None else None
        d.pro_shift = pro[1].data_ptr() if pro[1] is not None else None
        d.pro_act = pro[2]
    d.act = act
    if pre_add is not None:
        d.pre_add, d.pre_add_dtype, d.pre_add_bstride = pre_add.t.data_ptr(), pre_add.dtype, pre_add.bstride
    if residual is not None:
        if (residual.B, residual.H, residual.W, residual.C) != (out.B, out.H, out.W, out.C):
            raise LnsError("conv2d: residual shape mismatch")
        d.residual, d.res_dtype, d.res_bstride = residual.t.data_ptr(), residual.dtype, residual.bstride
    d.y, d.y_dtype, d.y_layout = out.t.data_ptr(), out.dtype, out.layout
    d.Hout, d.Wout, d.Cout, d.y_bstride = Hout, Wout, Cout, out.bstride
    if hi and engine == ENGINE_COARSE and _state.coarse_stats and out.layout == NHWC and out.group is None:
        # the statistics of the GroupNorm that follows, from the epilogue's registers
        nchunk = _C.lib().lns_conv_stats_chunks(Hout, Wout)
        part = torch.empty(x.B * nchunk * Cout * 4, dtype=torch.float32, device=x.t.device)
        d.stats = part.data_ptr()
        out.stats = (part, nchunk)
    tok = _mark(f"conv e{engine} {KH}x{KW} s{stride} d{dil} {Cin}->{Cout} @{Hout}x{Wout}"
                f"{' up' if virt is not None else ''}{' pro' if pro is not None else ''}"
                f"{' act' if act else ''}{' res' if residual is not None else ''} "
                f"{'h16' if x.t.dtype in H16_DTYPES else 'f32'}->{'h16' if out.t.dtype in H16_DTYPES else 'f32'}"
                f"{' split' if split_fmt else ''}",
                flops=2.0 * x.B * Hout * Wout * Cout * KH * KW * Cin, nbytes=_abytes(x, out, residual, pre_add))
    rc = _C.lib().lns_conv2d(ctypes.byref(d), _stream())
    check(rc, "lns_conv2d")
    _done(tok)
    _state.launches += 1
    return out


def _pointwise_proj(x, filt, use_bias, pro, out):
    """Cout <= 4 output projection (NHWC -> NCHW fp32) with the norm/activation prologue: lns_pointwise_proj."""
    Cout, Cin, _, _ = filt.dims()
    w = filt.get(W_SIMT_F32)  # [1 tap][Cin][Cout] fp32 == [Cin][Cout]; the kernel wants [Cout][Cin]
    ent = filt._cache.get("proj")
    if ent is None or ent[0] != w.data_ptr():
        ent = (w.data_ptr(), w.view(torch.float32).view(Cin, Cout).t().contiguous())
        filt._cache["proj"] = ent
    wt = ent[1]
    if out is None:
        out = Act(torch.empty(x.B * Cout * x.H * x.W, dtype=torch.float32, device=x.t.device), x.B, x.H, x.W, Cout,
                  layout=NCHW)
    bias = filt.bias() if use_bias else None
    if isinstance(pro, LazyNorm):
        pro = pro.as_tuple()
    sc, sh, pa = (pro if pro is not None else (None, None, ACT_NONE))
    tok = _mark(f"proj @{x.H}x{x.W}", flops=2.0 * x.B * x.H * x.W * Cin * Cout, nbytes=_abytes(x) + 4.0 * x.B * x.H * x.W * Cout)
    if out.group is not None:
        if out.B % out.group != 0:
            raise LnsError(f"proj: {out.B} samples are not whole steps of {out.group} trajectories")
        rc = _C.lib().lns_pointwise_proj_steps(_ptr(x.t), x.dtype, x.B, x.H * x.W, Cin, x.bstride, _ptr(wt), _ptr(bias), Cout,
                                               _ptr(sc), _ptr(sh), pa, _ptr(out.t), out.bstride, out.group, out.gstride,
                                               _stream())
    else:
        rc = _C.lib().lns_pointwise_proj(_ptr(x.t), x.dtype, x.B, x.H * x.W, Cin, x.bstride, _ptr(wt), _ptr(bias), Cout,
                                         _ptr(sc), _ptr(sh), pa, _ptr(out.t), out.bstride, _stream())
    check(rc, "lns_pointwise_proj")
    _done(tok)
    _state.launches += 1
    return out


# ---- normalisation ------------------------------------------------------------------------------------------
def chan_stats(x):
    """-> (partial [B,nchunk,C,2] fp32, nchunk)"""
    nchunk = _C.lib().lns_chan_stats_chunks(x.H, x.W)
    part = torch.empty(x.B * nchunk * x.C * 2, dtype=torch.float32, device=x.t.device)
    rc = _C.lib().lns_chan_stats(_ptr(x.t), x.dtype, x.B, x.H, x.W, x.C, x.bstride, _ptr(part), _stream())
    check(rc, "lns_chan_stats")
    _state.launches += 1
    return part, nchunk


def group_norm_affine(x, groups, eps, gamma=None, beta=None, prescale=None):
    """GroupNorm(groups) statistics of (x * prescale) folded with gamma/beta into per-(sample,channel) (scale, shift).
    One fused kernel for samples of <= 1024 pixels, statistics + finalize kernels above that."""
    scale = torch.empty(x.B * x.C, dtype=torch.float32, device=x.t.device)
    shift = torch.empty_like(scale)
    g = gamma.detach().float().contiguous() if gamma is not None else None
    b = beta.detach().float().contiguous() if beta is not None else None
    if x.stats is not None:
        # the producing kernel left per-channel partial sums: only the finalize runs (no read of x)
        part, nchunk = x.stats
        tok = _mark(f"gn_finalize C{x.C} @{x.H}x{x.W}", nbytes=4.0 * part.numel())
        rc = _C.lib().lns_norm_finalize_centred(_ptr(part), x.B, nchunk, x.C, x.H * x.W, groups, float(eps), _ptr(g), _ptr(b),
                                                _ptr(prescale), _ptr(scale), _ptr(shift), _stream())
        check(rc, "lns_norm_finalize_centred")
        _done(tok)
        _state.launches += 1
        return scale, shift
    nchunk = _C.lib().lns_chan_stats_chunks(x.H, x.W)
    part = None
    if nchunk > 1:
        part = torch.empty(x.B * nchunk * x.C * 2, dtype=torch.float32, device=x.t.device)
    tok = _mark(f"gn_stats C{x.C} @{x.H}x{x.W}", nbytes=_abytes(x))
    rc = _C.lib().lns_group_norm_affine(_ptr(x.t), x.dtype, x.B, x.H, x.W, x.C, x.bstride, groups, float(eps), _ptr(g),
                                        _ptr(b), _ptr(prescale), _ptr(part), _ptr(scale), _ptr(shift), _stream())
    check(rc, "lns_group_norm_affine")
    _done(tok)
    _state.launches += 1 if nchunk == 1 else 2
    return scale, shift


class LazyNorm:
    """A GroupNorm-family normalisation (+ activation) of `x` that has not been executed yet.  The consumer decides how:
    a CUDA-core conv folds the per-(sample,channel) affine into its gather (`as_tuple`), a tcgen05 conv needs the
    normalised tensor (`materialize`: ONE fused statistics+apply kernel when the sample fits in shared memory, else
    statistics kernel + affine kernel)."""

    def __init__(self, x, groups, eps, gamma=None, beta=None, prescale=None, act=ACT_NONE):
        self.x, self.groups, self.eps, self.gamma, self.beta, self.prescale, self.act = x, groups, eps, gamma, beta, prescale, act
        self._affine = None

    def affine(self):
        if self._affine is None:
            self._affine = group_norm_affine(self.x, self.groups, self.eps, self.gamma, self.beta, self.prescale)
        return self._affine

    def as_tuple(self):
        s, t = self.affine()
        return (s, t, self.act)

    def materialize(self, out_dtype=None):
        x = self.x
        if (self._affine is None and x.layout == NHWC and x.stats is None
                and _C.lib().lns_group_norm_act_supported(x.H, x.W, x.C) and x.bstride % 4 == 0):
            out = x.like(dtype=out_dtype)
            g = self.gamma.detach().float().contiguous() if self.gamma is not None else None
            b = self.beta.detach().float().contiguous() if self.beta is not None else None
            tok = _mark(f"gn_act_fused C{x.C} @{x.H}x{x.W}", nbytes=_abytes(x, out))
            rc = _C.lib().lns_group_norm_act(_ptr(x.t), x.dtype, x.B, x.H, x.W, x.C, x.bstride, self.groups, float(self.eps),
                                             _ptr(g), _ptr(b), _ptr(self.prescale), self.act, _ptr(out.t), out.dtype,
                                             out.bstride, _stream())
            check(rc, "lns_group_norm_act")
            _done(tok)
            _state.launches += 1
            return out
        s, t = self.affine()
        return affine_act(x, s, t, self.act, out_dtype=out_dtype)


def affine_act(x, scale, shift, act=ACT_NONE, out_dtype=None):
    """y = act(x*scale[b,c] + shift[b,c]) as a new contiguous Act."""
    out = x.like(dtype=out_dtype)
    tok = _mark(f"affine_act C{x.C} @{x.H}x{x.W}", nbytes=_abytes(x, out))
    rc = _C.lib().lns_affine_act(_ptr(x.t), x.dtype, x.bstride, x.B, x.H * x.W, x.C, _ptr(scale), _ptr(shift), act,
                                 _ptr(out.t), out.dtype, out.bstride, _stream())
    check(rc, "lns_affine_act")
    _done(tok)
    _state.launches += 1
    return out


def as_h16(x):
    """On the 16-bit paths: an fp32-stored activation (a high-precision stage's output) as a 16-bit copy, for the blocks
    whose fused kernels take 16-bit input; anything else is returned as it is."""
    if fast16() and x.layout == NHWC and x.t.dtype == torch.float32 and not x.tf32:
        return affine_act(x, None, None, ACT_NONE, out_dtype=act_dtype())
    return x


def layernorm(x, gamma, beta, eps, pe=None, out_dtype=None):
    """LayerNorm over C per pixel/token (+ pe[token]) -> new Act."""
    assert x.contiguous and x.layout == NHWC
    out = x.like(dtype=out_dtype)
    g = gamma.detach().float().contiguous() if gamma is not None else None
    b = beta.detach().float().contiguous() if beta is not None else None
    tok = _mark(f"layernorm")
    rc = _C.lib().lns_layernorm(_ptr(x.t), x.dtype, x.B, x.H * x.W, x.C, _ptr(g), _ptr(b), float(eps), _ptr(pe),
                                _ptr(out.t), out.dtype, _stream())
    check(rc, "lns_layernorm")
    _done(tok)
    _state.launches += 1
    return out


def channel_gate(x, gate):
    assert x.contiguous and x.layout == NHWC
    out = x.like()
    rc = _C.lib().lns_channel_gate(_ptr(x.t), x.dtype, x.B, x.H * x.W, x.C, _ptr(gate), _ptr(out.t), out.dtype,
                                   _stream())
    check(rc, "lns_channel_gate")
    _state.launches += 1
    return out


# ---- attention pieces -----------------------------------------------------------------------------------------
def attention(qkv, heads, dh, scale, out_dtype=None):
    """qkv: Act [B,H,W,3*heads*dh] (tokens = pixels) -> Act [B,H,W,heads*dh]"""
    assert qkv.contiguous and qkv.C == 3 * heads * dh
    out = qkv.like(C=heads * dh, dtype=out_dtype)
    tok = _mark(f"attention")
    rc = _C.lib().lns_attention(_ptr(qkv.t), qkv.dtype, qkv.B, qkv.H * qkv.W, heads, dh, float(scale), _ptr(out.t),
                                out.dtype, _stream())
    check(rc, "lns_attention")
    _done(tok)
    _state.launches += 1
    return out


def axis_mean(x, axis):
    """axis 0: mean over H -> Act [B, W, 1, C] fp32; axis 1: mean over W -> Act [B, H, 1, C] fp32"""
    keep = x.W if axis == 0 else x.H
    out = Act.empty(x.B, keep, 1, x.C, torch.float32, x.t.device)
    tok = _mark(f"axis_mean")
    rc = _C.lib().lns_axis_mean(_ptr(x.t), x.dtype, x.B, x.H, x.W, x.C, x.bstride, axis, _ptr(out.t), _stream())
    check(rc, "lns_axis_mean")
    _done(tok)
    _state.launches += 1
    return out


def lowrank_kernel(qk, heads, d, cos_tab, sin_tab, scaling=1.0):
    """qk: Act [B, n, 1, 2*heads*d] -> torch fp32 [B, heads, n, n]"""
    assert qk.contiguous and qk.C == 2 * heads * d
    n = qk.H * qk.W
    K = torch.empty(qk.B, heads, n, n, dtype=torch.float32, device=qk.t.device)
    tok = _mark(f"lowrank")
    rc = _C.lib().lns_lowrank_kernel(_ptr(qk.t), qk.dtype, qk.B, n, heads, d, _ptr(cos_tab), _ptr(sin_tab),
                                     float(scaling), _ptr(K), _stream())
    check(rc, "lns_lowrank_kernel")
    _done(tok)
    _state.launches += 1
    return K


def axial_contract(u, K, heads, axis, out_dtype=None):
    assert u.contiguous and u.layout == NHWC and u.C % heads == 0
    out = u.like(dtype=out_dtype)
    tok = _mark(f"axial C{u.C} @{u.H}x{u.W}")
    rc = _C.lib().lns_axial_contract(_ptr(u.t), u.dtype, u.B, u.H, u.W, heads, u.C // heads, _ptr(K), axis,
                                     _ptr(out.t), out.dtype, _stream())
    check(rc, "lns_axial_contract")
    _done(tok)
    _state.launches += 1
    return out


def fablock_core_supported(x, dim_head):
    return (x.t.dtype in H16_DTYPES and x.layout == NHWC and x.contiguous
            and bool(_C.lib().lns_fablock_core_supported(x.H, x.W, x.C, dim_head)))


def fablock_prepass(u, eps, gamma, beta, staged=False):
    """-> (scale [B*C], shift [B*C], pooled_x Act [B,H,1,C] fp32, pooled_y Act [B,W,1,C] fp32) in one read of u.
    staged=True: also the normalised sample as the shared-memory image fablock_full_staged reads (fifth value)."""
    dev = u.t.device
    scale = torch.empty(u.B * u.C, dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    px = Act.empty(u.B, u.H, 1, u.C, torch.float32, dev)   # fp32: tiny, and feeds per-sample low-rank kernels
    py = Act.empty(u.B, u.W, 1, u.C, torch.float32, dev)
    g = gamma.detach().float().contiguous() if gamma is not None else None
    b = beta.detach().float().contiguous() if beta is not None else None
    if staged:
        st = torch.empty(u.B * u.H * u.W * u.C, dtype=u.t.dtype, device=dev)
        tok = _mark(f"fablock_prepass @{u.H}x{u.W}", nbytes=2 * _abytes(u))
        rc = _C.lib().lns_fablock_prepass_staged(_ptr(u.t), u.dtype, u.B, u.H, u.W, u.C, u.bstride, float(eps), _ptr(g), _ptr(b),
                                                 _ptr(scale), _ptr(shift), _ptr(px.t), _ptr(py.t), _ptr(st), _stream())
        check(rc, "lns_fablock_prepass_staged")
        _done(tok)
        _state.launches += 1
        return scale, shift, px, py, st
    tok = _mark(f"fablock_prepass @{u.H}x{u.W}", nbytes=_abytes(u))
    rc = _C.lib().lns_fablock_prepass(_ptr(u.t), u.dtype, u.B, u.H, u.W, u.C, u.bstride, float(eps), _ptr(g), _ptr(b),
                                      _ptr(scale), _ptr(shift), _ptr(px.t), _ptr(py.t), _stream())
    check(rc, "lns_fablock_prepass")
    _done(tok)
    _state.launches += 1
    return scale, shift, px, py


def fablock_core(u, gn_scale, gn_shift, w_in_proj, Kx, Ky, heads, eps):
    """Fused in_proj -> axial contractions -> InstanceNorm of FABlock2D (16-bit paths): Act [B,H,W,64] -> [B,H,W,heads*64]."""
    out = Act.empty(u.B, u.H, u.W, heads * 64, u.t.dtype, u.t.device)
    w = w_in_proj.detach().float().contiguous()
    tok = _mark(f"fablock_core @{u.H}x{u.W}")
    rc = _C.lib().lns_fablock_core(_ptr(u.t), u.dtype, u.B, u.H, u.W, heads, _ptr(gn_scale), _ptr(gn_shift), _ptr(w), _ptr(Kx),
                                   _ptr(Ky), float(eps), _ptr(out.t), _stream())
    check(rc, "lns_fablock_core")
    _done(tok)
    _state.launches += 1
    return out


def sablock_fused_supported(x, heads, dim_head):
    return (fast16() and x.t.dtype == act_dtype() and x.layout == NHWC and x.contiguous
            and bool(_C.lib().lns_sablock_fused_supported(x.H * x.W, x.C, heads, dim_head)))


def sablock_fused(x, heads, ln_g, ln_b, ln_eps, pe, wqkv16, bv, wproj16, bproj, scale):
    """Whole SABlock (LN + pe, q|k|v, attention, projection, residual) in one kernel: Act [B,H,W,128] -> Act of the same shape."""
    out = x.like()
    n_ = x.H * x.W
    tok = _mark("sablock_fused", flops=x.B * (2.0 * n_ * x.C * 3 * heads * 64 + 4.0 * heads * n_ * n_ * 64 + 2.0 * n_ * heads * 64 * x.C),
                nbytes=_abytes(x, out))
    rc = _C.lib().lns_sablock_fused(_ptr(x.t), x.dtype, x.B, x.H * x.W, heads, _ptr(ln_g), _ptr(ln_b), float(ln_eps), _ptr(pe),
                                    _ptr(wqkv16), _ptr(bv), _ptr(wproj16), _ptr(bproj), float(scale), _ptr(out.t), _stream())
    check(rc, "lns_sablock_fused")
    _done(tok)
    _state.launches += 1
    return out


def ffn_fused_supported(x, f1, f2):
    """x Act, f1/f2 PackedFilters of the two bias-free 1x1 convs"""
    d1, d2 = f1.dims(), f2.dims()
    return (fast16() and x.t.dtype == act_dtype() and x.layout == NHWC and x.bstride % 8 == 0 and x.B * x.H * x.W < (1 << 22)
            and d1[2:] == (1, 1) and d2[2:] == (1, 1) and d1[0] == d2[1] and d1[1] == x.C and d2[0] == x.C
            and f1.bias() is None and f2.bias() is None
            and bool(_C.lib().lns_ffn_fused_supported(x.C, d1[0])))


def ffn_fused(x, scale, shift, f1, f2):
    """y = x + W2 . GELU(W1 . (x*scale + shift)): GroupNorm apply + both 1x1 convs + residual in one tcgen05 kernel."""
    fmt = W_UMMA_F16 if x.t.dtype == torch.float16 else W_UMMA_BF16
    w1, w2 = f1.get(fmt), f2.get(fmt)
    out = x.like()
    tok = _mark(f"ffn_fused C{x.C} @{x.H}x{x.W}")
    rc = _C.lib().lns_ffn_fused(_ptr(x.t), x.dtype, x.B, x.H * x.W, x.C, x.bstride, _ptr(scale), _ptr(shift), _ptr(w1), _ptr(w2),
                                _ptr(out.t), out.bstride, _stream())
    check(rc, "lns_ffn_fused")
    _done(tok)
    _state.launches += 1
    return out


def fa_axis_kernel_supported(n, dim, hidden, latent, heads, d):
    return bool(_C.lib().lns_fa_axis_kernel_supported(n, dim, hidden, latent, heads, d))


def fa_axis_kernel(pooled, heads, w1t, ln_g, ln_b, ln_eps, wf1t, wf2t, bf2, wqk16, cos_tab, sin_tab, scaling=1.0):
    """Pooled branch of one FABlock2D axis in one kernel: pooled Act [B,n,1,64] fp32 -> torch fp32 [B, heads, n, n]."""
    assert pooled.contiguous and pooled.t.dtype == torch.float32 and pooled.C == 64
    n = pooled.H * pooled.W
    K = torch.empty(pooled.B, heads, n, n, dtype=torch.float32, device=pooled.t.device)
    tok = _mark("fa_axis", flops=pooled.B * n * (2.0 * 64 * 64 + 4.0 * 64 * 128 + 2.0 * 64 * 2 * heads * 128 + 2.0 * heads * n * 128),
                nbytes=_abytes(pooled) + 4.0 * pooled.B * heads * n * n)
    rc = _C.lib().lns_fa_axis_kernel(_ptr(pooled.t), dt_code(wqk16.dtype), pooled.B, n, heads, _ptr(w1t), _ptr(ln_g), _ptr(ln_b),
                                     float(ln_eps), _ptr(wf1t), _ptr(wf2t), _ptr(bf2), _ptr(wqk16), _ptr(cos_tab), _ptr(sin_tab),
                                     float(scaling), _ptr(K), _stream())
    check(rc, "lns_fa_axis_kernel")
    _done(tok)
    _state.launches += 1
    return K


def fablock_full_supported(x, dim_head, dim_out):
    return (x.t.dtype in H16_DTYPES and x.layout == NHWC and x.contiguous
            and bool(_C.lib().lns_fablock_full_supported(x.H, x.W, x.C, dim_head, dim_out)))


def fablock_tc_supported(x, dim_head, dim_out):
    return (_state.fablock_tc and x.t.dtype in H16_DTYPES and x.layout == NHWC and x.contiguous
            and bool(_C.lib().lns_fablock_tc_supported(x.H, x.W, x.C, dim_head, dim_out)))


def fablock_tc(u, gn_scale, gn_shift, w_in_proj, Kx, Ky, heads, eps, w_out1, w_out2):
    """fablock_full with every contraction on tcgen05 (csrc/fablock_tc.cu; 16x16 and 32x32 samples)."""
    out = u.like()
    wi = w_in_proj.detach().float().contiguous()
    w1 = w_out1.detach().float().reshape(w_out1.shape[0], -1).contiguous()
    w2 = w_out2.detach().float().reshape(w_out2.shape[0], -1).contiguous()
    hw_ = u.H * u.W
    tok = _mark(f"fablock_tc @{u.H}x{u.W}",
                flops=u.B * (2.0 * hw_ * 64 * heads * 64 * 2 + 2.0 * heads * (u.H * u.H * u.W + u.H * u.W * u.W) * 64 + 2.0 * hw_ * 64 * 64),
                nbytes=_abytes(u, out) + 4.0 * u.B * heads * (u.H * u.H + u.W * u.W))
    rc = _C.lib().lns_fablock_tc(_ptr(u.t), u.dtype, u.B, u.H, u.W, heads, _ptr(gn_scale), _ptr(gn_shift), _ptr(wi), _ptr(Kx),
                                 _ptr(Ky), float(eps), _ptr(w1), _ptr(w2), _ptr(out.t), _stream())
    check(rc, "lns_fablock_tc")
    _done(tok)
    _state.launches += 1
    return out


def fablock_full(u, gn_scale, gn_shift, w_in_proj, Kx, Ky, heads, eps, w_out1, w_out2):
    """Whole FABlock2D after the pooled branch in one kernel per sample (16-bit paths): Act [B,H,W,64] -> Act [B,H,W,64]
    = to_out(InstanceNorm(Ky . Kx . in_proj(GN(u)))) + u, with to_out's two 1x1 convs on tcgen05 / TMEM."""
    out = u.like()
    wi = w_in_proj.detach().float().contiguous()
    w1 = w_out1.detach().float().reshape(w_out1.shape[0], -1).contiguous()
    w2 = w_out2.detach().float().reshape(w_out2.shape[0], -1).contiguous()
    hw_ = u.H * u.W
    tok = _mark(f"fablock_full @{u.H}x{u.W}",
                flops=u.B * (2.0 * hw_ * 64 * heads * 64 * 2 + 2.0 * heads * (u.H * u.H * u.W + u.H * u.W * u.W) * 64 + 2.0 * hw_ * 64 * 64),
                nbytes=_abytes(u, out) + 4.0 * u.B * heads * (u.H * u.H + u.W * u.W))
    rc = _C.lib().lns_fablock_full(_ptr(u.t), u.dtype, u.B, u.H, u.W, heads, _ptr(gn_scale), _ptr(gn_shift), _ptr(wi), _ptr(Kx),
                                   _ptr(Ky), float(eps), _ptr(w1), _ptr(w2), _ptr(out.t), _stream())
    check(rc, "lns_fablock_full")
    _done(tok)
    _state.launches += 1
    return out


def fablock_full_staged_supported(x, dim_head, dim_out):
    return (_state.fablock_staged and x.t.dtype in H16_DTYPES and x.layout == NHWC and x.contiguous
            and bool(_C.lib().lns_fablock_full_staged_supported(x.H, x.W, x.C, dim_head, dim_out)))


def fablock_staged_operands(w_in_proj, w_out1, heads, dt16):
    """Sample-independent operands of fablock_full_staged: in_proj slices as 16-bit rows padded to 72 elements, to_out[1]
    head-major in fp32 (prepare once per parameter version)."""
    with torch.no_grad():
        wi = w_in_proj.detach().float().reshape(heads, 64, 64)
        w_in16 = torch.zeros(heads, 64, 72, dtype=dt16, device=wi.device)
        w_in16[:, :, :64] = wi.to(dt16)
        w1h = w_out1.detach().float().reshape(64, heads, 64).permute(1, 0, 2).contiguous()
    return w_in16.contiguous(), w1h


def fablock_full_staged(u_staged, u, w_in16, Kx, Ky, heads, eps, w1h, w_out2):
    """fablock_full on pre-staged operands (csrc/fablock_full.cu, producer-warp kernel): u_staged from fablock_prepass(staged=True),
    (w_in16, w1h) from fablock_staged_operands; u is the raw block input (skip connection)."""
    out = u.like()
    w2 = w_out2.detach().float().reshape(w_out2.shape[0], -1).contiguous()
    hw_ = u.H * u.W
    tok = _mark(f"fablock_full @{u.H}x{u.W}",
                flops=u.B * (2.0 * hw_ * 64 * heads * 64 * 2 + 2.0 * heads * (u.H * u.H * u.W + u.H * u.W * u.W) * 64 + 2.0 * hw_ * 64 * 64),
                nbytes=_abytes(u, out) + 4.0 * u.B * heads * (u.H * u.H + u.W * u.W))
    rc = _C.lib().lns_fablock_full_staged(_ptr(u_staged), _ptr(u.t), u.dtype, u.B, u.H, u.W, heads, _ptr(w_in16), _ptr(Kx), _ptr(Ky),
                                          float(eps), _ptr(w1h), _ptr(w2), _ptr(out.t), _stream())
    check(rc, "lns_fablock_full_staged")
    _done(tok)
    _state.launches += 1
    return out


# ---- backward kernels (training rollout; csrc/backward.cu) ---------------------------------------------------------
def _f32_nhwc(a, what):
    if a.layout != NHWC or a.t.dtype != torch.float32 or a.tf32:
        raise LnsError(f"{what}: fp32 NHWC activations only")


def conv2d_wgrad(x, dy, dW, *, KH, KW, dil=1, pad=(0, 0, 0, 0), pad_mode=(PAD_ZEROS, PAD_ZEROS), pro=None, tensor_core=False,
                 out_scale=1.0, out_scale_dev=None):
    """dW (torch fp32 [Cout,Cin,KH,KW], accumulated into) += out_scale * filter gradient of the same-size stride-1 conv whose
    forward read pro(x) (pro = (scale[B,Cin] | None, shift | None, act)) and whose output gradient is dy.  tensor_core: TF32
    mma.sync with hi + lo split operands (fp32-class) instead of CUDA-core FMAs.  (lns_conv2d_wgrad)"""
    _f32_nhwc(x, "conv2d_wgrad")
    _f32_nhwc(dy, "conv2d_wgrad")
    if (x.B, x.H, x.W) != (dy.B, dy.H, dy.W) or tuple(dW.shape) != (dy.C, x.C, KH, KW) or not dW.is_contiguous():
        raise LnsError("conv2d_wgrad: shape mismatch")
    nbytes = _C.lib().lns_conv2d_wgrad_work_bytes(x.B, x.H, x.W, x.C, dy.C, KH, KW)
    work = torch.empty(nbytes, dtype=torch.uint8, device=x.t.device)
    sc, sh, pa = pro if pro is not None else (None, None, ACT_NONE)
    tok = _mark(f"wgrad {KH}x{KW} d{dil} {x.C}->{dy.C} @{x.H}x{x.W}", flops=2.0 * x.B * x.H * x.W * dy.C * KH * KW * x.C,
                nbytes=_abytes(x, dy))
    rc = _C.lib().lns_conv2d_wgrad(_ptr(x.t), x.bstride, _ptr(sc), _ptr(sh), pa, _ptr(dy.t), dy.bstride, x.B, x.H, x.W, x.C, dy.C,
                                   KH, KW, dil, pad[0], pad[2], pad_mode[0], pad_mode[1], 1 if tensor_core else 0, float(out_scale),
                                   _ptr(out_scale_dev), _ptr(work), _ptr(dW), _stream())
    check(rc, "lns_conv2d_wgrad")
    _done(tok)
    _state.launches += 2


def chan_sum_accum(dy, grad, out_scale=1.0, out_scale_dev=None):
    """grad[c] += out_scale * sum over samples and pixels of dy (bias gradient)."""
    _f32_nhwc(dy, "chan_sum_accum")
    work = torch.empty(_C.lib().lns_chan_sum_slices(dy.B) * dy.C, dtype=torch.float32, device=dy.t.device)
    rc = _C.lib().lns_chan_sum_accum(_ptr(dy.t), dy.bstride, dy.B, dy.H * dy.W, dy.C, float(out_scale), _ptr(out_scale_dev),
                                     _ptr(work), _ptr(grad), _stream())
    check(rc, "lns_chan_sum_accum")
    _state.launches += 2


def act_bwd(dy, pre, act):
    """dy * act'(pre) as a new Act (contiguous fp32 operands)."""
    _f32_nhwc(dy, "act_bwd")
    _f32_nhwc(pre, "act_bwd")
    if not (dy.contiguous and pre.contiguous) or (dy.B, dy.H, dy.W, dy.C) != (pre.B, pre.H, pre.W, pre.C):
        raise LnsError("act_bwd: contiguous activations of equal shape only")
    out = dy.like()
    rc = _C.lib().lns_act_bwd(_ptr(dy.t), _ptr(pre.t), dy.B * dy.H * dy.W * dy.C, act, _ptr(out.t), _stream())
    check(rc, "lns_act_bwd")
    _state.launches += 1
    return out


def absmax(t):
    """max |t| of a contiguous fp32 CUDA tensor as a Python float (one kernel + a 4-byte read back: it synchronises)."""
    out = torch.zeros(1, dtype=torch.int32, device=t.device)
    rc = _C.lib().lns_absmax(_ptr(t), t.numel(), _ptr(out), _stream())
    check(rc, "lns_absmax")
    _state.launches += 1
    return float(out.view(torch.float32).item())


def pixel_dot(dy, x, out, accumulate=True):
    """out[b][c] (+)= sum over pixels of dy * x (x None: of dy): gradients of a per-sample shift / gate.  out: torch fp32 [B*C]."""
    _f32_nhwc(dy, "pixel_dot")
    if x is not None:
        _f32_nhwc(x, "pixel_dot")
    rc = _C.lib().lns_pixel_dot(_ptr(dy.t), dy.bstride, _ptr(x.t) if x is not None else None, x.bstride if x is not None else 0,
                                dy.B, dy.H * dy.W, dy.C, 1 if accumulate else 0, _ptr(out), _stream())
    check(rc, "lns_pixel_dot")
    _state.launches += 1


def scale_add(x, scale=None, skip=None):
    """x * scale[b][c] + skip as a new contiguous fp32 Act."""
    _f32_nhwc(x, "scale_add")
    if not x.contiguous or (skip is not None and not skip.contiguous):
        raise LnsError("scale_add: contiguous activations only")
    out = x.like()
    rc = _C.lib().lns_scale_add(_ptr(x.t), _ptr(scale), _ptr(skip.t) if skip is not None else None, x.B, x.H * x.W, x.C, _ptr(out.t),
                                _stream())
    check(rc, "lns_scale_add")
    _state.launches += 1
    return out


def loss_scale(t, target=64.0):
    """Device-side loss scale of a contiguous fp32 CUDA tensor: -> torch fp32 [2] = (S, 1 / S), S = the power of two that brings
    max |t| to about `target` (no host synchronisation)."""
    bits = torch.zeros(1, dtype=torch.int32, device=t.device)
    s2 = torch.empty(2, dtype=torch.float32, device=t.device)
    rc = _C.lib().lns_absmax(_ptr(t), t.numel(), _ptr(bits), _stream())
    check(rc, "lns_absmax")
    rc = _C.lib().lns_loss_scale(_ptr(bits), float(target), _ptr(s2), _stream())
    check(rc, "lns_loss_scale")
    _state.launches += 2
    return s2


def scale_by(x, scalar_dev):
    """x * (*scalar_dev) as a new contiguous fp32 Act (scalar_dev: a one-element fp32 CUDA tensor / view)."""
    _f32_nhwc(x, "scale_by")
    if not x.contiguous:
        raise LnsError("scale_by: contiguous activations only")
    out = x.like()
    rc = _C.lib().lns_scale_by(_ptr(x.t), _ptr(scalar_dev), x.B * x.H * x.W * x.C, _ptr(out.t), _stream())
    check(rc, "lns_scale_by")
    _state.launches += 1
    return out


def group_norm_bwd(x, dy, groups, eps, gamma, dskip=None, dgamma=None, dbeta=None, out_scale=1.0, out_scale_dev=None):
    """Gradient of GroupNorm(groups, C, eps)(x) w.r.t. x (+ dskip), and out_scale * (dgamma, dbeta) accumulated into the given
    [C] tensors."""
    _f32_nhwc(x, "group_norm_bwd")
    _f32_nhwc(dy, "group_norm_bwd")
    out = x.like()
    part_g = part_b = None
    if dgamma is not None:
        part_g = torch.empty(x.B * x.C, dtype=torch.float32, device=x.t.device)
        part_b = torch.empty_like(part_g)
    g = gamma.detach().float().contiguous() if gamma is not None else None
    rc = _C.lib().lns_group_norm_bwd(_ptr(x.t), x.bstride, _ptr(dy.t), dy.bstride, _ptr(dskip.t) if dskip is not None else None,
                                     dskip.bstride if dskip is not None else 0, x.B, x.H * x.W, x.C, groups, float(eps), _ptr(g),
                                     _ptr(out.t), out.bstride, _ptr(part_g), _ptr(part_b), _stream())
    check(rc, "lns_group_norm_bwd")
    _state.launches += 1
    if dgamma is not None:
        for part, grad in ((part_g, dgamma), (part_b, dbeta)):
            rc = _C.lib().lns_batch_sum_accum(_ptr(part), x.B, x.C, float(out_scale), _ptr(out_scale_dev), _ptr(grad), _stream())
            check(rc, "lns_batch_sum_accum")
            _state.launches += 1
    return out


# ---- misc ----------------------------------------------------------------------------------------------------
def fourier_embedding(param, dim, max_period=10000.0):
    """param: torch fp32 [B] on CUDA -> torch fp32 [B, dim]"""
    _need_cuda(param, "fourier_embedding")
    p = param.detach().float().contiguous()
    out = torch.empty(p.shape[0], dim, dtype=torch.float32, device=p.device)
    rc = _C.lib().lns_fourier_embedding(_ptr(p), p.shape[0], dim, float(max_period), _ptr(out), _stream())
    check(rc, "lns_fourier_embedding")
    _state.launches += 1
    return out


def rows_act(t):
    """torch fp32 [B, C] -> Act [B,1,1,C] (no copy)"""
    t = t.contiguous()
    return Act(t, t.shape[0], 1, 1, t.shape[1])


def spectral_conv2d(x, w_modes, m1, m2, Co, emb=None):
    """x: Act NHWC [B,H,W,Ci]; w_modes: fp32 [2,m1,m2,Ci,Co,2]; emb: fp32 [B,m1,m2,2,2] or None -> Act fp32 [B,H,W,Co]"""
    assert x.contiguous and x.layout == NHWC
    nbytes = _C.lib().lns_spectral_work_bytes(x.B, x.H, x.W, x.C, Co, m1, m2)
    work = torch.empty(nbytes, dtype=torch.uint8, device=x.t.device)
    out = Act.empty(x.B, x.H, x.W, Co, torch.float32, x.t.device)
    rc = _C.lib().lns_spectral_conv2d(_ptr(x.t), x.dtype, x.B, x.H, x.W, x.C, Co, m1, m2, _ptr(w_modes), _ptr(emb),
                                      _ptr(work), _ptr(out.t), _stream())
    check(rc, "lns_spectral_conv2d")
    _state.launches += 3
    return out
