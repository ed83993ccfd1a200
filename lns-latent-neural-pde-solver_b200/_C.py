"""ctypes binding of liblns_b200.so (declared in include/lns_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  The library is built by
``lns_b200.build.build()`` (``python __graft_entry__.py`` does that)."""
import ctypes
import os

from .build import LIB

i32, i64, f32 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float
vp = ctypes.c_void_p


class LnsError(RuntimeError):
    pass


class ConvDesc(ctypes.Structure):
    """Mirror of LnsConvDesc (include/lns_b200.h)."""
    _fields_ = [
        ("x", vp), ("x_dtype", i32), ("x_layout", i32),
        ("B", i32), ("Hin", i32), ("Win", i32), ("Cin", i32),
        ("x_bstride", i64),
        ("Hv", i32), ("Wv", i32),
        ("KH", i32), ("KW", i32), ("stride", i32), ("dil", i32), ("pad_t", i32), ("pad_l", i32),
        ("pad_mode_h", i32), ("pad_mode_w", i32),
        ("w", vp), ("w_format", i32), ("engine", i32),
        ("bias", vp), ("sample_bias", vp),
        ("pro_scale", vp), ("pro_shift", vp), ("pro_act", i32),
        ("act", i32),
        ("pre_add", vp), ("pre_add_dtype", i32), ("pre_add_bstride", i64),
        ("residual", vp), ("res_dtype", i32), ("res_bstride", i64),
        ("y", vp), ("y_dtype", i32), ("y_layout", i32),
        ("Hout", i32), ("Wout", i32), ("Cout", i32),
        ("y_bstride", i64),
        ("stats", vp),
    ]


# name -> (restype, argtypes); every symbol include/lns_b200.h declares
SIGNATURES = {
    "lns_version": (ctypes.c_char_p, []),
    "lns_last_error": (ctypes.c_char_p, []),
    "lns_device_info": (i32, [ctypes.POINTER(i32)] * 3),
    "lns_conv2d": (i32, [ctypes.POINTER(ConvDesc), vp]),
    "lns_conv_stats_chunks": (i32, [i32, i32]),
    "lns_packed_weight_bytes": (i64, [i32, i32, i32, i32, i32]),
    "lns_pack_conv_weight": (i32, [vp, i32, i32, i32, i32, i32, vp, vp]),
    "lns_chan_stats_chunks": (i32, [i32, i32]),
    "lns_chan_stats": (i32, [vp, i32, i32, i32, i32, i32, i64, vp, vp]),
    "lns_norm_finalize": (i32, [vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
    "lns_norm_finalize_centred": (i32, [vp, i32, i32, i32, i32, i32, f32, vp, vp, vp, vp, vp, vp]),
    "lns_group_norm_affine": (i32, [vp, i32, i32, i32, i32, i32, i64, i32, f32, vp, vp, vp, vp, vp, vp, vp]),
    "lns_group_norm_act_supported": (i32, [i32, i32, i32]),
    "lns_group_norm_act": (i32, [vp, i32, i32, i32, i32, i32, i64, i32, f32, vp, vp, vp, i32, vp, i32, i64, vp]),
    "lns_pointwise_proj": (i32, [vp, i32, i32, i32, i32, i64, vp, vp, i32, vp, vp, i32, vp, i64, vp]),
    "lns_pointwise_proj_steps": (i32, [vp, i32, i32, i32, i32, i64, vp, vp, i32, vp, vp, i32, vp, i64, i32, i64, vp]),
    "lns_affine_act": (i32, [vp, i32, i64, i32, i32, i32, vp, vp, i32, vp, i32, i64, vp]),
    "lns_layernorm": (i32, [vp, i32, i32, i32, i32, vp, vp, f32, vp, vp, i32, vp]),
    "lns_attention": (i32, [vp, i32, i32, i32, i32, i32, f32, vp, i32, vp]),
    "lns_axis_mean": (i32, [vp, i32, i32, i32, i32, i32, i64, i32, vp, vp]),
    "lns_lowrank_kernel": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, f32, vp, vp]),
    "lns_axial_contract": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
    "lns_fablock_core_supported": (i32, [i32, i32, i32, i32]),
    "lns_fablock_prepass": (i32, [vp, i32, i32, i32, i32, i32, i64, f32, vp, vp, vp, vp, vp, vp, vp]),
    "lns_fablock_core": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp]),
    "lns_fablock_prepass_staged": (i32, [vp, i32, i32, i32, i32, i32, i64, f32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "lns_sablock_fused_supported": (i32, [i32, i32, i32, i32]),
    "lns_sablock_fused": (i32, [vp, i32, i32, i32, i32, vp, vp, f32, vp, vp, vp, vp, vp, f32, vp, vp]),
    "lns_ffn_fused_supported": (i32, [i32, i32]),
    "lns_ffn_fused": (i32, [vp, i32, i32, i32, i32, i64, vp, vp, vp, vp, vp, i64, vp]),
    "lns_fa_axis_kernel_supported": (i32, [i32, i32, i32, i32, i32, i32]),
    "lns_fa_axis_kernel": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, f32, vp, vp, vp, vp, vp, vp, f32, vp, vp]),
    "lns_fablock_full_supported": (i32, [i32, i32, i32, i32, i32]),
    "lns_fablock_full": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]),
    "lns_fablock_full_staged_supported": (i32, [i32, i32, i32, i32, i32]),
    "lns_fablock_full_staged": (i32, [vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, f32, vp, vp, vp, vp]),
    "lns_fablock_tc_supported": (i32, [i32, i32, i32, i32, i32]),
    "lns_fablock_tc": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp]),
    "lns_nchw_to_nhwc": (i32, [vp, i32, i32, i32, i32, i64, vp, i32, i64, vp]),
    "lns_nhwc_to_nchw": (i32, [vp, i32, i32, i32, i32, i32, i64, vp, i64, vp]),
    "lns_frame_sums": (i32, [vp, vp, i64, i32, vp, vp]),
    "lns_frame_sums_denorm": (i32, [vp, vp, i64, i32, i32, i32, vp, vp, vp, f32, f32, vp, vp]),
    "lns_fourier_embedding": (i32, [vp, i32, i32, f32, vp, vp]),
    "lns_channel_gate": (i32, [vp, i32, i32, i32, i32, vp, vp, i32, vp]),
    "lns_spectral_work_bytes": (i64, [i32, i32, i32, i32, i32, i32, i32]),
    "lns_spectral_conv2d": (i32, [vp] + [i32] * 8 + [vp] * 5),
    "lns_conv2d_wgrad_work_bytes": (i64, [i32] * 7),
    "lns_conv2d_wgrad": (i32, [vp, i64, vp, vp, i32, vp, i64] + [i32] * 13 + [f32, vp, vp, vp, vp]),
    "lns_chan_sum_slices": (i32, [i32]),
    "lns_chan_sum_accum": (i32, [vp, i64, i32, i32, i32, f32, vp, vp, vp, vp]),
    "lns_act_bwd": (i32, [vp, vp, i64, i32, vp, vp]),
    "lns_group_norm_bwd": (i32, [vp, i64, vp, i64, vp, i64, i32, i32, i32, i32, f32, vp, vp, i64, vp, vp, vp]),
    "lns_batch_sum_accum": (i32, [vp, i32, i32, f32, vp, vp, vp]),
    "lns_loss_scale": (i32, [vp, f32, vp, vp]),
    "lns_scale_by": (i32, [vp, vp, i64, vp, vp]),
    "lns_absmax": (i32, [vp, i64, vp, vp]),
    "lns_pixel_dot": (i32, [vp, i64, vp, i64, i32, i32, i32, i32, vp, vp]),
    "lns_scale_add": (i32, [vp, vp, vp, i32, i32, i32, vp, vp]),
}

_lib = None


def lib():
    """The loaded library; raises if it has not been built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is None:
        path = os.environ.get("LNS_B200_LIB", LIB)  # (instrumented builds for tools/; the product loads the in-tree library)
        if not os.path.exists(path):
            raise LnsError(f"{path} is missing: build it with `python __graft_entry__.py` "
                           "(lns_b200 has no CPU or PyTorch-eager fallback)")
        l = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().lns_last_error().decode(errors="replace")
        raise LnsError(f"{what} failed (rc={rc}): {msg}")
