"""ctypes binding of libzs3b200.so (the C ABI declared in include/zs3b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C zs3_b200/csrc``.  There is no
fallback: if the shared object is missing or a call fails, a ``Zs3NativeError`` is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libzs3b200.so")

ZS3_MAX_SEGMENTS = 6


class Zs3NativeError(RuntimeError):
    pass


class ConvSegment(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("x_cstride", C.c_int),
        ("w", C.c_void_p),
        ("cin_pad", C.c_int),
    ]


class ConvArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("Ho", C.c_int), ("Wo", C.c_int),
        ("R", C.c_int), ("S", C.c_int),
        ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("cout_pad", C.c_int),
        ("num_segments", C.c_int),
        ("seg", ConvSegment * ZS3_MAX_SEGMENTS),
        ("y", C.c_void_p),
        ("y_cstride", C.c_int),
        ("y_sp_stride", C.c_int),
        ("y_H", C.c_int), ("y_W", C.c_int),
        ("y_is_f32", C.c_int),
        ("accumulate", C.c_int),
        ("bias", C.c_void_p),
        ("stat_sum", C.c_void_p),
        ("stat_sqsum", C.c_void_p),
        ("ep_scale", C.c_void_p), ("ep_shift", C.c_void_p), ("ep_residual", C.c_void_p),
        ("ep_res_cstride", C.c_int), ("ep_relu", C.c_int),
        ("w_forward_layout", C.c_int),
        ("pre_scale", C.c_void_p * ZS3_MAX_SEGMENTS), ("pre_shift", C.c_void_p * ZS3_MAX_SEGMENTS),
        ("pre_relu", C.c_int),
    ]


class WgradArgs(C.Structure):
    _fields_ = [
        ("N", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ho", C.c_int), ("Wo", C.c_int),
        ("R", C.c_int), ("S", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("x", C.c_void_p),
        ("x_cstride", C.c_int),
        ("cin_pad", C.c_int),
        ("dy", C.c_void_p),
        ("dy_cstride", C.c_int),
        ("cout_pad", C.c_int),
        ("dw", C.c_void_p),
        ("k_splits", C.c_int),
        ("dw_ld", C.c_longlong),
        ("dw_ci_offset", C.c_int),
        ("cout_valid", C.c_int),
        ("cin_valid", C.c_int),
    ]


_lib = None


def _declare(lib):
    vp, i, ll, f, d = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double
    lib.zs3_last_error.restype = C.c_char_p
    lib.zs3_last_error.argtypes = []
    lib.zs3_abi_version.restype = i
    lib.zs3_device_supported.restype = i
    lib.zs3_launch_count.restype = C.c_ulonglong
    lib.zs3_launch_count.argtypes = []
    lib.zs3_sizeof.restype = C.c_ulonglong
    lib.zs3_sizeof.argtypes = [i]
    sigs = {
        "zs3_conv_fprop": [C.POINTER(ConvArgs), vp],
        "zs3_conv_wgrad": [C.POINTER(WgradArgs), vp],
        "zs3_pack_weight": [vp, i, i, i, i, i, i, vp, i, i, i, vp],
        "zs3_unpack_wgrad": [vp, i, i, vp, i, i, i, i, i, i, i, vp],
    }
    # entry points added by later translation units register themselves in _EXTRA_SIGS
    sigs.update(_EXTRA_SIGS(vp, i, ll, f, d))
    for name, argtypes in sigs.items():
        fn = getattr(lib, name)
        fn.restype = i
        fn.argtypes = argtypes
    lib.zs3_gmmn_train_workspace_size.restype = C.c_ulonglong
    lib.zs3_gmmn_train_workspace_size.argtypes = [i, i, i, i]
    sigs["zs3_gmmn_train_workspace_size"] = [i, i, i, i]
    lib.zs3_upsample4_ce_bwd_workspace_size.restype = C.c_ulonglong
    lib.zs3_upsample4_ce_bwd_workspace_size.argtypes = [i, i, i, i]
    sigs["zs3_upsample4_ce_bwd_workspace_size"] = [i, i, i, i]
    lib.zs3_augment_workspace_size.restype = C.c_ulonglong
    lib.zs3_augment_workspace_size.argtypes = [i, i, i, i]
    sigs["zs3_augment_workspace_size"] = [i, i, i, i]
    return sigs


class BnApplyArgs(C.Structure):
    _fields_ = [
        ("y", C.c_void_p), ("y_cstride", C.c_int),
        ("residual", C.c_void_p), ("res_cstride", C.c_int),
        ("out", C.c_void_p), ("out_cstride", C.c_int),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("M", C.c_longlong), ("C", C.c_int),
        ("relu", C.c_int), ("drop_mode", C.c_int), ("drop_p", C.c_float),
        ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
        ("keep_mask", C.c_void_p),
        ("offset_dev", C.c_void_p),
        ("stat_sum", C.c_void_p), ("stat_sqsum", C.c_void_p), ("count", C.c_longlong),
        ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float), ("momentum", C.c_float),
        ("running_mean", C.c_void_p), ("running_var", C.c_void_p), ("C_real", C.c_int),
        ("mean_out", C.c_void_p), ("invstd_out", C.c_void_p), ("scale_out", C.c_void_p), ("shift_out", C.c_void_p),
        ("reset_sum", C.c_void_p), ("reset_sqsum", C.c_void_p), ("reset_count", C.c_int),
        ("relu_mask_out", C.c_void_p),
        ("sync_clamp", C.c_int),
    ]


class BnBwdArgs(C.Structure):
    _fields_ = [
        ("dout", C.c_void_p), ("dout_cstride", C.c_int),
        ("out", C.c_void_p), ("out_cstride", C.c_int),
        ("y", C.c_void_p), ("y_cstride", C.c_int),
        ("mean", C.c_void_p), ("invstd", C.c_void_p), ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("M", C.c_longlong), ("C", C.c_int),
        ("relu", C.c_int), ("grad_scale", C.c_float), ("training", C.c_int),
        ("sum_dz", C.c_void_p), ("sum_dzx", C.c_void_p),
        ("dy", C.c_void_p), ("dy_cstride", C.c_int), ("dy_sp_stride", C.c_int),
        ("sp_Ho", C.c_int), ("sp_Wo", C.c_int), ("dy_H", C.c_int), ("dy_W", C.c_int),
        ("dres", C.c_void_p), ("dres_cstride", C.c_int), ("dres_accumulate", C.c_int),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("C_real", C.c_int), ("param_accumulate", C.c_int),
        ("reset_sum_dz", C.c_void_p), ("reset_sum_dzx", C.c_void_p), ("reset_count", C.c_int),
        ("relu_mask", C.c_void_p),
        ("stat_count", C.c_longlong),
    ]


class RowSource(C.Structure):
    _fields_ = [("base", C.c_void_p), ("rows", C.c_void_p), ("row_stride", C.c_longlong), ("col_stride", C.c_longlong)]


class GmmnItem(C.Structure):
    _fields_ = [
        ("emb", RowSource), ("noise", RowSource), ("real", RowSource),
        ("keep_mask", C.c_void_p), ("keep_rows", C.c_void_p), ("adj", C.c_void_p), ("out", C.c_void_p),
        ("rows", C.c_int), ("flags", C.c_int),
    ]


class GmmnTrainArgs(C.Structure):
    _fields_ = [
        ("items", C.c_void_p), ("n_items", C.c_int), ("max_rows", C.c_int),
        ("embed_dim", C.c_int), ("noise_dim", C.c_int), ("hidden", C.c_int), ("feat", C.c_int),
        ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
        ("apply_adam", C.c_int),
        ("adam_m", C.c_void_p * 4), ("adam_v", C.c_void_p * 4), ("grad", C.c_void_p * 4),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("step0", C.c_longlong),
        ("sigma", C.c_float * 8), ("nsigma", C.c_int),
        ("slope", C.c_float), ("drop_p", C.c_float),
        ("seed", C.c_ulonglong), ("offset", C.c_ulonglong),
        ("losses", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_ulonglong),
        ("phase_stamps", C.c_void_p),
        ("weights_in_out", C.c_int),
    ]


class ComponentsArgs(C.Structure):
    _fields_ = [
        ("labels", C.c_void_p), ("src_index", C.c_void_p), ("image_stride", C.c_longlong),
        ("B", C.c_int), ("h", C.c_int), ("w", C.c_int), ("max_nodes", C.c_int),
        ("n_nodes", C.c_void_p), ("node_label", C.c_void_p), ("node_seed", C.c_void_p), ("node_map", C.c_void_p),
        ("adj", C.c_void_p),
    ]


class AugItem(C.Structure):
    _fields_ = [
        ("image", C.c_void_p), ("label", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("flip", C.c_int),
        ("rw", C.c_int), ("rh", C.c_int), ("x1", C.c_int), ("y1", C.c_int), ("blur_radius", C.c_float),
    ]


class AugmentArgs(C.Structure):
    _fields_ = [
        ("items", C.c_void_p), ("items_host", C.POINTER(AugItem)), ("n", C.c_int), ("max_src_h", C.c_int),
        ("out_w", C.c_int), ("out_h", C.c_int), ("fill_label", C.c_int),
        ("lut", C.c_void_p), ("out_image", C.c_void_p), ("out_label", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_ulonglong),
    ]


class SgemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("lda", C.c_longlong), ("transA", C.c_int), ("idxA", C.c_void_p),
        ("B", C.c_void_p), ("ldb", C.c_longlong), ("transB", C.c_int), ("idxB", C.c_void_p),
        ("C", C.c_void_p), ("ldc", C.c_longlong),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("bias", C.c_void_p), ("accumulate", C.c_int),
        ("dyn_count", C.c_void_p), ("dyn_dim", C.c_int),
    ]


class BnActF32Args(C.Structure):
    _fields_ = [
        ("y", C.c_void_p), ("y_cstride", C.c_longlong),
        ("residual", C.c_void_p), ("res_cstride", C.c_longlong),
        ("scale", C.c_void_p), ("shift", C.c_void_p),
        ("out", C.c_void_p), ("out_cstride", C.c_longlong),
        ("pieces", C.c_void_p * 3), ("n_pieces", C.c_int), ("piece_cstride", C.c_longlong),
        ("M", C.c_longlong), ("C", C.c_int),
        ("relu", C.c_int),
        ("drop_mode", C.c_int), ("drop_p", C.c_float),
        ("seed", C.c_ulonglong), ("offset", C.c_ulonglong), ("offset_dev", C.c_void_p), ("keep_mask", C.c_void_p),
    ]


class BnBwdF32Args(C.Structure):
    _fields_ = [
        ("dout", C.c_void_p), ("dout_cstride", C.c_longlong),
        ("act", C.c_void_p), ("act_cstride", C.c_longlong),
        ("act_hi", C.c_void_p), ("act_hi_cstride", C.c_longlong),
        ("y", C.c_void_p), ("y_cstride", C.c_longlong),
        ("mean", C.c_void_p), ("invstd", C.c_void_p), ("scale", C.c_void_p),
        ("M", C.c_longlong), ("C", C.c_int),
        ("relu", C.c_int), ("grad_scale", C.c_float), ("training", C.c_int),
        ("sum_dz", C.c_void_p), ("sum_dzx", C.c_void_p),
        ("dy", C.c_void_p), ("dy_cstride", C.c_longlong),
        ("dy_pieces", C.c_void_p * 3), ("n_pieces", C.c_int), ("piece_cstride", C.c_longlong),
        ("dres", C.c_void_p), ("dres_cstride", C.c_longlong),
        ("dgamma", C.c_void_p), ("dbeta", C.c_void_p), ("C_real", C.c_int), ("param_accumulate", C.c_int),
    ]


def _EXTRA_SIGS(vp, i, ll, f, d):
    return {
        "zs3_debug_im2col_probe": [vp, i, i, i, i, i, i, i, i, i, i, i, i, i, i, i, vp, vp],
        "zs3_nchw_f32_to_nhwc_bf16": [vp, vp, i, i, ll, i, vp],
        "zs3_nhwc_bf16_to_nchw_f32": [vp, vp, i, i, ll, i, vp],
        "zs3_bn_finalize": [vp, vp, ll, vp, vp, f, f, vp, vp, vp, vp, vp, vp, i, i, i, vp],
        "zs3_bn_eval_coeffs": [vp, vp, vp, vp, f, vp, vp, vp, vp, i, i, vp],
        "zs3_bn_stats": [vp, i, ll, i, vp, vp, vp],
        "zs3_bn_apply": [C.POINTER(BnApplyArgs), vp],
        "zs3_bn_bwd_reduce": [C.POINTER(BnBwdArgs), vp],
        "zs3_bn_bwd_apply": [C.POINTER(BnBwdArgs), vp],
        "zs3_stem_im2col": [vp, vp, i, i, i, i, i, i, i, i, i, i, i, vp],
        "zs3_maxpool_fwd": [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_maxpool_bwd": [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_bilinear_fwd": [vp, vp, i, i, i, i, i, i, i, i, vp],
        "zs3_bilinear_bwd": [vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_upsample_logits_fwd": [vp, vp, i, i, i, i, i, i, i, vp],
        "zs3_upsample_logits_bwd": [vp, vp, i, i, i, i, i, i, i, vp],
        "zs3_spatial_sum": [vp, vp, i, i, i, i, i, f, vp],
        "zs3_spatial_broadcast": [vp, vp, i, i, i, i, i, f, i, vp],
        "zs3_ce_fwd": [vp, vp, vp, i, i, ll, i, f, vp, vp, vp],
        "zs3_ce_bwd": [vp, vp, vp, i, i, ll, i, f, vp, vp, vp, vp],
        "zs3_upsample_ce_fwd": [vp, vp, vp, i, i, i, i, i, i, i, i, f, vp, vp, vp],
        "zs3_upsample_ce_bwd": [vp, vp, vp, i, i, i, i, i, i, i, i, f, vp, vp, vp, vp],
        "zs3_upsample4_ce_bwd": [vp, vp, vp, i, i, i, i, i, i, i, i, f, vp, vp, vp, vp, C.c_ulonglong, vp],
        "zs3_cast_f32_to_bf16": [vp, vp, ll, vp],
        "zs3_sgd_step": [vp, vp, vp, ll, f, f, f, i, i, f, vp],
        "zs3_sgd_step_lrdev": [vp, vp, vp, ll, vp, f, f, i, i, f, vp],
        "zs3_adam_step": [vp, vp, vp, vp, ll, f, f, f, f, i, f, vp],
        "zs3_sgemm": [C.POINTER(SgemmArgs), vp],
        "zs3_find_active_rows": [vp, i, i, vp, vp, vp],
        "zs3_col_sum": [vp, ll, vp, vp, i, i, vp, i, vp],
        "zs3_leaky_dropout_fwd": [vp, vp, ll, f, i, f, C.c_ulonglong, C.c_ulonglong, vp, vp],
        "zs3_leaky_dropout_bwd": [vp, vp, vp, i, i, vp, vp, f, f, vp],
        "zs3_mmd_fwd": [vp, vp, i, i, i, C.POINTER(C.c_float), i, vp, vp, vp, vp],
        "zs3_mmd_bwd": [vp, vp, i, i, i, vp, vp, vp, vp, vp, vp],
        "zs3_concat2": [vp, i, vp, i, vp, ll, vp],
        "zs3_gmmn_train_fused": [C.POINTER(GmmnTrainArgs), vp],
        "zs3_label_components": [C.POINTER(ComponentsArgs), vp],
        "zs3_debug_wgrad_tile": [ll, i, i, i, C.POINTER(C.c_int), C.POINTER(C.c_int)],
        "zs3_augment_batch": [C.POINTER(AugmentArgs), vp],
        "zs3_argmax_confusion": [vp, vp, i, i, ll, vp, vp, vp],
        "zs3_confusion_from_pred": [vp, vp, ll, i, vp, vp],
        "zs3_split3_f32": [vp, vp, vp, vp, ll, vp],
        "zs3_pack_weight_component": [vp, i, i, i, i, i, i, vp, i, i, i, i, vp],
        "zs3_bn_apply_f32": [vp, i, vp, i, vp, i, vp, vp, ll, i, i, vp],
        "zs3_maxpool_f32": [vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_bilinear_f32": [vp, vp, i, i, i, i, i, i, i, i, vp],
        "zs3_spatial_sum_f32": [vp, vp, i, i, i, f, vp],
        "zs3_spatial_broadcast_f32": [vp, vp, i, i, i, vp],
        "zs3_nchw_to_nhwc_f32": [vp, vp, i, i, ll, i, vp],
        "zs3_stem_im2col_f32": [vp, vp, i, i, i, i, i, i, i, i, i, i, i, vp],
        "zs3_split_f32": [vp, C.POINTER(C.c_void_p), i, ll, vp],
        "zs3_bn_act_f32": [C.POINTER(BnActF32Args), vp],
        "zs3_bn_bwd_f32": [C.POINTER(BnBwdF32Args), vp],
        "zs3_channel_sums_f32": [vp, ll, ll, i, vp, vp],
        "zs3_maxpool_arg_f32": [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_maxpool_bwd_f32": [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_bilinear_bwd_f32": [vp, vp, i, i, i, i, i, i, i, i, i, vp],
        "zs3_spatial_broadcast_acc_f32": [vp, vp, i, i, i, f, i, vp],
    }


# ids of include/zs3b200.h (ZS3_STRUCT_*) -> ctypes mirror; checked against zs3_sizeof() when the library loads
STRUCT_IDS = {1: ConvArgs, 2: WgradArgs, 3: BnApplyArgs, 4: BnBwdArgs, 5: SgemmArgs, 6: GmmnItem, 7: GmmnTrainArgs,
              8: ComponentsArgs, 9: ConvSegment, 10: RowSource, 11: BnActF32Args, 12: BnBwdF32Args,
              13: AugItem, 14: AugmentArgs}


def lib():
    """Load (once) and return the native library; raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Zs3NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or PyTorch fallback for the zs3_b200 kernels)")
        l = C.CDLL(LIB_PATH)
        declared = _declare(l)
        l._zs3_declared = sorted(declared) + ["zs3_last_error", "zs3_abi_version", "zs3_device_supported",
                                                "zs3_launch_count", "zs3_sizeof"]
        for which, mirror in STRUCT_IDS.items():
            if l.zs3_sizeof(which) != C.sizeof(mirror):
                raise Zs3NativeError(f"ctypes mirror {mirror.__name__} is {C.sizeof(mirror)} bytes, the library's "
                                     f"struct is {l.zs3_sizeof(which)}: binding and libzs3b200.so are out of step")
        _lib = l
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().zs3_last_error().decode("utf-8", "replace")
        raise Zs3NativeError(f"{what} failed with code {rc}: {msg}")


_raw_stream = None


def stream_ptr():
    """cudaStream_t of torch's CURRENT stream on the current device (follows torch.cuda.stream(...) and graph capture).
    Called once per kernel launch: torch.cuda.current_stream() costs ~19 us of Python per call (measured on the box:
    3 ms per config-5 step, tools/profile_config5_host.py), the raw accessor well under 1 us."""
    global _raw_stream
    if _raw_stream is None:
        import torch
        get_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
        get_dev = getattr(torch._C, "_cuda_getDevice", None)
        if get_stream is not None and get_dev is not None:
            _raw_stream = lambda: get_stream(get_dev())  # noqa: E731
        else:
            _raw_stream = lambda: torch.cuda.current_stream().cuda_stream  # noqa: E731
    return C.c_void_p(_raw_stream())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())
