"""ctypes binding of libgwd_b200.so (the C ABI declared in include/gwd_b200.h).

The product path has NO fallback: if the shared library is missing or a call
fails, an exception is raised.  Build it with `python __graft_entry__.py build`
(nvcc -gencode arch=compute_100a,code=sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgwd_b200.so")

ACT_NONE, ACT_RELU, ACT_GELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3, 4
RES_NONE, RES_BEFORE_NORM, RES_AFTER, RES_MUL_ACTGRAD = 0, 1, 2, 3

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int32
c_i64 = ctypes.c_int64
c_float = ctypes.c_float


class GemmDesc(ctypes.Structure):
    """struct gwd_gemm_desc (include/gwd_b200.h)"""
    _fields_ = [
        ("x", c_void_p), ("B", c_int), ("H", c_int), ("W", c_int),
        ("x_cstride", c_int), ("x_coff", c_int), ("cin", c_int),
        ("w", c_void_p), ("taps", c_int), ("n_pad", c_int), ("n", c_int),
        ("bias", c_void_p), ("ln_g", c_void_p), ("ln_b", c_void_p), ("ln_eps", c_float),
        ("pre_act", c_int), ("post_act", c_int), ("out_scale", c_float),
        ("res", c_void_p), ("res_cstride", c_int), ("res_coff", c_int), ("res_mode", c_int),
        ("y", c_void_p), ("y_cstride", c_int), ("y_coff", c_int), ("y_f32", c_int),
        ("y_raw", c_void_p), ("yraw_cstride", c_int), ("yraw_coff", c_int),
        ("store_n", c_int), ("w_per_image", c_int), ("upsample2", c_int),
        ("x_wstride", ctypes.c_int64), ("x_hstride", ctypes.c_int64), ("x_bstride", ctypes.c_int64),
        ("ag_act", c_int), ("ag_from_input", c_int), ("ag_y_mul", c_float), ("ag_scale", c_float),
    ]


class AttnDesc(ctypes.Structure):
    """struct gwd_attn_desc (include/gwd_b200.h)"""
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("o", c_void_p),
        ("items", c_int), ("heads", c_int), ("Lq", c_int), ("Lk", c_int), ("hd", c_int),
        ("q_item_stride", c_i64), ("q_row_stride", c_i64), ("k_item_stride", c_i64), ("k_row_stride", c_i64),
        ("v_item_stride", c_i64), ("v_row_stride", c_i64), ("o_item_stride", c_i64), ("o_row_stride", c_i64),
        ("bias", c_void_p), ("mask", c_void_p), ("mask_windows", c_int), ("key_padding", c_void_p),
        ("scale", c_float),
        ("dropout_seed", c_void_p), ("dropout_site", ctypes.c_uint32), ("dropout_p", c_float),
    ]


class AttnBwdDesc(ctypes.Structure):
    """struct gwd_attn_bwd_desc (include/gwd_b200.h)"""
    _fields_ = [
        ("q", c_void_p), ("k", c_void_p), ("v", c_void_p), ("d_o", c_void_p),
        ("dq", c_void_p), ("dk", c_void_p), ("dv", c_void_p),
        ("items", c_int), ("heads", c_int), ("Lq", c_int), ("Lk", c_int), ("hd", c_int),
        ("q_item_stride", c_i64), ("q_row_stride", c_i64), ("k_item_stride", c_i64), ("k_row_stride", c_i64),
        ("v_item_stride", c_i64), ("v_row_stride", c_i64), ("do_item_stride", c_i64), ("do_row_stride", c_i64),
        ("dq_item_stride", c_i64), ("dq_row_stride", c_i64), ("dk_item_stride", c_i64), ("dk_row_stride", c_i64),
        ("dv_item_stride", c_i64), ("dv_row_stride", c_i64),
        ("scale", c_float), ("o", c_void_p), ("o_item_stride", c_i64), ("o_row_stride", c_i64),
        ("dq_mul", c_float), ("dk_mul", c_float),
        ("dropout_seed", c_void_p), ("dropout_site", ctypes.c_uint32), ("dropout_p", c_float),
        ("key_padding", c_void_p), ("stats_ws", c_void_p),
    ]


P, I, L, F_ = c_void_p, c_int, c_i64, c_float
# name -> (restype, argtypes); must list every symbol of include/gwd_b200.h
SIGNATURES = {
    "gwd_last_error": (ctypes.c_char_p, []),
    "gwd_version": (c_int, []),
    "gwd_launch_count": (c_i64, []),
    "gwd_reset_launch_count": (None, []),
    "gwd_conv_gemm": (c_int, [ctypes.POINTER(GemmDesc), P]),
    "gwd_attention": (c_int, [ctypes.POINTER(AttnDesc), P]),
    "gwd_token_attention": (c_int, [P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, F_, P]),
    "gwd_ref_scores": (c_int, [P, L, P, L, P, I, I, I, I, I, I, F_, P]),
    "gwd_ref_diffuse": (c_int, [P, P, P, P, P, P, I, I, I, I, P]),
    "gwd_ref_requery": (c_int, [P, P, L, P, L, I, I, I, I, I, I, F_, P]),
    "gwd_layernorm": (c_int, [P, L, P, L, P, P, F_, I, P, L, L, I, I, P]),
    "gwd_add_rows": (c_int, [P, L, P, L, L, P, L, L, I, P]),
    "gwd_window_gather": (c_int, [P, L, P, P, F_, P, L, I, I, I, I, I, I, I, P]),
    "gwd_window_merge": (c_int, [P, L, P, L, P, L, P, P, F_, P, L, I, I, I, I, I, I, I, P]),
    "gwd_upsample_nearest": (c_int, [P, L, I, I, I, P, L, I, I, I, P, L, P]),
    "gwd_avgpool": (c_int, [P, L, I, I, I, I, P, L, I, P]),
    "gwd_bilinear_up": (c_int, [P, L, I, I, I, P, L, I, I, I, P]),
    "gwd_avgpool_pyramid": (c_int, [P, L, I, I, I, P, P, P, P, I, P]),
    "gwd_bilinear_up4": (c_int, [P, P, P, P, P, I, P, L, I, I, I, P]),
    "gwd_sample_bilinear": (c_int, [P, L, I, P, L, I, I, I, I, P, I, P, P]),
    "gwd_sample_scalar": (c_int, [P, I, I, I, P, I, P, P]),
    "gwd_line_ref_gather": (c_int, [P, L, P, L, P, I, P, L, I, I, I, I, I, I, P]),
    "gwd_anchor_mix": (c_int, [P, L, P, I, L, I, P, P]),
    "gwd_nchw_to_nhwc": (c_int, [P, I, I, L, P, I, P]),
    "gwd_pil_bilinear_ksize": (c_int, [I, I]),
    "gwd_pil_bilinear_coeffs": (c_int, [I, I, P, P, P]),
    "gwd_pil_nearest_index": (c_int, [I, I, P]),
    "gwd_resample_u8": (c_int, [P, L, I, I, I, P, I, I, P, P, P, I, I, P]),
    "gwd_gather2d": (c_int, [P, L, I, I, I, P, I, I, P, P, I, I, P]),
    "gwd_jitter_u8": (c_int, [P, L, I, P, P, P, P, P]),
    "gwd_images_to_batch": (c_int, [P, I, I, I, ctypes.POINTER(c_float), ctypes.POINTER(c_float), P, P, P]),
    "gwd_stem_conv_pool": (c_int, [P, P, P, P, I, I, I, P]),
    "gwd_certain_sample": (c_int, [P, I, I, P, I, I, I, I, ctypes.POINTER(c_float), I, P, P, P]),
    "gwd_match_cost": (c_int, [P, P, P, P, P, I, I, I, I, F_, F_, P, P, P]),
    "gwd_lsap_batch": (c_int, [P, P, P, I, I, P, P, P, I]),
    "gwd_depth_metrics": (c_int, [P, P, I, L, F_, F_, P, P, P]),
    "gwd_seg_confusion": (c_int, [P, L, L, L, P, I, L, I, I, P, P]),
    "gwd_silog_sums": (c_int, [P, I, I, I, P, I, I, F_, F_, I, P, P]),
    "gwd_layernorm_bwd": (c_int, [P, L, P, L, P, P, I, F_, P, L, P, L, P, P, L, I, I, P]),
    "gwd_act_bwd": (c_int, [P, I, L, P, I, L, I, P, L, L, I, I, F_, F_, I, P]),
    "gwd_transpose": (c_int, [P, L, P, L, L, L, I, P, P]),
    "gwd_transpose_batch": (c_int, [P, P, I, I, P]),
    "gwd_linear_wgrad": (c_int, [P, L, P, L, L, I, I, P, L, P, P]),
    "gwd_conv3x3_wgrad": (c_int, [P, L, P, L, I, I, I, I, I, P, P, P]),
    "gwd_attention_bwd": (c_int, [ctypes.POINTER(AttnBwdDesc), P]),
    "gwd_set_loss": (c_int, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, P, P, P, P]),
    "gwd_sumsq": (c_int, [P, L, P, P]),
    "gwd_adamw_step": (c_int, [P, P, P, P, P, L, F_, F_, F_, F_, F_, I, F_, F_, P, P]),
    "gwd_silog_bwd": (c_int, [P, I, I, I, P, I, I, F_, F_, I, P, F_, F_, F_, P, I, P, P]),
    "gwd_seg_ce": (c_int, [P, L, L, L, P, I, L, I, I, F_, P, P, I, P, P]),
    "gwd_bilinear_up_bwd": (c_int, [P, L, I, I, I, P, L, I, I, I, P]),
    "gwd_avgpool_bwd": (c_int, [P, L, I, F_, P, L, P, L, I, I, I, I, P]),
    "gwd_anchor_mix_bwd": (c_int, [P, L, P, P, I, L, I, I, P, L, P, P]),
    "gwd_sample_bilinear_bwd": (c_int, [P, P, I, P, L, I, I, I, I, P]),
    "gwd_sample_scalar_bwd": (c_int, [P, P, I, P, P, I, I, I, P]),
    "gwd_window_attention_bwd": (c_int, [P, L, P, L, P, L, P, P, I, P, I, I, I, I, F_, P]),
    "gwd_token_attention_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, L, L, L, L, L, L, L, F_, P]),
    "gwd_ref_diffuse_dev": (c_int, [P, P, P, P, P, I, I, I, I, P]),
    "gwd_ref_diffuse_conv_dev": (c_int, [P, P, P, P, P, I, I, I, I, P]),
    "gwd_diffuse_filter_pack": (c_int, [P, P, P, P, P]),
    "gwd_ref_diffuse_bwd": (c_int, [P, P, P, P, P, P, P, P, P, P, I, I, I, I, P]),
    "gwd_ref_affine": (c_int, [P, L, P, P, P, L, I, P]),
    "gwd_ref_affine_bwd": (c_int, [P, P, L, P, P, P, P, I, I, P]),
    "gwd_ref_requery_bwd": (c_int, [P, P, L, P, L, P, P, L, I, I, I, I, I, F_, P]),
    "gwd_ref_scores_bwd": (c_int, [P, P, L, P, L, P, L, P, L, I, I, I, I, I, F_, P]),
    "gwd_line_ref_scatter": (c_int, [P, L, P, I, P, L, I, I, I, I, I, I, P]),
    "gwd_subsample2": (c_int, [P, P, I, I, I, I, P]),
    "gwd_zero_stuff2": (c_int, [P, P, P, I, I, I, I, P]),
    "gwd_scale_rows": (c_int, [P, P, L, P]),
    "gwd_fold_mirror": (c_int, [P, P, P, L, P]),
    "gwd_im2col3x3_s2": (c_int, [P, P, I, I, I, I, P]),
    "gwd_col2im3x3_s2": (c_int, [P, P, P, I, I, I, I, P]),
    "gwd_select_lines": (c_int, [P, I, P, I, I, I, I, I, P, P, P]),
    "gwd_dropout": (c_int, [P, P, P, L, P, ctypes.c_uint32, F_, P]),
}

_lib = None


class GwdError(RuntimeError):
    pass


def lib():
    """Load the shared library once; fail loudly when it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GwdError(
                "libgwd_b200.so is not built (%s). Run `python __graft_entry__.py build`; "
                "there is no CPU or PyTorch fallback for the GW-Depth hot path." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().gwd_last_error()
        raise GwdError("%s failed (%d): %s" % (what, status, msg.decode() if msg else ""))


def launch_count():
    return int(lib().gwd_launch_count())


def reset_launch_count():
    lib().gwd_reset_launch_count()
