"""ctypes binding of libttdg_sm100.so (include/ttdg_b200.h).

There is NO fallback: if the library is missing, ``lib()`` raises, and every op in ``ttdg_b200.ops`` raises with
it.  The library is built in-tree by ``ttdg_b200._build`` (``__graft_entry__.build()``)."""
import ctypes
import os

from ctypes import c_int, c_int64, c_uint64, c_float, c_double, c_void_p, c_char_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("TTDG_LIB") or os.path.join(_HERE, "lib", "libttdg_sm100.so")     # TTDG_LIB: A/B builds of the same ABI (tools/)
_lib = None

P = c_void_p
# name -> (restype, argtypes); mirrors include/ttdg_b200.h one to one
SIGNATURES = {
    "ttdg_version": (c_int, []),
    "ttdg_build_info": (c_char_p, []),
    "ttdg_limit": (c_int, [c_char_p]),
    "ttdg_launch_count": (ctypes.c_longlong, []),
    "ttdg_sgd_step": (c_int, [P, P, P, c_int64, c_float, c_float, c_float, c_float, c_int, P]),
    "ttdg_sinkhorn_small_fwd": (c_int, [P, P, P, c_int, c_int, c_double, c_int, c_int, P]),
    "ttdg_sinkhorn_small_bwd": (c_int, [P, P, P, P, c_int, c_int, c_double, c_int, c_int, P]),
    "ttdg_sinkhorn_stream_scratch_bytes": (c_int64, [c_int, c_int, c_int]),
    "ttdg_sinkhorn_stream_fwd": (c_int, [P, P, c_int, c_int, c_int, c_float, c_int, c_int, P, P]),
    "ttdg_lap_solve": (c_int, [P, P, P, c_int, P]),
    "ttdg_linear_f64acc": (c_int, [P, c_int, P, c_int, P, P, c_int, c_int, c_int, c_int, P]),
    "ttdg_gemm_f64acc": (c_int, [c_int, c_int, c_int, c_int, c_int, P, c_int, c_int, P, c_int, c_int, P, c_int, c_int,
                                 c_int, P]),
    "ttdg_attn_adjacency": (c_int, [P, P, c_int, c_int, c_float, P, P, c_float, c_uint64, c_uint64, P, P]),
    "ttdg_affinity_hidden": (c_int, [P, P, P, P, c_int, c_int, c_int, P, P]),
    "ttdg_affinity_pairs_fwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P, P]),
    "ttdg_affinity_bwd_scratch_bytes": (c_int64, [c_int, c_int]),
    "ttdg_affinity_pairs_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, c_int64, P, P, P, P, P]),
    "ttdg_gagm_scratch_bytes": (c_int64, [c_int, c_int]),
    "ttdg_gagm_set_lap_fast": (c_int, [c_int]),
    "ttdg_gagm_read_profile": (c_int, [P]),
    "ttdg_gagm_solve": (c_int, [P, P, P, P, c_int, c_int, c_int, c_double, c_double, c_double, c_int, c_int, c_double,
                                c_double, c_int, c_int, P, P, P, P, P, c_int, P]),
    "ttdg_matching_loss_scratch_bytes": (c_int64, [c_int]),
    "ttdg_matching_loss_fwd": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P]),
    "ttdg_matching_loss_bwd": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P]),
    "ttdg_focal_bce_scratch_bytes": (c_int64, []),
    "ttdg_focal_bce_fwd": (c_int, [P, P, c_int64, P, P, P]),
    "ttdg_focal_bce_bwd": (c_int, [P, P, c_int64, P, P, P]),
    "ttdg_conv_fwd": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_conv_dgrad": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_conv_wgrad": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_relu_bn_bwd": (c_int, [P, P, P, c_int, c_int64, P, P]),
    "ttdg_relu_bn_bwd2": (c_int, [P, P, P, c_int, c_int64, P, P, P]),
    "ttdg_bias_grad": (c_int, [P, c_int64, c_int, P, P]),
    "ttdg_maxpool3x3s2": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_resample2": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "ttdg_preprocess": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_float, P, P]),
    "ttdg_stem_tc": (c_int, [P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_conv_tc_supported": (c_int, [c_int, c_int, c_int]),
    "ttdg_conv_tc_set_cluster": (c_int, [c_int]),
    "ttdg_conv_tc_set_epilogue": (c_int, [c_int]),
    "ttdg_set_sm_limit": (c_int, [c_int]),
    "ttdg_conv_tc_set_trace": (c_int, [c_void_p, c_int]),
    "ttdg_conv_tc": (c_int, [P, P, P, P, P, P] + [c_int] * 15 + [P, P]),
    "ttdg_wgrad_tc_supported": (c_int, [c_int, c_int, c_int]),
    "ttdg_wgrad_tc": (c_int, [P, P] + [c_int] * 10 + [P, P]),
    "ttdg_tf32_split": (c_int, [P, P, P, c_int64, P]),
    "ttdg_weight_refresh": (c_int, [P, c_int, c_int64, P]),
    "ttdg_rpn_select": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, c_int, P, P, P, P, P]),
    "ttdg_sort_candidates": (c_int, [P, P, P, P, c_int, c_int, c_int, c_float, P, P, P, P, P]),
    "ttdg_rpn_nms_levels": (c_int, [P, P, P, c_int, c_int, c_float, c_int, P, P]),
    "ttdg_top_candidates": (c_int, [P, P, P, c_int, c_int, c_int, c_float, P, P, P, P]),
    "ttdg_resize_ksize": (c_int, [c_int, c_int]),
    "ttdg_resize_coeffs_u8": (c_int, [c_int, c_int, P, P]),
    "ttdg_resize_bilinear_u8": (c_int, [P, c_int, c_int, c_int, P, P, c_int, P, P, c_int, c_int, c_int, P, P, c_int, c_int, P]),
    "ttdg_gather_kept": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_float, P, P, P, P, P]),
    "ttdg_rois_from_padded": (c_int, [P, P, c_int, c_int, P, P]),
    "ttdg_mask_padded_candidates": (c_int, [P, P, c_int, c_int, c_int, P]),
    "ttdg_conv_tc_bf16": (c_int, [P, P, P, P, P] + [c_int] * 16 + [P, c_int, P]),
    "ttdg_weight_transpose_bf16": (c_int, [P, c_int, c_int, c_int, P, P]),
    "ttdg_stem_tc2": (c_int, [P, c_int, P, P, P, P, c_int, c_int, c_int, c_int, P, c_int, P]),
    "ttdg_maxpool3x3s2_bf16": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_relu_bn_bwd_bf16y": (c_int, [P, P, P, c_int, c_int64, P, P]),
    "ttdg_weight_transpose_split": (c_int, [P, c_int, c_int, c_int, P, P, P]),
    "ttdg_rpn_decode": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_int, c_int, P, c_float, c_float, P, P, P]),
    "ttdg_box_predict": (c_int, [P, c_int, P, c_int, P, c_int, c_int, c_float, c_float, c_float, P, P, P]),
    "ttdg_nms_scratch_bytes": (c_int64, [c_int, c_int]),
    "ttdg_nms": (c_int, [P, P, c_int, c_int, c_float, c_int, P, P, P, P]),
    "ttdg_roi_align": (c_int, [P, P, P, c_int, c_int, c_int, P, P]),
    "ttdg_pixel_shuffle2": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "ttdg_mask_paste": (c_int, [P, c_int, c_int, P, P, c_int, c_int, c_int, c_float, P, P]),
    "ttdg_mask_gt_stats": (c_int, [P, c_int, c_int, c_int, P, P]),
    "ttdg_mask_pair_counts": (c_int, [P, P, P, c_int, P, c_int, c_int, P, P]),
    "ttdg_sampler_select": (c_int, [P, P, P, c_int, P, c_int, c_int, P, P, P, P]),
    "ttdg_sampler_gather": (c_int, [P, P, P, c_int, c_int, c_int, P, P, c_int, P, P, P, P]),
    "ttdg_sampler_scatter_bwd": (c_int, [P, P, P, P, c_int, c_int, c_int, P, c_int, P, P]),
}


class TTDGError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise TTDGError(
                f"{SO_PATH} not found: the CUDA library is not built (run `python __graft_entry__.py build`). "
                "ttdg_b200 has no CPU or PyTorch fallback.")
        l = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the header and the library drifted apart
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc, what):
    if rc == 0:
        return
    if rc == -1:
        raise ValueError(f"{what}: bad argument (TTDG_E_ARG)")
    if rc == -2:
        raise ValueError(f"{what}: size above a compiled-in limit (TTDG_E_LIMIT)")
    raise TTDGError(f"{what}: CUDA error {rc}")


def limit(name):
    return lib().ttdg_limit(name.encode())
