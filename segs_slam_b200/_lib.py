"""ctypes binding of the C-ABI library (include/segs_raster.h).

The library is built in-tree (segs_slam_b200/libsegs_raster.so) by
``make -C segs_slam_b200/csrc`` / ``__graft_entry__.build()``.  There is no fallback:
if the shared object is missing, loading raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsegs_raster.so")

ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)

_f32p = C.c_void_p  # device pointers travel as plain addresses
_i32p = C.c_void_p

_PARAM_NAMES = ("opacity_w1", "opacity_b1", "opacity_w2", "opacity_b2", "cov_w1", "cov_b1", "cov_w2", "cov_b2",
                "color_w1", "color_b1", "color_w2", "color_b2", "app_w", "app_b", "bank_w1", "bank_b1", "bank_w2",
                "bank_b2")


class DecodeParams(C.Structure):
    """segs_decode_params (include/segs_raster.h)."""
    _fields_ = [(n, C.c_void_p) for n in _PARAM_NAMES] + [
        ("appearance_dim", C.c_int), ("use_feat_bank", C.c_int), ("add_opacity_dist", C.c_int),
        ("add_cov_dist", C.c_int), ("add_color_dist", C.c_int)]


class DecodeGrads(C.Structure):
    """segs_decode_grads (include/segs_raster.h)."""
    _fields_ = [(n, C.c_void_p) for n in _PARAM_NAMES]


class MapperViewArgs(C.Structure):
    """segs_mapper_view_args (include/segs_raster.h)."""
    _fields_ = [("A", C.c_int), ("anchor", C.c_void_p), ("anchor_feat", C.c_void_p), ("offset", C.c_void_p),
                ("scaling", C.c_void_p), ("scaling_is_log", C.c_int), ("filter_scales", C.c_void_p),
                ("filter_rotations", C.c_void_p), ("params", C.POINTER(DecodeParams)),
                ("width", C.c_int), ("height", C.c_int), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("pose", C.POINTER(C.c_float)), ("background", C.c_void_p), ("gt_image", C.c_void_p),
                ("row_mask", C.c_void_p), ("lambda_dssim", C.c_float), ("scaling_reg_weight", C.c_float),
                ("grad_anchor", C.c_void_p), ("grad_anchor_feat", C.c_void_p), ("grad_offset", C.c_void_p),
                ("grad_scaling", C.c_void_p), ("grad_params", C.POINTER(DecodeGrads)), ("loss_accum", C.c_void_p),
                ("image_out", C.c_void_p), ("loss_terms_out", C.c_void_p), ("dL_dmean2D_out", C.c_void_p),
                ("radii_out", C.c_void_p), ("stat_opacity_accum", C.c_void_p), ("stat_anchor_demon", C.c_void_p),
                ("stat_offset_gradient_accum", C.c_void_p), ("stat_offset_denom", C.c_void_p),
                ("lambda_frequency_high", C.c_float), ("use_multi_resolution", C.c_int), ("freq_scale_num", C.c_int),
                ("gt_freq_mag", C.c_void_p)]


class RasterViewArgs(C.Structure):
    """segs_raster_view_args (include/segs_raster.h)."""
    _fields_ = [("P", C.c_int), ("means3D", C.c_void_p), ("colors_precomp", C.c_void_p), ("opacities", C.c_void_p),
                ("scales", C.c_void_p), ("rotations", C.c_void_p), ("background", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("tan_fovx", C.c_float), ("tan_fovy", C.c_float),
                ("viewmatrix", C.c_void_p), ("projmatrix", C.c_void_p), ("campos", C.c_void_p),
                ("dL_dout", C.c_void_p), ("image_out", C.c_void_p), ("radii_out", C.c_void_p),
                ("grad_means3D", C.c_void_p), ("grad_means2D", C.c_void_p), ("grad_colors", C.c_void_p),
                ("grad_opacity", C.c_void_p), ("grad_scales", C.c_void_p), ("grad_rotations", C.c_void_p)]


class MapperViewResult(C.Structure):
    """segs_mapper_view_result (include/segs_raster.h)."""
    _fields_ = [("n_visible", C.c_int), ("n_gaussians", C.c_int), ("num_rendered", C.c_int)]


class AdamTensor(C.Structure):
    """segs_adam_tensor (include/segs_raster.h)."""
    _fields_ = [("param", C.c_void_p), ("offset", C.c_ulonglong), ("count", C.c_ulonglong), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
                ("step", C.c_longlong)]


_PROTOTYPES = {
    "segs_version": (C.c_int, []),
    "segs_last_error": (C.c_char_p, []),
    "segs_raster_forward": (
        C.c_int,
        [ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p,
         C.c_int, C.c_int, C.c_int,
         _f32p, C.c_int, C.c_int,
         _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, _f32p,
         _f32p, _f32p, _f32p,
         C.c_float, C.c_float, C.c_int,
         _f32p, _i32p, C.POINTER(C.c_int), C.c_void_p],
    ),
    "segs_raster_backward": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, C.c_int,
         _f32p, C.c_int, C.c_int,
         _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p,
         C.c_float, C.c_float, _i32p,
         C.c_void_p, C.c_void_p, C.c_void_p,
         _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p],
    ),
    "segs_visible_filter": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, _f32p,
         C.c_float, C.c_float, C.c_int, _i32p, C.c_void_p],
    ),
    "segs_mark_visible": (C.c_int, [C.c_int, _f32p, _f32p, _f32p, C.c_void_p, C.c_void_p]),
    "segs_project": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
         _f32p, _f32p, _f32p, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, _f32p, _f32p,
         C.c_float, C.c_float, C.c_int, _f32p, _f32p, _i32p, C.c_void_p],
    ),
    "segs_knn_mean_dist2": (C.c_int, [C.c_int, _f32p, _f32p, ALLOC_FN, C.c_void_p, C.c_void_p]),
    "segs_decode_state_bytes": (C.c_size_t, [C.c_int]),
    "segs_decode_set_variant": (C.c_int, [C.c_int]),
    "segs_decode_get_variant": (C.c_int, []),
    "segs_decode_forward": (
        C.c_int,
        [C.c_int, C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(DecodeParams),
         _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p],
    ),
    "segs_decode_backward": (
        C.c_int,
        [C.c_int, C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(DecodeParams),
         C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
         _f32p, _f32p, _f32p, _f32p, C.POINTER(DecodeGrads), ALLOC_FN, C.c_void_p, C.c_void_p],
    ),
    "segs_decode_backward_ex": (
        C.c_int,
        [C.c_int, C.c_void_p, _f32p, _f32p, _f32p, _f32p, _f32p, C.POINTER(C.c_float), C.POINTER(DecodeParams),
         C.c_void_p, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
         _f32p, _f32p, _f32p, _f32p, C.POINTER(DecodeGrads), ALLOC_FN, C.c_void_p, C.c_int, C.c_void_p],
    ),
    "segs_training_statis": (
        C.c_int,
        [C.c_int, C.c_void_p, C.c_int, _f32p, _i32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, C.c_void_p],
    ),
    "segs_ingest_image": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_void_p]),
    "segs_resize_bilinear": (C.c_int, [C.c_int, C.c_int, C.c_int, _f32p, C.c_int, C.c_int, _f32p, C.c_void_p]),
    "segs_freq_plan_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_void_p)]),
    "segs_freq_plan_destroy": (C.c_int, [C.c_void_p]),
    "segs_freq_mag_floats": (C.c_size_t, [C.c_void_p]),
    "segs_freq_target": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_void_p]),
    "segs_freq_loss": (C.c_int, [C.c_void_p, _f32p, _f32p, _f32p, C.c_float, _f32p, _f32p, _f32p, C.c_void_p]),
    "segs_anchor_growing_level": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p,
         C.c_float, C.c_float, C.c_float, C.c_float, ALLOC_FN, C.c_void_p, ALLOC_FN, C.c_void_p,
         C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p],
    ),
    "segs_prune_scratch_words": (C.c_size_t, [C.c_int]),
    "segs_prune_plan": (
        C.c_int,
        [C.c_int, C.c_int, _f32p, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
         C.POINTER(C.c_int), C.c_void_p],
    ),
    "segs_compact_rows": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_void_p, _f32p, _f32p, C.c_int, C.c_float, C.c_void_p]),
    "segs_workspace_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "segs_workspace_destroy": (C.c_int, [C.c_void_p]),
    "segs_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "segs_mapper_view": (C.c_int, [C.c_void_p, C.POINTER(MapperViewArgs), C.POINTER(MapperViewResult), C.c_void_p]),
    "segs_mapper_views": (
        C.c_int,
        [C.c_int, C.POINTER(MapperViewArgs), C.POINTER(MapperViewResult), C.c_int, C.POINTER(C.c_void_p),
         C.POINTER(C.c_void_p), C.c_void_p],
    ),
    "segs_raster_views": (
        C.c_int,
        [C.c_int, C.POINTER(RasterViewArgs), C.POINTER(MapperViewResult), C.c_int, C.POINTER(C.c_void_p),
         C.POINTER(C.c_void_p), C.c_void_p],
    ),
    "segs_loss_state_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "segs_loss_l1_ssim_forward": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_float, C.c_float, C.c_float, _f32p, C.c_void_p, C.c_void_p],
    ),
    "segs_loss_l1_ssim_backward": (
        C.c_int,
        [C.c_int, C.c_int, C.c_int, _f32p, _f32p, _f32p, C.c_float, C.c_float, _f32p, C.c_void_p, _f32p, C.c_void_p],
    ),
    "segs_scaling_reg": (C.c_int, [C.c_int, _f32p, C.c_float, _f32p, _f32p, _f32p, C.c_void_p]),
    "segs_adam_step": (
        C.c_int,
        [C.c_int, C.POINTER(AdamTensor), _f32p, _f32p, _f32p, C.c_float, C.c_int, C.c_void_p],
    ),
    "segs_accumulate": (
        C.c_int,
        [C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_ulonglong), C.c_int, C.c_void_p],
    ),
    "segs_launch_count": (C.c_ulonglong, []),
    "segs_profile_enable": (C.c_int, [C.c_int]),
    "segs_set_blocking_sync": (C.c_int, [C.c_int]),
    "segs_profile_read": (C.c_int, [C.POINTER(C.c_float)]),
    "segs_debug_blend_stats": (C.c_int, [C.POINTER(C.c_ulonglong), C.c_int]),
    "segs_buffer_section": (
        C.c_int,
        [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
         C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)],
    ),
}

_lib = None


def load() -> C.CDLL:
    """Load libsegs_raster.so (once).  Raises if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `make -C segs_slam_b200/csrc` "
                "(or __graft_entry__.build()). There is no CPU fallback."
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols() -> list[str]:
    return sorted(_PROTOTYPES)


class SegsError(RuntimeError):
    pass


def check(status: int) -> None:
    if status != 0:
        msg = load().segs_last_error()
        raise SegsError(f"segs status {status}: {msg.decode() if msg else ''}")
