"""ctypes binding of libmrfa_b200.so (C ABI declared in include/mrfa_b200.h).

There is deliberately no fallback: if the library is missing this module raises, and every
wrapper raises on non-CUDA tensors.  Build with ``python mrfa_b200/build.py``.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmrfa_b200.so")

ABI_VERSION = 8
COORD_NORM_ACF, COORD_NORM_ACT, COORD_PIXEL = 0, 1, 2
PAD_ZEROS, PAD_REFLECTION = 0, 1
TPS_L1, TPS_L2SQ = 0, 1
MAP_ROWMAJOR, MAP_TILED = 0, 1


class GridStrides(ctypes.Structure):
    _fields_ = [("sn", c_int64), ("sy", c_int64), ("sx", c_int64), ("sc", c_int64)]


# name -> (restype, argtypes); mirrors include/mrfa_b200.h line by line
SIGNATURES = {
    "mrfa_abi_version": (c_int, []),
    "mrfa_error_string": (c_char_p, [c_int]),
    "mrfa_grid_sample_fwd": (c_int, [c_void_p, c_void_p, GridStrides, c_void_p] + [c_int] * 11 + [c_void_p]),
    "mrfa_grid_sample_bwd": (c_int, [c_void_p, c_void_p, c_void_p, GridStrides, c_void_p, c_void_p]
                             + [c_int] * 11 + [c_void_p]),
    "mrfa_dual_warp_fwd": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_int64, c_void_p]),
    "mrfa_coords_grid": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p]),
    "mrfa_make_coordinate_grid": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "mrfa_kp2gaussian": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "mrfa_dense_motion_prior": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_float, c_void_p]),
    "mrfa_dense_motion_prior_bwd_workspace": (c_int64, [c_int, c_int]),
    "mrfa_dense_motion_prior_bwd": (c_int, [c_void_p] * 15 + [c_int] * 5 + [c_float, c_void_p]),
    "mrfa_kp2gaussian_bwd": (c_int, [c_void_p] * 3 + [c_int] * 3 + [c_float, c_void_p]),
    "mrfa_tps_motion_prior_bwd_workspace": (c_int64, [c_int, c_int]),
    "mrfa_tps_motion_prior_bwd": (c_int, [c_void_p] * 13 + [c_int] * 5 + [c_float, c_void_p]),
    "mrfa_tps_solve": (c_int, [c_void_p] * 4 + [c_int, c_void_p]),
    "mrfa_tps_motion_prior": (c_int, [c_void_p] * 8 + [c_int] * 5 + [c_float, c_void_p]),
    "mrfa_prior_to_flow": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "mrfa_corr_rows_total": (c_int64, [c_int, c_int]),
    "mrfa_corr_row_offset": (c_int64, [c_int, c_int, c_int]),
    "mrfa_corr_map_layout": (c_int, [c_int, c_int]),
    "mrfa_corr_map_offset": (c_int64, [c_int] * 5),
    "mrfa_corr_pack": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "mrfa_corr_pack_bias": (c_int, [c_void_p] * 6 + [c_int] * 4 + [c_void_p]),
    "mrfa_corr_volume": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_float, c_int, c_void_p]),
    "mrfa_corr_bwd_rows_pad": (c_int64, [c_int, c_int]),
    "mrfa_corr_bwd_pack": (c_int, [c_void_p] * 4 + [c_int] * 3 + [c_float, c_void_p]),
    "mrfa_transpose_bf16": (c_int, [c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "mrfa_corr_bwd_gemm": (c_int, [c_void_p] * 3 + [c_int] * 4 + [c_int64, c_int64, c_int, c_void_p]),
    "mrfa_corr_bwd_unpack": (c_int, [c_void_p] * 4 + [c_int] * 4 + [c_void_p]),
    "mrfa_avg_pool2x2": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "mrfa_corr_lookup_fwd": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p] + [c_int] * 4
                             + [c_int64, c_int64, c_int, c_int, c_int, c_void_p]),
    "mrfa_channel_affine": (c_int, [c_void_p] * 5 + [c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "mrfa_occlusion_blend_subpixel": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_int64, c_void_p]),
    "mrfa_avg_pool2x2_nhwc": (c_int, [c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "mrfa_antialias_down": (c_int, [c_void_p] * 3 + [c_int] * 7 + [c_void_p]),
    "mrfa_resize_bilinear": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 8 + [c_void_p]),
    "mrfa_resize_bilinear_strip": (c_int, [c_void_p, c_void_p, c_int64] + [c_int] * 4 + [c_int64, c_int64, c_int, c_void_p]),
    "mrfa_random_warp_grid": (c_int, [c_void_p] * 4 + [c_int] * 5 + [c_void_p]),
    "mrfa_conv7x7_small_kpad": (c_int, [c_int]),
    "mrfa_conv7x7_small": (c_int, [c_void_p, GridStrides] + [c_void_p] * 3 + [c_int] * 7 + [c_void_p]),
    "mrfa_subpixel_shuffle_cat": (c_int, [c_void_p, c_void_p, GridStrides, c_void_p] + [c_int] * 5 + [c_void_p]),
    "mrfa_avg_pool2x2_nhwc_bwd": (c_int, [c_void_p, c_void_p] + [c_int] * 4 + [c_void_p]),
    "mrfa_cat2_nhwc": (c_int, [c_void_p] * 3 + [c_int64, c_int, c_int, c_void_p]),
    "mrfa_flow_update": (c_int, [c_void_p] * 3 + [GridStrides] + [c_void_p] * 3 + [c_int] * 4 + [c_void_p]),
    "mrfa_flow_carry": (c_int, [c_void_p, GridStrides] + [c_void_p] * 8 + [c_int] * 3 + [c_float, c_int, c_void_p]),
    "mrfa_final_blend_s2d": (c_int, [c_void_p] * 5 + [c_int] * 5 + [c_void_p]),
    "mrfa_occlusion_blend": (c_int, [c_void_p] * 4 + [c_int64, c_int, c_int, c_int, c_void_p]),
    "mrfa_corr_lookup_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int] + [c_void_p] * 4 + [c_int] * 4
                             + [c_int64, c_int64, c_int, c_int, c_void_p]),
}


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA kernels are the product and there is no fallback. "
            "Build them with `python mrfa_b200/build.py` (needs nvcc, no GPU required).")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == header / library drift
        fn.restype, fn.argtypes = res, args
    if lib.mrfa_abi_version() != ABI_VERSION:
        raise ImportError(f"libmrfa_b200.so ABI {lib.mrfa_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    return lib


lib = _load()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise RuntimeError(f"{what or 'mrfa_b200 kernel'} failed with code {rc}: {lib.mrfa_error_string(rc).decode()}")
