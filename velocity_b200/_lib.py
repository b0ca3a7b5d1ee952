"""ctypes binding of libvelocity_b200.so (include/velocity_b200.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvelocity_b200.so")

VEL_MAX_LEVELS = 8


class PyrLayout(C.Structure):
    _fields_ = [
        ("max_level", C.c_int32),
        ("width", C.c_int32 * VEL_MAX_LEVELS),
        ("height", C.c_int32 * VEL_MAX_LEVELS),
        ("pitch", C.c_int32 * VEL_MAX_LEVELS),
        ("offset", C.c_int64 * VEL_MAX_LEVELS),
        ("bytes", C.c_int64),
    ]


class LkParams(C.Structure):
    _fields_ = [
        ("win_w", C.c_int32),
        ("win_h", C.c_int32),
        ("max_level", C.c_int32),
        ("max_count", C.c_int32),
        ("eps", C.c_double),
        ("min_eig_threshold", C.c_float),
        ("fb_threshold", C.c_float),
    ]


_P = C.c_void_p
_I32, _I64 = C.c_int32, C.c_int64

# name -> (restype, argtypes); mirrors include/velocity_b200.h one to one
SIGNATURES = {
    "vel_version": (C.c_int, []),
    "vel_last_error": (C.c_char_p, []),
    "vel_pyr_layout_make": (C.c_int, [_I32, _I32, _I32, _I32, _I32, C.POINTER(PyrLayout)]),
    "vel_pyramid_u8": (C.c_int, [_P, _I64, _I32, _I32, C.POINTER(PyrLayout), _P, _I64, _P]),
    "vel_decimate4_u8": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _I32, _I32, _P]),
    "vel_bgr2gray_u8": (C.c_int, [_P, _I64, _I32, _I32, _I32, _I32, _P, _I64, _I32, _P]),
    "vel_good_features_workspace": (C.c_size_t, [_I32, _I32, _I32]),
    "vel_good_features_harris_u8": (C.c_int, [_P, _I32, _I32, _I32, _I32, C.c_double, _I32, C.c_double, _P, C.c_size_t, _P, _P, _P, _P]),
    "vel_corner_subpix_u8": (C.c_int, [_P, _I32, _I32, _I32, _P, _I32, _I32, _I32, _I32, C.c_double, _P]),
    "vel_estimate_affine2d_ransac": (C.c_int, [_P, _P, _I32, C.c_double, C.c_double, _I32, _I32, _P, _P, _P, _P]),
    "vel_estimate_affine2d_ransac_masked_workspace": (C.c_size_t, [_I32]),
    "vel_estimate_affine2d_ransac_masked": (C.c_int, [_P, _P, _P, _I32, C.c_float, C.c_float, C.c_float, C.c_double, C.c_double, _I32, _I32, _P,
                                                     C.c_size_t, _P, _P, _P, _P, _P]),
    "vel_lk_track": (C.c_int, [_P, _I64, _I32, _P, _I64, _P, _I64, _I32, _P, _I64, C.POINTER(PyrLayout), _I32, _P, _I64, _I32,
                               C.POINTER(LkParams), _P, _P, _P, _P, _P]),
    "vel_klt_regional_workspace": (C.c_size_t, [_I32, _I32, _I32, _I32, _I32, _I32]),
    "vel_klt_regional": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P, _I32, _I32, _I32, _I32, _I32, C.POINTER(C.c_float), _I32,
                                   C.POINTER(LkParams), _P, C.c_size_t, _P, _P, _P, _P]),
    "vel_remap_affine_u8": (C.c_int, [_P, _I32, _I32, _I32, C.POINTER(C.c_float), _I32, _I32, _I32, _I32, _P, _I32, _P]),
    "vel_nls_t": (C.c_int, [_P, _P, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "vel_nls_rt": (C.c_int, [_P, _P, _P, _P, _P, _I32, _P, _P, _P, _P]),
    "vel_triangulate_2v": (C.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "vel_triangulate_nv": (C.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "vel_msv1_t": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _P, _I32, _P, _P, _P, _P]),
    "vel_ba_accumulate": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "vel_ba_solve_workspace": (C.c_size_t, [_I32, _I32]),
    "vel_ba_solve": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P, _P, _P, C.c_size_t, _P]),
    "vel_ba_iterate_workspace": (C.c_size_t, [_I32, _I32]),
    "vel_ba_iterate": (C.c_int, [_P, _P, _I32, _I32, _P, _I32, C.c_double, _P, _P, _P, C.c_size_t, _P]),
    "vel_syrk_lower_sub_workspace": (C.c_size_t, [_I32, _I32]),
    "vel_syrk_lower_sub": (C.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P, C.c_size_t, _P]),
    "vel_syrk_tile_rows": (C.c_int, [_I32, C.POINTER(_I32), C.POINTER(_I32)]),
    "vel_syrk_lower_sub_rows": (C.c_int, [_P, _I64, _I32, _I32, _P, _I64, _P, C.c_size_t, _I32, _I32, _P]),
    "vel_ba_solve_layout": (C.c_int, [_I32, _I32, C.POINTER(_I64), C.POINTER(_I64)]),
    "vel_ba_reduce": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _I32, _I32, _P, C.c_size_t, _P]),
    "vel_ba_factor": (C.c_int, [_I32, _I32, _P, C.c_size_t, _P]),
    "vel_ba_update": (C.c_int, [_P, _I32, _I32, _P, _P, _P, C.c_size_t, _P]),
    "vel_spd_solve": (C.c_int, [_P, _I64, _I32, _P, _P, _P]),
    "vel_fp64_mma_peak_tflops": (C.c_double, []),
    "vel_ba2_accumulate": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "vel_ba2_solve": (C.c_int, [_P, _P, _P, _P, _I32, _I32, _P, _P, _P, C.c_size_t, _P]),
    "vel_match_knn2_hamming256": (C.c_int, [_P, _I32, _P, _I32, _P, _P, _P]),
    "vel_match_knn2_l2": (C.c_int, [_P, _I32, _P, _I32, _I32, _P, _P, _P]),
    "vel_klt_sequence": (C.c_int, [_P, _I64, _I32, _P, _I64, C.POINTER(PyrLayout), _I32, _I32, C.POINTER(LkParams), _P, _P, _P, _P, _P]),
    "vel_seq_pose_t": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I32, C.POINTER(C.c_double), _P, _P, _P, _P, _P]),
    "vel_seq_stats": (C.c_int, [_P, _P, _I32, _I32, _P, _P]),
    "vel_seq_select": (C.c_int, [_P, _P, _I32, _P, _P, _P]),
    "vel_seq_rays": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "vel_seq_pack_ba": (C.c_int, [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P, _P]),
    "vel_seq_ba_cameras": (C.c_int, [_P, _I32, _I32, _P, _P, _P, _P, _P]),
    "vel_seq_export_P": (C.c_int, [_P, _P, _P, _I32, _I32, _P, _P]),
}

# entry points declared in the header whose kernels have not landed yet (shrinks to empty)
PENDING = set()

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "velocity_b200: %s is missing -- build it with `python -m velocity_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if name in PENDING:
                continue
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().vel_last_error().decode("utf-8", "replace")
        raise RuntimeError("velocity_b200 %s failed (code %d): %s" % (what, rc, msg))


def pyr_layout(width, height, win, max_level):
    lay = PyrLayout()
    check(lib().vel_pyr_layout_make(width, height, win[0], win[1], max_level, C.byref(lay)), "vel_pyr_layout_make")
    return lay
