"""ctypes front-end of oracle/velocity_oracle.c (CPU ORACLE -- test infrastructure only).

Each function takes/returns numpy arrays with the shapes and dtypes of the cv2 call it restates.
See velocity_oracle.c for the reference call sites (utils/KLT.py:45,48,73,111,113,16-25).
"""
import ctypes as C
import os

import numpy as np

from .build_oracle import OUT, build

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = OUT if os.path.exists(OUT) else build()
        try:
            path = build()  # rebuild if the source is newer
        except Exception:
            pass
        L = C.CDLL(path)
        u8p, f32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32)
        L.orc_pyrdown_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.orc_pyrdown_u8.restype = None
        L.orc_decimate4_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int, C.c_int, C.c_int]
        L.orc_decimate4_u8.restype = None
        L.orc_scharr_s16.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int16)]
        L.orc_scharr_s16.restype = None
        L.orc_lk_effective_max_level.argtypes = [C.c_int] * 5
        L.orc_lk_effective_max_level.restype = C.c_int
        L.orc_calc_optical_flow_pyr_lk.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int,
                                                   C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, f32p, u8p, f32p,
                                                   C.c_int]
        L.orc_calc_optical_flow_pyr_lk.restype = C.c_int
        L.orc_remap_affine_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_int, u8p,
                                          C.c_int]
        L.orc_remap_affine_u8.restype = None
        L.orc_bgr2gray_u8.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p, C.c_int]
        L.orc_bgr2gray_u8.restype = None
        L.orc_knn2_hamming.argtypes = [u8p, C.c_int, u8p, C.c_int, C.c_int, i32p, i32p]
        L.orc_knn2_hamming.restype = None
        L.orc_knn2_l2.argtypes = [f32p, C.c_int, f32p, C.c_int, C.c_int, i32p, f32p]
        L.orc_knn2_l2.restype = None
        _lib = L
    return _lib


def _u8(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _i32(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def _rowmajor_u8(im):
    """Accept any 2-D uint8 array whose last axis is contiguous (numpy ROI slices included)."""
    im = np.asarray(im)
    assert im.dtype == np.uint8 and im.ndim == 2
    if im.strides[1] != 1 or im.strides[0] < im.shape[1]:
        im = np.ascontiguousarray(im)
    return im, int(im.strides[0])


def pyrDown(im):
    im, pitch = _rowmajor_u8(im)
    h, w = im.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyrdown_u8(_u8(im), w, h, pitch, _u8(out), out.shape[1])
    return out


def decimate4(im):
    im, pitch = _rowmajor_u8(im)
    h, w = im.shape
    dw, dh = int(np.rint(w * 0.25)), int(np.rint(h * 0.25))
    out = np.empty((dh, dw), np.uint8)
    lib().orc_decimate4_u8(_u8(im), w, h, pitch, _u8(out), dw, dh, dw)
    return out


def scharr(im):
    im, pitch = _rowmajor_u8(im)
    h, w = im.shape
    out = np.empty((h, w, 2), np.int16)
    lib().orc_scharr_s16(_u8(im), w, h, pitch, out.ctypes.data_as(C.POINTER(C.c_int16)))
    return out


def effective_max_level(w, h, win, max_level):
    return lib().orc_lk_effective_max_level(w, h, win[0], win[1], max_level)


def calcOpticalFlowPyrLK(im1, im2, p1, winSize=(21, 21), maxLevel=3, criteria=(3, 30, 0.01), minEigThreshold=1e-4,
                         nthreads=0):
    """Same return convention as cv2: (nextPts [N,2] f32, status [N,1] u8, err [N,1] f32)."""
    a, pa = _rowmajor_u8(im1)
    b, pb = _rowmajor_u8(im2)
    assert a.shape == b.shape
    h, w = a.shape
    pts = np.ascontiguousarray(np.asarray(p1, np.float32).reshape(-1, 2))
    n = pts.shape[0]
    out = np.zeros((n, 2), np.float32)
    st = np.zeros((n, 1), np.uint8)
    err = np.zeros((n, 1), np.float32)
    _, max_count, eps = criteria
    lib().orc_calc_optical_flow_pyr_lk(_u8(a), _u8(b), w, h, pa, pb, _f32(pts), n, winSize[0], winSize[1], maxLevel,
                                       int(max_count), float(eps), float(minEigThreshold), _f32(out), _u8(st), _f32(err),
                                       nthreads)
    return out, st, err


def remap_affine(im, T, x0, x1, y0, y1):
    """utils/KLT.py:70-73: warp `im` through the 3x2 row-vector affine T (float32) onto the grid
    x in [x0,x1), y in [y0,y1)."""
    im, pitch = _rowmajor_u8(im)
    h, w = im.shape
    T = np.ascontiguousarray(np.asarray(T, np.float32).reshape(3, 2))
    out = np.empty((y1 - y0, x1 - x0), np.uint8)
    lib().orc_remap_affine_u8(_u8(im), w, h, pitch, _f32(T), x0, y0, x1 - x0, y1 - y0, _u8(out), out.shape[1])
    return out


def knn2_hamming(q, t):
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    idx = np.empty((q.shape[0], 2), np.int32)
    dist = np.empty((q.shape[0], 2), np.int32)
    lib().orc_knn2_hamming(_u8(q), q.shape[0], _u8(t), t.shape[0], q.shape[1], _i32(idx), _i32(dist))
    return idx, dist


def knn2_l2(q, t):
    q = np.ascontiguousarray(q, np.float32)
    t = np.ascontiguousarray(t, np.float32)
    idx = np.empty((q.shape[0], 2), np.int32)
    dist = np.empty((q.shape[0], 2), np.float32)
    lib().orc_knn2_l2(_f32(q), q.shape[0], _f32(t), t.shape[0], q.shape[1], _i32(idx), _f32(dist))
    return idx, dist


def bgr2gray(bgr):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    lib().orc_bgr2gray_u8(_u8(bgr), w, h, 3 * w, _u8(out), w)
    return out


def cornerSubPix(im, corners, winSize=(5, 5), zeroZone=(-1, -1), criteria=(3, 100, 0.001), nthreads=0):
    """cv2.cornerSubPix as vidExample.py:113-115 calls it (zeroZone (-1,-1)); returns refined float32 [n, 2]."""
    assert tuple(zeroZone) == (-1, -1)
    im, pitch = _rowmajor_u8(im)
    h, w = im.shape
    pts = np.ascontiguousarray(np.asarray(corners, np.float32).reshape(-1, 2)).copy()
    ctype, count, eps = criteria
    L = lib()
    L.orc_corner_subpix_u8.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_double]
    L.orc_corner_subpix_u8.restype = None
    L.orc_corner_subpix_u8(_u8(im), w, h, pitch, _f32(pts), pts.shape[0], int(winSize[0]), int(winSize[1]),
                           int(count) if (ctype & 1) else 100, float(eps) if (ctype & 2) else 0.0)
    return pts
