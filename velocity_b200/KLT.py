"""Drop-in for the reference's utils/KLT.py: same function names, arguments, return types.

    cv2calcOpticalFlowPyrLK  utils/KLT.py:37-51
    KLTregional              utils/KLT.py:55-95
    KLTmain                  utils/KLT.py:99-134
    estimateAffine2D_SURF    utils/KLT.py:10-33

All pixel/point arithmetic runs in the sm_100a kernels of libvelocity_b200.so (K1 pyramid, K2
Lucas-Kanade with fused forward-backward gate, K3 affine remap, K4 descriptor matcher).  Images
may be numpy arrays (uploaded per call) or CUDA uint8 tensors (used in place, ROI slices are free).
Outputs are numpy arrays with the reference's shapes and dtypes.  The RANSAC affine fit
(cv2.estimateAffine2D, utils/KLT.py:116,127) runs on the GPU too (K10, ransac.py): OpenCV's loop with
its fixed-seed RNG, bit-identical inlier masks.  There is no CPU fallback for the kernels.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .common import addcol1
from .device import image_view, ptr, stream_ptr
from .images import boundingRect
from .lk import lk_params
from .ransac import estimateAffine2D, estimateAffine2D_masked_device

TERM_CRITERIA_COUNT, TERM_CRITERIA_EPS = 1, 2

LK_COARSE = dict(winSize=(15, 15), maxLevel=4, criteria=(TERM_CRITERIA_EPS | TERM_CRITERIA_COUNT, 10, 0.1))
LK_FINE = dict(winSize=(51, 51), maxLevel=0, criteria=(TERM_CRITERIA_EPS | TERM_CRITERIA_COUNT, 30, 0.001))


def _as_cuda_image(im):
    t, _, w, h, pitch = image_view(im)
    return t  # [h, w] uint8 CUDA tensor, last stride 1


_scratch = {}


def _workspace(nbytes, device):
    """Grow-only scratch for vel_klt_regional, one buffer per (device, CUDA stream): calls issued on the same stream are ordered,
    calls on different streams must not share scratch."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty((max(nbytes, 1 << 20),), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


def _regional_device(d0, dn, pts_dev, x0, x1, y0, y1, T32, translate, fbt, lk_param, pts_in_roi=False):
    """vel_klt_regional on two CUDA frames: crop, shift / remap, pyramids of the crops, LK with the fused forward-backward gate.
    Returns CUDA tensors (pa in ROI coordinates [N,2] f32, status [N] u8, err [N] f32); nothing is synchronised."""
    params = lk_params(fbt=fbt, **lk_param)
    n = pts_dev.shape[0]
    dev = d0.device
    L = _lib.lib()
    nbytes = int(L.vel_klt_regional_workspace(x1 - x0, y1 - y0, params.win_w, params.win_h, params.max_level, n))
    if nbytes == 0:
        raise RuntimeError("vel_klt_regional_workspace failed: %s" % L.vel_last_error().decode())
    work = _workspace(nbytes, dev)
    status = torch.empty((n,), dtype=torch.uint8, device=dev)
    pa_c = torch.empty((n, 2), dtype=torch.float32, device=dev)
    err_c = torch.empty((n,), dtype=torch.float32, device=dev)
    Tc = (C.c_float * 6)(*[float(v) for v in np.asarray(T32, np.float32).reshape(6)])
    _lib.check(L.vel_klt_regional(ptr(d0), ptr(dn), d0.shape[1], d0.shape[0], d0.stride(0), dn.stride(0), ptr(pts_dev), n, x0, x1, y0, y1, Tc,
                                  (1 if translate else 0) | (2 if pts_in_roi else 0), C.byref(params), ptr(work), work.numel(), ptr(pa_c), ptr(status), ptr(err_c),
                                  stream_ptr()), "vel_klt_regional")
    return pa_c, status, err_c


def lk_device(im1, im2, p1, fbt=None, **lk_param):
    """Device-resident core of cv2calcOpticalFlowPyrLK: CUDA tensors in, CUDA tensors out
    (next [N,2] f32, status [N] u8 -- already forward-backward gated when fbt is given --, err [N] f32)."""
    a, b = _as_cuda_image(im1), _as_cuda_image(im2)
    if a.shape != b.shape:
        raise ValueError("image sizes differ: %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    pts = p1 if isinstance(p1, torch.Tensor) and p1.is_cuda else torch.from_numpy(
        np.ascontiguousarray(np.asarray(p1, np.float32).reshape(-1, 2))).cuda()
    pts = pts.to(torch.float32).contiguous()
    if pts.shape[0] == 0:
        dev = a.device
        return (torch.empty((0, 2), dtype=torch.float32, device=dev), torch.empty((0,), dtype=torch.uint8, device=dev),
                torch.empty((0,), dtype=torch.float32, device=dev))
    # the whole frame as the "region", zero shift: one C call builds both pyramids and tracks
    return _regional_device(a, b, pts, 0, a.shape[1], 0, a.shape[0], np.float32([[1, 0], [0, 1], [0, 0]]), True, fbt, lk_param)


def cv2calcOpticalFlowPyrLK(im1, im2, p1, p2hat=None, fbt=None, **lk_param):
    """Pyramidal LK forward (and, with fbt, backward + forward-backward gate).  Returns
    (p2 float32 [N,2], v bool [N], err float32 [N,1]) like the reference."""
    out, status, err = lk_device(im1, im2, p1, fbt=fbt, **lk_param)
    packed = torch.cat([out, err.unsqueeze(1), status.to(torch.float32).unsqueeze(1)], 1).cpu().numpy()  # one D2H copy
    return np.ascontiguousarray(packed[:, 0:2]), packed[:, 3] != 0, np.ascontiguousarray(packed[:, 2:3])


def _remap_affine_device(im, T32, x0, x1, y0, y1):
    src, p, w, h, pitch = image_view(im)
    out = torch.empty((y1 - y0, x1 - x0), dtype=torch.uint8, device=src.device)
    Tc = (C.c_float * 6)(*[float(v) for v in np.asarray(T32, np.float32).reshape(6)])
    _lib.check(_lib.lib().vel_remap_affine_u8(C.c_void_p(p), w, h, pitch, Tc, x0, y0, x1 - x0, y1 - y0, ptr(out),
                                              out.stride(0), stream_ptr()), "vel_remap_affine_u8")
    return out


def _klt_regional_enqueue(d0, dn, p0, T, lk_param, fbt, translateFlag):
    """The device part of KLTregional, nothing synchronised: returns (pa_d [N,2] f32 in ROI coordinates, st_d [N] u8, xy0, (dx, dy))."""
    x0, x1, y0, y1 = boundingRect(p0, tuple(dn.shape), border=(50, 50))
    xy0 = np.float32([x0, y0])
    dx = dy = 0
    if translateFlag:
        dx, dy = int(T[2, 0]), int(T[2, 1])
        if y0 + dy < 0 or x0 + dx < 0 or y1 + dy > dn.shape[0] or x1 + dx > dn.shape[1]:
            raise ValueError("KLTregional: shifted ROI leaves the frame (cv2 asserts on mismatched pyramid sizes here)")
    # p0 - xy0 in numpy's arithmetic (float64 when the caller's points are float64), then float32 as cv2 takes them
    p0_roi = np.ascontiguousarray(np.asarray(p0 - xy0, np.float32).reshape(-1, 2))
    pts = torch.from_numpy(p0_roi).to(d0.device)
    pa_d, st_d, _ = _regional_device(d0, dn, pts, x0, x1, y0, y1, T, translateFlag, fbt, lk_param, pts_in_roi=True)
    return pa_d, st_d, xy0, (dx, dy)


def KLTregional(im0, im, p0, T, lk_param, fbt=1.0, translateFlag=False):
    """ROI tracker: warp the current frame into the previous frame's coordinates (integer shift or
    affine remap), LK forward+backward on the ROI, map the result back through T.  The device work is one call
    (vel_klt_regional); the ROI rectangle (utils/KLT.py:60) and the map-back (:88-93) are the reference's numpy."""
    T = np.asarray(T).astype(np.float32)
    p0 = np.asarray(p0)
    d0, dn = _as_cuda_image(im0), _as_cuda_image(im)
    pa_d, st_d, xy0, (dx, dy) = _klt_regional_enqueue(d0, dn, p0, T, lk_param, fbt, translateFlag)
    packed = torch.cat([pa_d, st_d.to(torch.float32).unsqueeze(1)], 1).cpu().numpy()      # one D2H copy
    pa, v = np.ascontiguousarray(packed[:, 0:2]), packed[:, 2] != 0
    if translateFlag:
        p = pa + (xy0 + [dx, dy]).astype(np.float32)
    else:
        p = addcol1(pa + xy0) @ T
    return p, v


def _decimate4_device(im):
    src, p, w, h, pitch = image_view(im)
    dw, dh = int(np.rint(w * 0.25)), int(np.rint(h * 0.25))
    out = torch.empty((dh, dw), dtype=torch.uint8, device=src.device)
    _lib.check(_lib.lib().vel_decimate4_u8(C.c_void_p(p), w, h, pitch, ptr(out), dw, dh, out.stride(0), stream_ptr()),
               "vel_decimate4_u8")
    return out


def KLTmain(im, im0, im0_small, p0):
    """Three-stage tracker: quarter-scale LK -> RANSAC -> translated-ROI LK (FB 1.0) -> RANSAC ->
    affine-ROI fine LK (FB 0.3).  Returns (p[v] float32, v bool [N], im_small) like the reference;
    im_small is returned in the form it was produced (CUDA tensor), pass it back as im0_small.
    Each LK stage hands its points and status mask to the robust fit ON THE DEVICE (vel_estimate_affine2d_ransac_masked: the
    `estimateAffine2D(p0[v], p[v]); v[v] = inliers` of utils/KLT.py:116-117 in one launch), so a frame costs three read-backs -- what
    the host needs to size the next stage (the mean translation, the 2x3 affine) and the result -- instead of five, and one upload of
    p0 for both fits."""
    p0 = np.asarray(p0)
    d_im, d_im0 = _as_cuda_image(im), _as_cuda_image(im0)
    scale = 1 / 4
    im_small = _decimate4_device(d_im)
    if im0_small is None:
        im0_small = _decimate4_device(d_im0)
    p0_dev = torch.from_numpy(np.ascontiguousarray(np.asarray(p0, np.float32).reshape(-1, 2))).to(d_im.device)
    n = p0_dev.shape[0]
    if n == 0:                                              # the reference fails the same way (estimateAffine2D returns None, utils/KLT.py:117)
        raise AttributeError("KLTmain: no points to track")
    # stage 1 (utils/KLT.py:113-121): p0 * scale is exact in float32 (a power of two), so the device copy serves
    out, status, _ = lk_device(im0_small, im_small, p0_dev * scale, **LK_COARSE)
    mask1, p1_dev, tail1 = estimateAffine2D_masked_device(p0_dev, out, status, to_scale=1 / scale)
    host = torch.cat([p1_dev.reshape(-1), mask1.to(torch.float32), tail1.view(torch.float32)]).cpu().numpy()       # read-back 1
    p, v = host[:2 * n].reshape(n, 2), host[2 * n:3 * n] != 0
    if not host[3 * n + 12:3 * n + 16].view(np.int32)[0]:
        raise AttributeError("KLTmain: estimateAffine2D found no model after the coarse stage (the reference fails on None here, utils/KLT.py:117)")
    translation = p[v] - p0[v]
    T = np.eye(3, 2)
    T[2] = translation.mean(0)

    # stage 2 (:122-130): translated ROI, forward-backward 1 px; the fit only has to deliver T23
    pa_d, st_d, xy0, (dx, dy) = _klt_regional_enqueue(d_im0, d_im, p0, T.astype(np.float32), LK_COARSE, 1, True)
    off = (xy0 + [dx, dy]).astype(np.float32)
    _, _, tail2 = estimateAffine2D_masked_device(p0_dev, pa_d, st_d, to_off=(off[0], off[1]))
    t2 = tail2.cpu().numpy()                                                                                      # read-back 2
    found, _, _, kept = (int(x) for x in t2[6:8].view(np.int32))
    if kept > 10:
        if not found:
            raise AttributeError("KLTmain: estimateAffine2D found no model after the translated stage (the reference fails on None.T here, utils/KLT.py:132)")
        T23 = t2[:6].reshape(2, 3).copy()
    else:
        print("KLT coarse-affine failure, running SURF matches full scale.")
        T23, inliers = estimateAffine2D_SURF(d_im0, d_im, p0, scale=1)

    p, v = KLTregional(d_im0, d_im, p0, T23.T, LK_FINE, fbt=0.3)                                                    # read-back 3
    return p[v], v, im_small


def knnMatch2(des1, des2):
    """cv2.BFMatcher(norm).knnMatch(des1, des2, k=2) on the GPU: uint8 [.,32] -> Hamming,
    float32 [.,D] -> L2.  Returns (idx int32 [nq,2], dist [nq,2])."""
    from .match import knn2_hamming256, knn2_l2

    des1, des2 = np.asarray(des1), np.asarray(des2)
    if des1.dtype == np.uint8:
        return knn2_hamming256(des1, des2)
    return knn2_l2(des1.astype(np.float32), des2.astype(np.float32))


def estimateAffine2D_SURF(im1, im2, p1, scale=1.0):
    """Descriptor-matching fallback (utils/KLT.py:10-33).  The reference asks for SURF, which is a
    non-free module absent from opencv-python; this build extracts ORB descriptors on the host
    (BASELINE.json config 4 names ORB) and runs the all-pairs 2-NN match on the GPU (K4)."""
    import cv2

    def host(im):
        return im.cpu().numpy() if isinstance(im, torch.Tensor) else np.asarray(im)

    im1 = cv2.resize(host(im1), (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_NEAREST)
    im2 = cv2.resize(host(im2), (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_NEAREST)
    orb = cv2.ORB_create(nfeatures=8192)
    kp2, des2 = orb.detectAndCompute(im2, mask=None)
    a, ngood, good = 0, 0, []
    while ngood < 10:
        border = int(a * scale)
        x0, x1, y0, y1 = boundingRect(np.asarray(p1) * scale, im1.shape, border=(border, border))
        kp1, des1 = orb.detectAndCompute(np.ascontiguousarray(im1[y0:y1, x0:x1]), mask=None)
        if des1 is not None and len(des1) >= 1 and des2 is not None and len(des2) >= 2:
            idx, dist = knnMatch2(des1, des2)
            good = [(q, int(idx[q, 0])) for q in range(len(des1)) if dist[q, 0] < 0.6 * dist[q, 1]]
        ngood = len(good)
        a += 10
        if x0 <= 1 and y0 <= 1 and x1 >= im1.shape[1] and y1 >= im1.shape[0] and ngood < 10:
            raise RuntimeError("estimateAffine2D_SURF: fewer than 10 ratio-test matches on the full frame")
    m1 = np.float32([kp1[q].pt for q, _ in good]) + np.float32([x0, y0])
    m2 = np.float32([kp2[t].pt for _, t in good])
    return estimateAffine2D(m1 / scale, m2 / scale)
