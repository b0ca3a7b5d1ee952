"""CPU ORACLE (test infrastructure) for the tracker layer of the reference, utils/KLT.py.

Restates the reference's three tracker entry points on top of oracle/velocity_oracle.c instead of
cv2 (opencv-python 4.13.0 is the un-vendored third-party dependency whose arithmetic that C file
restates).  Pinned by tests/golden/lk_*.npz, regional_*.npz, kltmain_pair.npz.

    lk_forward_backward   <- utils/KLT.py:37-51   cv2calcOpticalFlowPyrLK
    klt_regional          <- utils/KLT.py:55-95   KLTregional
    klt_main              <- utils/KLT.py:99-134  KLTmain

cv2.estimateAffine2D (RANSAC, utils/KLT.py:116,127) is restated in oracle/ransac_oracle.py (inlier masks
bit-exact against cv2, T to ~1e-12) and used by klt_main below.
"""
import numpy as np

from . import cv_oracle as cvo


def bounding_rect(pts, imshape, border=(0, 0)):
    """utils/images.py:9-19 over cv2.boundingRect(float32 points): x0=floor(min), w=floor(max)-x0+1."""
    pts = np.asarray(pts)
    fx, fy = np.floor(pts[:, 0]), np.floor(pts[:, 1])
    x0, y0 = int(fx.min()), int(fy.min())
    bw, bh = int(fx.max()) - x0 + 1, int(fy.max()) - y0 + 1
    xa, ya, xb, yb = x0 - border[0], y0 - border[1], x0 + bw + border[0], y0 + bh + border[1]
    return max(xa, 1), min(xb, imshape[1]), max(ya, 1), min(yb, imshape[0])


def lk_forward_backward(im1, im2, p1, fbt=None, nthreads=0, **lk):
    """(p2, v, err): forward LK, and when fbt is given a backward pass from p2 with
    v = st_fwd & st_bwd & (||p1 - p1'||_2 < fbt)."""
    p1 = np.asarray(p1, np.float32)
    p2, st, err = cvo.calcOpticalFlowPyrLK(im1, im2, p1, nthreads=nthreads, **lk)
    v = st.ravel() != 0
    if fbt is not None:
        back, st2, _ = cvo.calcOpticalFlowPyrLK(im2, im1, p2, nthreads=nthreads, **lk)
        d = p1 - back
        fbe = np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1])  # float32 throughout, as numpy evaluates it
        v = v & (st2.ravel() != 0) & (fbe < np.float32(fbt))
    return p2, v, err


def klt_regional(im0, im, p0, T, lk, fbt=1.0, translate=False, nthreads=0):
    T = np.asarray(T).astype(np.float32)
    x0, x1, y0, y1 = bounding_rect(p0, im.shape, border=(50, 50))
    roi_prev = im0[y0:y1, x0:x1]
    origin = np.float32([x0, y0])
    if translate:
        dx, dy = int(T[2, 0]), int(T[2, 1])
        roi_next = im[y0 + dy:y1 + dy, x0 + dx:x1 + dx]
        if roi_next.shape != roi_prev.shape:
            raise ValueError("shifted ROI leaves the frame (cv2 would assert on mismatched pyramid sizes)")
    else:
        roi_next = cvo.remap_affine(im, T, x0, x1, y0, y1)
    pa, v, _ = lk_forward_backward(roi_prev, roi_next, p0 - origin, fbt=fbt, nthreads=nthreads, **lk)
    if translate:
        p = pa + (origin + [dx, dy]).astype(np.float32)
    else:
        q = pa + origin
        p = np.concatenate([q, np.ones((q.shape[0], 1), q.dtype)], 1) @ T
    return p, v


LK_COARSE = dict(winSize=(15, 15), maxLevel=4, criteria=(3, 10, 0.1))
LK_FINE = dict(winSize=(51, 51), maxLevel=0, criteria=(3, 30, 0.001))


def klt_main(im, im0, im0_small, p0, nthreads=0):
    from .ransac_oracle import estimate_affine_2d

    p0 = np.asarray(p0, np.float32)
    s = 1 / 4
    im_small = cvo.decimate4(im)
    if im0_small is None:
        im0_small = cvo.decimate4(im0)
    p, v, _ = lk_forward_backward(im0_small, im_small, p0 * s, nthreads=nthreads, **LK_COARSE)
    p = p / s
    T23, inl = estimate_affine_2d(p0[v], p[v])
    v[v] = inl.ravel().astype(bool)
    T = np.eye(3, 2)
    T[2] = (p[v] - p0[v]).mean(0)
    p, v = klt_regional(im0, im, p0, T, LK_COARSE, fbt=1, translate=True, nthreads=nthreads)
    if v.sum() > 10:
        T23, inl = estimate_affine_2d(p0[v], p[v])
    else:
        raise RuntimeError("KLT coarse-affine failure (descriptor fallback is exercised separately)")
    p, v = klt_regional(im0, im, p0, T23.T, LK_FINE, fbt=0.3, nthreads=nthreads)
    return p[v], v, im_small
