"""Robust affine fit on the GPU: the call the tracker makes after every LK stage,

    T23, inliers = cv2.estimateAffine2D(p0[v], p[v], method=cv2.RANSAC)              utils/KLT.py:116,127 (and :33)

with cv2's signature, defaults and return layout.  OpenCV's RANSAC loop (fixed-seed RNG, 3-point closed-form model,
float32 residuals, adaptive iteration budget) runs in one CTA of libvelocity_b200.so (K10, csrc/ransac.cu): the inlier
mask is bit-identical to cv2's, T agrees to ~1e-12 (identical after the float32 cast KLTregional applies).  No CPU
fallback.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import require_cuda, stream_ptr

RANSAC = 8   # cv2.RANSAC


def estimateAffine2D(from_, to, inliers=None, method=RANSAC, ransacReprojThreshold=3.0, maxIters=2000, confidence=0.99, refineIters=10):
    """Returns (T float64 [2,3] or None, inliers uint8 [n,1]) like cv2.estimateAffine2D."""
    require_cuda()
    if method != RANSAC:
        raise NotImplementedError("only method=cv2.RANSAC (what the reference passes, utils/KLT.py:116) is implemented")
    a = np.ascontiguousarray(np.asarray(from_, np.float32).reshape(-1, 2))
    b = np.ascontiguousarray(np.asarray(to, np.float32).reshape(-1, 2))
    n = a.shape[0]
    if b.shape[0] != n:
        raise ValueError("estimateAffine2D: point sets differ in length (%d vs %d)" % (n, b.shape[0]))
    if n < 3:
        return None, None                      # cv2: fewer correspondences than the model needs
    packed = torch.from_numpy(np.concatenate([a.ravel(), b.ravel()])).cuda()       # one H2D copy
    out = torch.empty(6 * 8 + 3 * 4 + n, dtype=torch.uint8, device=packed.device)   # T | info | mask: one D2H copy
    base = out.data_ptr()
    _lib.check(_lib.lib().vel_estimate_affine2d_ransac(C.c_void_p(packed.data_ptr()), C.c_void_p(packed.data_ptr() + 8 * n), n,
                                                       float(ransacReprojThreshold), float(confidence), int(maxIters),
                                                       1 if refineIters else 0, C.c_void_p(base + 60), C.c_void_p(base),
                                                       C.c_void_p(base + 48), stream_ptr()), "vel_estimate_affine2d_ransac")
    host = out.cpu().numpy()
    T = host[:48].view(np.float64).reshape(2, 3).copy()
    info = host[48:60].view(np.int32)
    mask = host[60:].reshape(n, 1).copy()
    if not info[0]:
        return None, mask
    return T, mask
