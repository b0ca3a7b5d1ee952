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


def estimateAffine2D_masked_device(from_dev, to_dev, mask_dev, to_scale=1.0, to_off=(0.0, 0.0), ransacReprojThreshold=3.0, maxIters=2000,
                                   confidence=0.99, refineIters=10):
    """`T, inl = estimateAffine2D(from_[v], to[v]); v[v] = inl` (utils/KLT.py:116-117) on DEVICE data in one launch, nothing synchronised:
    from_dev / to_dev float32 [n,2] CUDA, mask_dev uint8 [n] CUDA (the LK status).  `to` is first mapped as to * to_scale + to_off in
    float32.  Returns CUDA tensors (mask_out uint8 [n], to_mapped float32 [n,2], T float64 [6], info int32 [4] = found, inliers,
    iterations, rows kept)."""
    require_cuda()
    n = from_dev.shape[0]
    dev = from_dev.device
    L = _lib.lib()
    nbytes = int(L.vel_estimate_affine2d_ransac_masked_workspace(n))
    work = torch.empty((nbytes + 7) // 8, dtype=torch.float64, device=dev)
    mask_out = torch.empty((n,), dtype=torch.uint8, device=dev)
    to_mapped = torch.empty((n, 2), dtype=torch.float32, device=dev)
    tail = torch.empty(8, dtype=torch.float64, device=dev)                 # T (6 doubles) | info (4 int32)
    _lib.check(L.vel_estimate_affine2d_ransac_masked(C.c_void_p(from_dev.data_ptr()), C.c_void_p(to_dev.data_ptr()), C.c_void_p(mask_dev.data_ptr()), n,
                                                     float(to_scale), float(to_off[0]), float(to_off[1]), float(ransacReprojThreshold),
                                                     float(confidence), int(maxIters), 1 if refineIters else 0, C.c_void_p(work.data_ptr()),
                                                     work.numel() * 8, C.c_void_p(mask_out.data_ptr()), C.c_void_p(to_mapped.data_ptr()),
                                                     C.c_void_p(tail.data_ptr()), C.c_void_p(tail.data_ptr() + 48), stream_ptr()),
               "vel_estimate_affine2d_ransac_masked")
    return mask_out, to_mapped, tail
