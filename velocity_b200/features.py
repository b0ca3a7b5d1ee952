"""Feature initialisation on the GPU: the detector call of the reference's driver,

    p = cv2.goodFeaturesToTrack(roi, 1000, 0.01, 0, blockSize=5, useHarrisDetector=True)      vidExample.py:110

with cv2's signature and return layout ([n, 1, 2] float32, strongest first).  The Harris response, the 3x3
non-maximum suppression and the ordered top-K selection all run in libvelocity_b200.so (K9, csrc/features.cu) and
return the same corners in the same order as opencv-python 4.13 (tests/test_features_gpu.py).  The sub-pixel
refinement that follows in the reference (cv2.cornerSubPix, vidExample.py:113-115) is cornerSubPix below: one
thread per corner, bit-identical to cv2.  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from .device import image_view, ptr, require_cuda, stream_ptr


def harris_corners_device(image, maxCorners, qualityLevel, blockSize=5, k=0.04, want_response=False):
    """Device-resident core: returns (xy CUDA float32 [maxCorners, 2], count CUDA int32 [1], response or None)."""
    require_cuda()
    t, p, w, h, pitch = image_view(image)
    L = _lib.lib()
    nbytes = L.vel_good_features_workspace(w, h, int(maxCorners))
    work = torch.empty(nbytes, dtype=torch.uint8, device=t.device)
    out = torch.empty((int(maxCorners), 2), dtype=torch.float32, device=t.device)
    cnt = torch.zeros(1, dtype=torch.int32, device=t.device)
    resp = torch.empty((h, w), dtype=torch.float32, device=t.device) if want_response else None
    _lib.check(L.vel_good_features_harris_u8(p, w, h, pitch, int(maxCorners), float(qualityLevel), int(blockSize), float(k), ptr(work),
                                             nbytes, ptr(resp), ptr(out), ptr(cnt), stream_ptr()), "vel_good_features_harris_u8")
    return out, cnt, resp


def goodFeaturesToTrack(image, maxCorners, qualityLevel, minDistance, mask=None, blockSize=3, useHarrisDetector=False, k=0.04):
    """cv2.goodFeaturesToTrack for the configuration the reference uses (Harris detector, blockSize 5,
    minDistance 0, no mask).  Returns float32 [n, 1, 2] like cv2 (None when nothing passes, like cv2)."""
    if mask is not None or minDistance != 0 or not useHarrisDetector or blockSize != 5 or maxCorners <= 0:
        raise NotImplementedError("only the reference's call is implemented: Harris detector, blockSize=5, minDistance=0, "
                                  "no mask, maxCorners > 0 (vidExample.py:110)")
    out, cnt, _ = harris_corners_device(image, maxCorners, qualityLevel, blockSize, k)
    packed = torch.cat([out.reshape(-1), cnt.to(torch.float32)]).cpu().numpy()      # one D2H copy
    n = int(packed[-1])
    if n == 0:
        return None
    return np.ascontiguousarray(packed[:2 * n].reshape(n, 1, 2))


def cornerSubPix(image, corners, winSize, zeroZone, criteria):
    """cv2.cornerSubPix for the reference's call (vidExample.py:113-115): uint8 image, zeroZone (-1,-1).  Returns the
    refined corners as a new float32 array of the input's shape (cv2 also refines its argument in place; numpy inputs
    are left untouched here)."""
    require_cuda()
    if tuple(zeroZone) != (-1, -1):
        raise NotImplementedError("only zeroZone=(-1,-1) (the reference's value, vidExample.py:114) is implemented")
    t, p, w, h, pitch = image_view(image)
    c = np.asarray(corners, np.float32)
    pts = torch.from_numpy(np.ascontiguousarray(c.reshape(-1, 2))).to(t.device)
    ctype, count, eps = criteria
    _lib.check(_lib.lib().vel_corner_subpix_u8(p, w, h, pitch, ptr(pts), pts.shape[0], int(winSize[0]), int(winSize[1]),
                                               int(count) if (ctype & 1) else 100, float(eps) if (ctype & 2) else 0.0, stream_ptr()),
               "vel_corner_subpix_u8")
    return pts.cpu().numpy().reshape(c.shape)
