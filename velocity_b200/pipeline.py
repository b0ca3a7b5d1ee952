"""End-to-end speed estimation: the frame loop of the reference's vidExample.py (:75-170) driven
through this package's GPU modules.

    frame 0 : Harris corners (+ sub-pixel) in the plate ROI, 6-dof plate pose (NLS.fcnNLS_Rt), planar 3-D points
    frame i : KLT.KLTmain (3-stage GPU tracker) -> NLS.estimateWorldCameraPose (3-dof, GPU) -> speed
    frame 5 : MSV.fcnMSV1_t re-triangulates every track (GPU) and enables all points

Feature initialisation (vidExample.py:110-115): the Harris detector and the sub-pixel refinement run on
the GPU (K9, features.py); everything inside the per-frame loop that the reference delegates to cv2.calcOpticalFlowPyrLK / remap / numpy LM runs
in libvelocity_b200.so.  The bookkeeping arrays keep the reference's layout (P [5,npts,n],
B [n,14], S [n,9]) so downstream code (plots) can consume them unchanged.
"""
import time

import numpy as np

from . import KLT, MSV, NLS
from .common import addcol0, image2world, norm, worldPointsLicensePlate
from .images import boundingRect, insidebbox

HEADER = ("image", "procTime", "pointTracks", "metric", "dt", "time", "dx", "distance", "speed",
          "#", "(s)", "#", "(pixels)", "(s)", "(s)", "(m)", "(m)", "(km/h)")


def detect_plate_features(im, q, max_corners=1000, detector=None):
    """vidExample.py:107-116: Harris corners in the plate neighbourhood, refined to sub-pixel -- both on the GPU (K9).
    `detector` = (goodFeaturesToTrack, cornerSubPix) replaces them (the test-suite passes the CPU oracle's)."""
    if detector is None:
        from . import features

        detector = (features.goodFeaturesToTrack, features.cornerSubPix)
    gftt, subpix = detector
    boxa = boundingRect(q, im.shape, border=(0, 0))
    boxb = boundingRect(q, im.shape, border=(700, 500))
    roi = np.ascontiguousarray(im[boxb[2]:boxb[3], boxb[0]:boxb[1]])
    p = gftt(roi, max_corners, 0.01, 0, blockSize=5, useHarrisDetector=True).squeeze()
    p = p + np.float32([boxb[0], boxb[2]])
    p = subpix(im, p, (5, 5), (-1, -1), (2 + 1, 100, 0.001))     # cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER
    return np.concatenate((q, p)), boxa, boxb


def run_speed_estimation(frames, q, K, frame_times, msv_frame=5, verbose=True, country="Chile", modules=None):
    """frames: sequence of uint8 [H,W] images; q: float32 [4,2] plate corners in frame 0;
    K: 3x3 row-vector intrinsics; frame_times: seconds per frame.  Returns a dict with the
    reference's S / B / P arrays and the summary statistics it prints.  `modules` = (KLT, NLS, MSV, (gftt, subpix))
    lets the test-suite drive the same loop with the CPU oracle; the default is the GPU path."""
    klt, nls, msv, detector = modules if modules is not None else (KLT, NLS, MSV, None)
    n = len(frames)
    q = np.asarray(q, np.float32)
    B = np.zeros([n, 14], dtype=np.float32)
    S = np.zeros([n, 9], dtype=np.float32)
    if verbose:
        print(("\n" + "%13s" * 9) * 2 % HEADER)
    im0 = im0_small = None
    on_gpu = modules is None
    for i in range(n):
        tic = time.time()
        im = frames[i]
        B[i, 12] = frame_times[i]
        if on_gpu and i > 0:
            # one upload per frame: the tracker takes CUDA tensors in place, and this frame is next step's previous frame
            # (numpy frames were uploaded twice per KLTmain call)
            import torch

            im = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im)).cuda()
        if i == 0:
            p, boxa, boxb = detect_plate_features(im, q, detector=detector)
            t, R, residuals, _ = nls.estimateWorldCameraPose(K, q, worldPointsLicensePlate(country), findR=True)
            p3 = addcol0(image2world(K, R, t, p).astype(float)) @ R + t
            R = np.eye(3)
            B[0, 0:3] = t
            vg = np.ones(p.shape[0], dtype=bool)
            vp = insidebbox(p, boxa)
            p_ = p[vp]
            P = np.full([5, p.shape[0], n], np.nan, dtype=np.float32)
            dt, dr, r, t0 = np.nan, 0, 0, B[0, 12]
        else:
            p, v, im0_small = klt.KLTmain(im, im0, im0_small, p)
            vg[vg] = v
            vp = vp & vg
            t, R, residuals, p_ = nls.estimateWorldCameraPose(K, p[vp[vg]], p3[vp], R=R, findR=False)
            dt = B[i, 12] - B[i - 1, 12]
            dr = norm(t + B[0, 0:3] - B[i - 1, 0:3])
            r += dr
            B[i, 3:6] = t
            B[i, 0:3] = B[0, 0:3] + t
        P[0:2, vg, i] = p.T
        P[2:4, vp, i] = p_.T
        P[4, vg, i] = i
        if i == msv_frame:
            _tmsv, p3hat = msv.fcnMSV1_t(K, P, B, vg, i)
            p3[vg] = p3hat - t
            vp = vg
        S[i, :] = (i, time.time() - tic, vg.sum(), residuals, dt, B[i, 12] - t0, dr, r, dr / dt * 3.6)
        if verbose:
            print("{:13g}{:13.3f}{:13g}{:13.3f}{:13.3f}{:13.3f}{:13.2f}{:13.2f}{:13.1f}".format(*tuple(S[i, :])))
        if on_gpu and i == 0:
            import torch

            im = im if isinstance(im, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(im)).cuda()
        im0 = im
    out = dict(S=S, B=B, P=P, speed_mean=float(S[1:, 8].mean()), speed_std=float(S[1:, 8].std()),
               res_mean=float(S[1:, 3].mean()), tracks=S[:, 2].astype(int))
    if verbose:
        print(f"\nSpeed = {out['speed_mean']:.2f} +/- {out['speed_std']:.2f} km/h\nRes = {out['res_mean']:.3f} pixels")
    return out
