"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by velocity_b200/) for the robust affine fit of the tracker,

    T23, inliers = cv2.estimateAffine2D(p0[v], p[v], method=cv2.RANSAC)            utils/KLT.py:116,127 (and :33)

with cv2's defaults (threshold 3 px, 2000 iterations, confidence 0.99, 10 refinement iterations).  The arithmetic is
OpenCV's (calib3d ptsetreg.cpp; un-vendored, unpinned: requirements.txt:5; 4.13.0.92 here), restated from its public
algorithm and pinned by black-box comparison -- tests/golden/ransac.npz and, where opencv-python is importable, cv2
itself on the test host:

  * RNG: cv::RNG seeded with 2^64-1, multiply-with-carry x <- (uint32)x * 4164903690 + (x >> 32), uniform(a,b) = a + next % (b-a)
  * subset: three distinct indices (re-draw on a repeat); rejected when the third point is (nearly) collinear with
    the first two in either point set: |dx2*dy1 - dy2*dx1| <= FLT_EPSILON*(|dx1|+|dy1|+|dx2|+|dy2|), differences in float32
  * model: the closed-form 3-point affine in float64 (Cramer), cast to float32 for the residuals
    e = ((F0*x + F1*y) + F2 - X)^2 + ((F3*x + F4*y) + F5 - Y)^2 in float32; inlier <=> e <= float32(thr^2)
  * a model is kept when it has more inliers than max(best, 2); the iteration budget shrinks to
    round(log(1-conf) / log(1 - (1-outlier_ratio)^3)) after every improvement
  * refinement: cv2 runs 10 Levenberg-Marquardt steps on the inliers from the kept model; the problem is linear, so that
    converges to the least-squares affine of the inliers -- restated as a centred float64 least-squares fit.

Parity: inlier masks bit-exact (150/150 random trials incl. 0-50 % outliers); T within 1e-9 of cv2 (observed 3.5e-13),
identical after the float32 cast KLTregional applies (utils/KLT.py:58).
"""
import math
import sys

import numpy as np

f32 = np.float32
FLT_EPS = float(np.finfo(np.float32).eps)


class CvRNG:
    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def _third_point_collinear(p):
    """cv::haveCollinearPoints for a 3-point subset: only the last point is tested against the first pair."""
    dx1, dy1 = float(f32(p[1][0]) - f32(p[2][0])), float(f32(p[1][1]) - f32(p[2][1]))
    dx2, dy2 = float(f32(p[0][0]) - f32(p[2][0])), float(f32(p[0][1]) - f32(p[2][1]))
    return abs(dx2 * dy1 - dy2 * dx1) <= FLT_EPS * (abs(dx1) + abs(dy1) + abs(dx2) + abs(dy2))


def _affine_from_3(fr, to):
    (x1, y1), (x2, y2), (x3, y3) = [(float(a), float(b)) for a, b in fr]
    (X1, Y1), (X2, Y2), (X3, Y3) = [(float(a), float(b)) for a, b in to]
    d = 1. / (x1 * (y2 - y3) + x2 * (y3 - y1) + x3 * (y1 - y2))
    return [d * (X1 * (y2 - y3) + X2 * (y3 - y1) + X3 * (y1 - y2)), d * (X1 * (x3 - x2) + X2 * (x1 - x3) + X3 * (x2 - x1)),
            d * (X1 * (x2 * y3 - x3 * y2) + X2 * (x3 * y1 - x1 * y3) + X3 * (x1 * y2 - x2 * y1)),
            d * (Y1 * (y2 - y3) + Y2 * (y3 - y1) + Y3 * (y1 - y2)), d * (Y1 * (x3 - x2) + Y2 * (x1 - x3) + Y3 * (x2 - x1)),
            d * (Y1 * (x2 * y3 - x3 * y2) + Y2 * (x3 * y1 - x1 * y3) + Y3 * (x1 * y2 - x2 * y1))]


def _residuals(M, fr, to):
    F = [f32(m) for m in M]
    a = (F[0] * fr[:, 0] + F[1] * fr[:, 1]) + F[2] - to[:, 0]
    b = (F[3] * fr[:, 0] + F[4] * fr[:, 1]) + F[5] - to[:, 1]
    return a * a + b * b


def _update_iters(p, ep, model_points, max_iters):
    p, ep = min(max(p, 0.), 1.), min(max(ep, 0.), 1.)
    num, denom = max(1. - p, sys.float_info.min), 1. - pow(1. - ep, model_points)
    if denom < sys.float_info.min:
        return 0
    num, denom = math.log(num), math.log(denom)
    return max_iters if (denom >= 0 or -num >= max_iters * (-denom)) else int(np.rint(num / denom))


def ransac_affine(fr, to, thresh=3.0, conf=0.99, max_iters=2000):
    """The RANSAC stage: (model [6] float64 or None, inlier mask bool [n] or None, iterations run)."""
    fr, to = np.asarray(fr, f32).reshape(-1, 2), np.asarray(to, f32).reshape(-1, 2)
    count = len(fr)
    if count < 3:
        return None, None, 0
    if count == 3:
        return _affine_from_3(fr, to), np.ones(3, bool), 0
    rng, niters, best, best_mask, max_good = CvRNG(), max(max_iters, 1), None, None, 0
    t = f32(thresh * thresh)
    it = 0
    while it < niters:
        found = False
        for _ in range(10000):
            idx = []
            for _i in range(3):
                v = rng.uniform(0, count)
                while v in idx:
                    v = rng.uniform(0, count)
                idx.append(v)
            if _third_point_collinear(fr[idx]) or _third_point_collinear(to[idx]):
                continue
            found = True
            break
        if not found:
            if it == 0:
                return None, None, 0
            break
        M = _affine_from_3(fr[idx], to[idx])
        mask = _residuals(M, fr, to) <= t
        good = int(mask.sum())
        if good > max(max_good, 2):
            best, best_mask, max_good = M, mask, good
            niters = _update_iters(conf, (count - good) / count, 3, niters)
        it += 1
    return best, best_mask, it


def estimate_affine_2d(fr, to, thresh=3.0, conf=0.99, max_iters=2000, refine=True):
    """cv2.estimateAffine2D(fr, to, method=cv2.RANSAC): (T [2,3] float64, inliers uint8 [n,1]) or (None, None)."""
    fr, to = np.asarray(fr, f32).reshape(-1, 2), np.asarray(to, f32).reshape(-1, 2)
    M, mask, _ = ransac_affine(fr, to, thresh, conf, max_iters)
    if M is None:
        return None, (np.zeros((len(fr), 1), np.uint8) if len(fr) >= 3 else None)     # cv2 leaves an all-zero mask behind
    T = np.array(M, np.float64).reshape(2, 3)
    if refine and len(fr) > 3 and mask.any():
        x, y = fr[mask].astype(np.float64), to[mask].astype(np.float64)
        mx, my = x.mean(0), y.mean(0)
        xc, yc = x - mx, y - my
        A = np.linalg.solve(xc.T @ xc, xc.T @ yc).T          # centred normal equations: y - my = A (x - mx)
        T = np.concatenate([A, (my - A @ mx)[:, None]], 1)
    return T, mask.astype(np.uint8).reshape(-1, 1)
