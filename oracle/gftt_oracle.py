"""CPU oracle (TEST INFRASTRUCTURE ONLY -- never imported by velocity_b200/) for the feature
initialisation of the reference, vidExample.py:110:

    cv2.goodFeaturesToTrack(roi, 1000, 0.01, 0, blockSize=5, useHarrisDetector=True)   (k = 0.04, Sobel 3)

The arithmetic lives in opencv-python (un-vendored, unpinned: requirements.txt:5; 4.13.0.92 here).  It was pinned
by black-box comparison against cv2 4.13.0 on this image's x86-64 build and is restated below step by step; every
step is bit-exact against cv2 (tests/test_oracle_gftt.py) with ONE documented exception: in the LAST image row the
final (width mod 16) response values come from cv2's scalar / narrower-SIMD tail code whose rounding depends on
the CPU dispatch -- they differ by <= 1 ulp, are never corner candidates (the detector skips the 1-px frame) and
did not change the selected corners on any input tried.

  1. Sobel pair, CV_32F, scale s = 1 / (4 * blockSize * 255), BORDER_REFLECT_101, float32 kernel (1, 2, 1) * s:
       Dx = fma(r[y-1] + r[y+1], k1, round(r[y] * k0)),  r = p[x+1] - p[x-1]                    (column pass fused)
       Dy = t[y+1] - t[y-1],  t = fma(p[x+1], k1, fma(p[x], k0, round(p[x-1] * k1)))            (row pass fused) for
            columns < 32 * (w // 32), and the unfused ((p[x-1]*k1 + p[x]*k0) + p[x+1]*k1) in the remaining columns
  2. cov = (Dx*Dx, Dx*Dy, Dy*Dy) in float32
  3. 5x5 box sums, BORDER_REFLECT_101, accumulated in float64: row sums left to right, then cv2's running column
     sum (SUM += entering row; out = float32(SUM); SUM -= leaving row) from the top of the image
  4. R = (a*c - b*b) - k32 * ((a + c) * (a + c)), float32, no fused operations
  5. threshold: keep R > float32(max(R) * quality) (cv2.threshold THRESH_TOZERO), 3x3 local maxima (R == dilate(R),
     R != 0) inside the 1-px frame, ordered by value descending, ties by HIGHER raster address first, first
     maxCorners kept (minDistance = 0: no spacing filter).  Returns (x, y) float32 like cv2.
"""
import numpy as np

f32, f64 = np.float32, np.float64


def _reflect101_pad(a, n):
    return np.pad(a, n, mode="reflect")


def _fma(a, b, c):
    """float32 fma(a, b, c): the float64 product of two float32 is exact, one rounding to float32 at the end
    (the float64 addition is exact here because |a*b| and |c| are within 2^24 of each other or one is 0; the
    test-suite pins the result against cv2 bit for bit)."""
    return (a.astype(f64) * f64(b) + c.astype(f64)).astype(f32)


def sobel_pair(im, block_size=5):
    scale = 1.0 / (4 * block_size * 255.0)
    p = _reflect101_pad(np.asarray(im, np.uint8), 1).astype(f32)
    ky = np.array([1, 2, 1], f32) * f32(scale)
    k1, k0 = ky[0], ky[1]
    r = p[:, 2:] - p[:, :-2]
    dx = _fma(r[:-2] + r[2:], k1, (r[1:-1].astype(f64) * f64(k0)).astype(f32))
    s0, s1, s2 = p[:, :-2], p[:, 1:-1], p[:, 2:]
    t = _fma(s2, k1, _fma(s1, k0, s0 * k1))
    t0 = (im.shape[1] // 32) * 32
    t[:, t0:] = ((s0 * k1 + s1 * k0) + s2 * k1)[:, t0:]
    dy = t[2:] - t[:-2]
    return dx, dy


def box5_sum(c):
    p = _reflect101_pad(c, 2).astype(f64)
    rs = (((p[:, 0:-4] + p[:, 1:-3]) + p[:, 2:-2]) + p[:, 3:-1]) + p[:, 4:]
    h = c.shape[0]
    out = np.empty(c.shape, f32)
    acc = np.zeros(c.shape[1], f64)
    for i in range(4):
        acc = acc + rs[i]
    for y in range(h):
        s0 = acc + rs[y + 4]
        out[y] = s0.astype(f32)
        acc = s0 - rs[y]
    return out


def harris_response(im, block_size=5, k=0.04):
    dx, dy = sobel_pair(im, block_size)
    a, b, c = box5_sum(dx * dx), box5_sum(dx * dy), box5_sum(dy * dy)
    ac = a + c
    return (a * c - b * b) - f32(k) * (ac * ac)


def good_features_to_track(im, max_corners=1000, quality=0.01, block_size=5, k=0.04, response=None):
    """Returns float32 [n, 2] (x, y) in cv2's order."""
    R = harris_response(im, block_size, k) if response is None else response
    h, w = R.shape
    thr = f32(f64(R.max()) * quality)
    eig = np.where(R > thr, R, f32(0))
    pad = np.pad(eig, 1, mode="constant", constant_values=-np.inf)
    dil = np.max(np.stack([pad[dy:dy + h, dx:dx + w] for dy in range(3) for dx in range(3)]), axis=0)
    cand = (eig != 0) & (eig == dil)
    cand[0, :] = cand[-1, :] = False
    cand[:, 0] = cand[:, -1] = False
    ys, xs = np.nonzero(cand)
    vals = eig[ys, xs]
    addr = ys.astype(np.int64) * w + xs
    order = np.lexsort((-addr, -vals.astype(f64)))
    order = order[:max_corners] if max_corners > 0 else order
    return np.stack([xs[order], ys[order]], 1).astype(f32)
