"""Seeded synthetic inputs (SURVEY.md section 8(d) "Synthetic sequence generator").

Host-side input generation only -- nothing here is on the accelerated path.  A textured plane at
depth Z0 approaches the camera by v*dt per frame (the reference's own 3-D model is planar:
vidExample.py:119), rendered with cv2.warpPerspective; tracks are the N strongest Harris corners
of frame 0 (the detector the reference uses at vidExample.py:110).
"""
import numpy as np

K_1080P = np.array([[1700.0, 0, 0], [0, 1700.0, 0], [960.5, 540.5, 1]])  # row-vector convention (utils/images.py:148)


def texture(h, w, seed, sigma=2.0):
    """uint8 band-limited noise: uniform noise -> Gaussian blur(sigma) -> min/max stretch."""
    import cv2

    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w), dtype=np.uint8)
    a = cv2.GaussianBlur(a, (0, 0), sigma).astype(np.float32)
    a = (a - a.min()) / (a.max() - a.min()) * 255.0
    return a.astype(np.uint8)


def plane_homography(K_row, Z0, Z, tex_scale, tex_center):
    """Texture pixel -> image pixel for a fronto-parallel plane at depth Z whose texture has
    `tex_scale` pixels per metre at reference depth Z0 (identity scale at Z == Z0)."""
    f = K_row[0, 0]
    cx, cy = K_row[2, 0], K_row[2, 1]
    s = Z0 / Z
    # image = c + s * (tex - tex_center)
    H = np.array([[s, 0, cx - s * tex_center[0]], [0, s, cy - s * tex_center[1]], [0, 0, 1.0]])
    return H


def plane_sequence(n_frames, h=1080, w=1920, seed=2025, Z0=10.0, v_kmh=40.0, dt=1 / 29.97, margin=64, K_row=None,
                   reset_every=24):
    """List of n_frames uint8 [h,w] frames of the approaching plane, depths Z[i] (metres).
    Depth is reset every `reset_every` frames so tracks survive long sequences (stated in DESIGN.md)."""
    import cv2

    K_row = K_1080P if K_row is None else K_row
    tex = texture(h + 2 * margin, w + 2 * margin, seed)
    tc = ((w + 2 * margin - 1) / 2.0, (h + 2 * margin - 1) / 2.0)
    frames, depths = [], []
    step = v_kmh / 3.6 * dt
    for i in range(n_frames):
        Z = Z0 - step * (i % reset_every)
        H = plane_homography(K_row, Z0, Z, 1.0, tc)
        frames.append(cv2.warpPerspective(tex, H, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101))
        depths.append(Z)
    return frames, np.array(depths)


def harris_tracks(im, n, quality=0.001, min_distance=5, border=32):
    """N strongest Harris corners (float32 [n,2], x,y), kept `border` px away from the frame edge."""
    import cv2

    p = cv2.goodFeaturesToTrack(im, 0, quality, min_distance, blockSize=5, useHarrisDetector=True).reshape(-1, 2)
    h, w = im.shape
    keep = (p[:, 0] > border) & (p[:, 0] < w - border) & (p[:, 1] > border) & (p[:, 1] < h - border)
    p = p[keep]
    if p.shape[0] < n:
        rng = np.random.default_rng(n)
        extra = np.stack([rng.uniform(border, w - border, n - p.shape[0]), rng.uniform(border, h - border, n - p.shape[0])], 1)
        p = np.concatenate([p, extra.astype(np.float32)])
    return np.ascontiguousarray(p[:n], dtype=np.float32)


def scene_points(n, seed=7):
    """3-D tie points of the survey's scratch scene: X,Y in U(-2,2)xU(-1,1), Z in U(8,12)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-2, 2, n), rng.uniform(-1, 1, n), rng.uniform(8, 12, n)], 1)


def scene_observations(pw, n_cams, K_row=None, step=0.37, noise=0.1, seed=11, rpy_sigma=0.0):
    """Project `pw` into n_cams cameras translating by (0,0,-step*j) (row-vector convention
    p_cam = p_world @ R + t, utils/NLS.py:213).  Returns P [5, nt, n_cams] float32 laid out like the
    reference's bookkeeping array (vidExample.py:128,151-153), camera positions cw [n_cams,3]."""
    K_row = K_1080P if K_row is None else K_row
    rng = np.random.default_rng(seed)
    nt = pw.shape[0]
    P = np.full((5, nt, n_cams), np.nan, np.float32)
    cw = np.zeros((n_cams, 3))
    for j in range(n_cams):
        cw[j] = (0.0, 0.0, -step * j)
        pc = pw + cw[j]
        uv = (pc @ K_row)
        uv = uv[:, :2] / uv[:, 2:3] + rng.normal(0, noise, (nt, 2))
        P[0:2, :, j] = uv.T
        P[4, :, j] = j
    return P, cw


# ---- connected C3 sequence: one continuous approach, no depth reset ----------------------------------------------
def approach_sequence(n_frames, h=1080, w=1920, seed=2025, z_start=200.0, v_kmh=40.0, dt=1 / 29.97, sigma=2.0, margin=64,
                      K_row=None):
    """n_frames uint8 [h,w] frames of the SURVEY 8(d) textured plane approaching the camera by v*dt per frame in ONE
    run from depth z_start (no reset: tracks are propagated frame to frame through the whole sequence, the way
    vidExample.py:134-135 does).  The texture is defined at the mid-sequence depth so that the rendering minifies
    (first half) and magnifies (second half) by the same factor.  Returns (frames [n,h,w] uint8, depths [n])."""
    import cv2

    K_row = K_1080P if K_row is None else K_row
    step = v_kmh / 3.6 * dt
    Z = z_start - step * np.arange(n_frames)
    if Z[-1] <= 0:
        raise ValueError("approach_sequence: the plane passes the camera (z_start too small for %d frames)" % n_frames)
    z_ref = float(Z[n_frames // 2])
    s_min = z_ref / Z.max()
    th, tw = int(h / s_min) + 2 * margin, int(w / s_min) + 2 * margin
    tex = texture(th, tw, seed, sigma)
    tc = ((tw - 1) / 2.0, (th - 1) / 2.0)
    frames = np.empty((n_frames, h, w), np.uint8)
    for i in range(n_frames):
        Hm = plane_homography(K_row, z_ref, Z[i], 1.0, tc)
        frames[i] = cv2.warpPerspective(tex, Hm, (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    return frames, Z


def approach_tracks(frame0, n, zoom, K_row=None, border=40):
    """The n strongest Harris corners of frame 0 (quality 0.001, minDistance 5, blockSize 5) among those
    that stay `border` px inside the frame after the image has grown by `zoom` about the principal point."""
    import cv2

    K_row = K_1080P if K_row is None else K_row
    h, w = frame0.shape
    cx, cy = K_row[2, 0], K_row[2, 1]
    p = cv2.goodFeaturesToTrack(frame0, 0, 0.001, 5, blockSize=5, useHarrisDetector=True).reshape(-1, 2)
    bx, by = (min(cx, w - 1 - cx) - border) / zoom, (min(cy, h - 1 - cy) - border) / zoom
    p = p[(np.abs(p[:, 0] - cx) < bx) & (np.abs(p[:, 1] - cy) < by)]
    if p.shape[0] < n:
        raise ValueError("approach_tracks: only %d corners stay in frame (asked for %d)" % (p.shape[0], n))
    return np.ascontiguousarray(p[:n], dtype=np.float32)


def plane_fiducials(K_row, Z, size=(8.0, 4.0)):
    """Four corner marks of a size[0] x size[1] metre rectangle on the plane (the synthetic stand-in for the licence
    plate whose known geometry fixes the metric scale, vidExample.py:118): (pixels float32 [4,2], world [4,3]) with the
    corner order of utils/common.py:150-156."""
    corners = np.array([[1, -1, 0], [1, 1, 0], [-1, 1, 0], [-1, -1, 0]], np.float64) * np.array([size[0], size[1], 0.0]) / 2
    cam = corners + np.array([0.0, 0.0, Z])
    uv = cam @ K_row
    return (uv[:, :2] / uv[:, 2:3]).astype(np.float32), corners
