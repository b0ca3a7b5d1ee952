import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def lk_kwargs(g):
    return dict(winSize=tuple(int(x) for x in g["win"]), maxLevel=int(g["max_level"]),
                criteria=(3, int(g["max_count"]), float(g["eps"])))


def lk_images(g):
    if "roi0" in g:
        a, b = g["roi0"], g["roi1"]
        return g["im0"][a[0]:a[1], a[2]:a[3]], g["im1"][b[0]:b[1], b[2]:b[3]]
    return g["im0"], g["im1"]


def fbt_of(g):
    f = float(g["fbt"])
    return None if f < 0 else f


LK_CASES = ["lk_coarse", "lk_coarse_fb", "lk_c2_fb", "lk_fine_fb", "lk_odd_21", "lk_rect_win", "lk_roi_views"]
PRIM_CASES = ["prim_even", "prim_odd", "prim_tiny"]

# Tolerances (SURVEY.md 8(c) "Tolerance policy")
LK_POINT_TOL_PX = 2e-3   # vs cv2 (golden): cv2 accumulates the 2x2 system in float32 SIMD lanes
LK_ERR_TOL = 1e-2


def lk_sweep_cases():
    """Seeded sweep over cv2.calcOpticalFlowPyrLK's parameter space (window, levels, iteration cap, eps, minimum
    eigenvalue threshold, forward-backward threshold) on small-shift / large-shift / unrelated pairs, with points
    anywhere (also outside the frame).  Shared by the CPU test (oracle vs cv2) and the GPU test (CUDA vs oracle)."""
    from velocity_b200 import synth

    rng = np.random.default_rng(2024)
    base = synth.texture(260, 372, 31)
    pairs = [(base, np.roll(base, (1, -2), (0, 1))), (base, np.roll(base, (-7, 9), (0, 1))), (base, synth.texture(260, 372, 32))]
    wins = [(15, 15), (15, 15), (15, 15), (21, 21), (9, 13), (33, 17)]
    for k in range(18):
        im0, im1 = pairs[k % 3]
        lk = dict(winSize=wins[k % len(wins)], maxLevel=int(rng.integers(0, 5)),
                  criteria=(3, int(rng.choice([1, 3, 10, 30])), float(rng.choice([0.5, 0.03, 0.001]))),
                  minEigThreshold=float(rng.choice([1e-4, 1e-4, 3e-3])))
        fbt = [None, 0.2, 2.0][int(rng.integers(0, 3))]
        h, w = im0.shape
        pts = np.concatenate([synth.harris_tracks(im0, 120, border=3),
                              np.stack([rng.uniform(-20, w + 20, 60), rng.uniform(-20, h + 20, 60)], 1)]).astype(np.float32)
        yield k, im0, im1, pts, lk, fbt


def ba_c3_inputs():
    """Seeded bundle-adjustment problem of BASELINE configs[2] size (nt=4096, nc=299); tests/golden/ba_c3_sparse.npz holds
    the oracle's result for exactly these inputs (CRCs stored beside it)."""
    from velocity_b200 import synth

    K = synth.K_1080P.copy()
    nt, nf = 4096, 300
    pw = synth.scene_points(nt, seed=7)
    P, cw = synth.scene_observations(pw, nf, step=0.02, noise=0.1, seed=11)
    rng = np.random.default_rng(3)
    pw0 = pw + rng.normal(0, 0.05, pw.shape)
    cw0 = cw + rng.normal(0, 0.01, cw.shape)
    cw0[0] = 0
    return K, P, pw0, cw0
