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
