"""Per-frame cost of the reference's 3-stage tracker KLTmain (utils/KLT.py:99-134) on the GPU path, with the robust fit
(K10) timed separately against cv2 on the host:  python tools/kltmain_bench.py [tracks]"""
import sys
import time

import cv2
import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import KLT, ransac, synth  # noqa: E402

NT = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
frames, _ = synth.plane_sequence(8, seed=77, Z0=40.0)
dev = [torch.from_numpy(f).cuda() for f in frames]
p0 = synth.harris_tracks(frames[0], NT)
small = None
for rep in range(2):
    p, small = p0, None
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(1, len(frames)):
        q, v, small = KLT.KLTmain(dev[i], dev[i - 1], small, p)
        p = q
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / (len(frames) - 1)
print("KLTmain (GPU path, frames resident, %d tracks at start, %d at end): %.2f ms/frame wall" % (NT, len(p), ms))

rng = np.random.default_rng(1)
for n in (170, 1000, 4096):
    fr = rng.uniform(0, 1900, (n, 2)).astype(np.float32)
    to = (fr * 1.01 + 3 + rng.normal(size=(n, 2)) * 0.3).astype(np.float32)
    to[rng.choice(n, n // 5, replace=False)] += rng.uniform(-60, 60, (n // 5, 2)).astype(np.float32)
    for _ in range(3):
        T, inl = ransac.estimateAffine2D(fr, to)
    t0 = time.perf_counter()
    for _ in range(20):
        T, inl = ransac.estimateAffine2D(fr, to)
    g = (time.perf_counter() - t0) / 20 * 1e3
    t0 = time.perf_counter()
    for _ in range(20):
        Tc, ic = cv2.estimateAffine2D(fr, to, method=cv2.RANSAC)
    c = (time.perf_counter() - t0) / 20 * 1e3
    print("estimateAffine2D n=%d (20%% outliers): GPU %.3f ms wall incl. H2D/D2H, cv2 %.3f ms, masks equal %s" % (n, g, c, np.array_equal(inl, ic)))
