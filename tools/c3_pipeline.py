"""BASELINE config 3 shape on one GPU: 300 synthetic 1080p frames, 4096 tracks --
KLT (C2 parameters, fwd+bwd) on every consecutive pair, fcnNLS_t per frame, fcnNvintercept over all
frames, fcnNLS_batch (10 LM iterations, nt=4096, nc=299).  Prints per-stage device times.

    python tools/c3_pipeline.py [--frames 300] [--cpu]     (--cpu also times the oracle on a bounded sample)
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from velocity_b200 import MSV, NLS, synth  # noqa: E402
from velocity_b200.common import pixel2uvec  # noqa: E402
from velocity_b200.lk import FrameBatch, lk_params, track_pairs  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=300)
ap.add_argument("--tracks", type=int, default=4096)
ap.add_argument("--cpu", action="store_true")
args = ap.parse_args()
F, NT = args.frames, args.tracks
K = synth.K_1080P
LK = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))


def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


# ---- stage 1: KLT over the image sequence ------------------------------------------------------------
uniq, _ = synth.plane_sequence(24, seed=2025, Z0=40.0)
idx = [i % 24 if (i // 24) % 2 == 0 else 23 - (i % 24) for i in range(F)]
frames = torch.from_numpy(np.stack(uniq)[idx]).cuda()
pts = torch.from_numpy(synth.harris_tracks(uniq[0], NT)).cuda()
params = lk_params(fbt=1.0, **LK)
fb = FrameBatch(frames, LK["winSize"], LK["maxLevel"])


def klt():
    fb.build()
    return track_pairs(fb, fb, pts, params, 0, 1, F - 1)


t_klt, (p2, st, err, _) = timed(klt)
print("KLT  : %d pairs x %d tracks fwd+bwd   %8.2f ms  (%.1f us/pair, valid %.3f)" % (F - 1, NT, t_klt, t_klt * 1e3 / (F - 1), st.float().mean().item()))

# ---- synthetic geometry for the solver stages (SURVEY 8c.3 scratch scene, 0.02 m / frame) -------------
pw = synth.scene_points(NT, seed=7)
P, cw = synth.scene_observations(pw, F, step=0.02, noise=0.1, seed=11)
# stage 2: fcnNLS_t per frame (299 independent problems of 4096 points)
p_all = torch.from_numpy(np.ascontiguousarray(P[0:2].transpose(2, 1, 0).reshape(-1, 2)).astype(np.float64)).cuda()   # [F*NT, 2]
pw_all = torch.from_numpy(np.tile(pw, (F, 1))).cuda()
first = torch.arange(F, dtype=torch.int32, device="cuda") * NT
count = torch.full((F,), NT, dtype=torch.int32, device="cuda")
x0 = torch.zeros((F, 3), dtype=torch.float64, device="cuda")
Kd = torch.from_numpy(K).cuda()
t_nls, (xt, iters) = timed(lambda: NLS.nls_batch_device(Kd, p_all, pw_all, first, count, x0, 3))
terr = np.abs(xt.cpu().numpy() - cw).max()
print("NLS_t: %d frames x %d points           %8.2f ms  (%.1f us/frame, iterations %d..%d, |t - truth| max %.2e m)" % (
    F, NT, t_nls, t_nls * 1e3 / F, iters.min().item(), iters.max().item(), terr))

# stage 3: N-view triangulation over all frames
U = np.zeros((3, F, NT))
for j in range(F):
    U[:, j] = pixel2uvec(K, P[0:2, :, j].T.astype(np.float64)).T
dU, dA = torch.from_numpy(U).cuda(), torch.from_numpy(-cw).cuda()
t_tri, C0 = timed(lambda: MSV.fcnNvintercept(dA, dU))
print("triNv: %d rays x %d points             %8.2f ms  (|C0 - truth| max %.2e m)" % (F, NT, t_tri, np.abs(C0.cpu().numpy() - pw).max()))
if F <= 64:
    t_tri2, C2 = timed(lambda: MSV.fcn2vintercept(dA, dU))
    print("tri2v: %d pairs x %d points           %8.2f ms" % (F * (F - 1) // 2, NT, t_tri2))

# stage 4: bundle adjustment, 10 LM iterations
rng = np.random.default_rng(3)
z = np.concatenate((P[0].T.ravel(), P[1].T.ravel())).astype(np.float64)
xinit = np.concatenate((pw + rng.normal(0, 0.05, pw.shape), (cw + rng.normal(0, 0.01, cw.shape))[1:], np.zeros((F - 1, 3)))).ravel()
ba = NLS.BundleAdjuster(K, z, xinit, NT, F - 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
hist = []
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(10):
    ev[0].record(); ba.accumulate(); ev[1].record(); ba.solve(); ev[2].record()
    torch.cuda.synchronize()
    hist.append((ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), float(np.sqrt(ba.cost.item() / z.size)), ba.rms_delta.item()))
t_ba = (time.perf_counter() - t0) * 1e3
print("BA   : nt=%d nc=%d nx=%d, 10 iterations  %8.2f ms  (accumulate %.2f ms, solve %.2f ms per iteration)" % (
    NT, F - 1, xinit.size, t_ba, np.mean([h[0] for h in hist[1:]]), np.mean([h[1] for h in hist[1:]])))
print("       rms residual per iteration: " + " ".join("%.4f" % h[2] for h in hist))
xs = ba.x.cpu().numpy()
print("       point error vs truth: init %.3e -> final %.3e m (gauge-free: relative to camera 0)" % (
    np.abs(xinit[:3 * NT].reshape(-1, 3) - pw).max(), np.abs(xs[:3 * NT].reshape(-1, 3) - pw).max()))
total = t_klt + t_nls + t_tri + t_ba
print("TOTAL: %.1f ms for %d frames -> %.0f frames/s (KLT %.0f%%, NLS %.0f%%, tri %.0f%%, BA %.0f%%)" % (
    total, F, F / total * 1e3, 100 * t_klt / total, 100 * t_nls / total, 100 * t_tri / total, 100 * t_ba / total))

if args.cpu:
    import cv2

    from oracle import sfm_oracle as S

    a, b = uniq[0], uniq[1]
    ph = pts.cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(3):
        q2, s1, _ = cv2.calcOpticalFlowPyrLK(a, b, ph, None, **LK)
        cv2.calcOpticalFlowPyrLK(b, a, q2, None, **LK)
    c_klt = (time.perf_counter() - t0) / 3 * 1e3
    t0 = time.perf_counter()
    S.solve_translation(K, P[0:2, :, 1].T.astype(float), pw, np.zeros(3))
    c_nls = (time.perf_counter() - t0) * 1e3
    nt_s, nf_s = 512, 20
    Ps, cws = synth.scene_observations(pw[:nt_s], nf_s, step=0.02, noise=0.1, seed=11)
    t0 = time.perf_counter()
    S.bundle_sparse(K, Ps, pw[:nt_s] + 0.01, cws, max_iter=2)
    c_ba_small = (time.perf_counter() - t0) / 2 * 1e3
    print("CPU  : cv2 KLT fwd+bwd %.1f ms/pair (%d threads); oracle fcnNLS_t %.1f ms/frame; oracle sparse BA nt=%d nf=%d %.0f ms/iteration"
          % (c_klt, cv2.getNumThreads(), c_nls, nt_s, nf_s, c_ba_small))
    print("       (the reference's own dense BA needs a 277 GB Jacobian at nt=4096, nf=300: infeasible as written)")
