"""CPU ORACLE (test infrastructure) for the connected sequence pipeline (BASELINE.json configs[2], "C3").

A restatement of the reference's frame loop, vidExample.py:75-166, reduced to the stages C3 names and built from the
pinned oracle pieces (oracle/klt_oracle.py, oracle/sfm_oracle.py), so it needs no fixture of its own beyond theirs:

    track_sequence       <- vidExample.py:134-135     p, v = KLT(im, im0, p); vg[vg] = v   with
                            utils/KLT.py:37-51        cv2calcOpticalFlowPyrLK (forward-backward gate) as the tracker
    pose_table           <- vidExample.py:139-146,151-153,164 + utils/NLS.py:9-33 (estimateWorldCameraPose, findR=False)
    triangulate_rays_vec <- utils/MSV.py:146-175      fcnNvintercept, vectorised (== sfm_oracle.triangulate_rays)
    run_sequence         <- the whole chain: tracking, per-frame pose + speed, triangulation over all frames,
                            fcnNLS_batch (sfm_oracle.bundle_sparse; the dense form needs a 277 GB Jacobian at C3 size)

Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs import this file.
The reference COMPACTS failed tracks (p = p[v]); this restatement keeps full-length arrays with NaN for dead tracks,
which is the same computation on the surviving ones (LK treats points independently).
"""
import numpy as np

from . import sfm_oracle as S


def track_sequence(frames, p0, fbt=1.0, lk_fn=None, nthreads=0, **lk):
    """tracks [n,npts,2] float32 (NaN once a track has failed), alive [n,npts] bool."""
    if lk_fn is None:
        from .klt_oracle import lk_forward_backward

        def lk_fn(a, b, p):
            return lk_forward_backward(a, b, p, fbt=fbt, nthreads=nthreads, **lk)[:2]
    n = len(frames)
    p0 = np.asarray(p0, np.float32)
    npts = p0.shape[0]
    tracks = np.full((n, npts, 2), np.nan, np.float32)
    alive = np.zeros((n, npts), bool)
    tracks[0], alive[0] = p0, True
    vg = np.ones(npts, bool)
    p = p0.copy()
    for i in range(1, n):
        p2, v = lk_fn(frames[i - 1], frames[i], p)       # vidExample.py:134 on the compacted point list
        vg[vg] = v                                       # :135
        p = p2[v]
        tracks[i, vg], alive[i] = p, vg
    return tracks, alive


def pose_table(K, tracks, alive, p3, frame_times, t0=(0.0, 0.0, 0.0), subset=None):
    """B [n,14], S [n,9] float32 and the reprojections proj [n,npts,2] float32 (NaN where a point is not in the fit)."""
    n, npts, _ = tracks.shape
    K = np.asarray(K, float)
    B = np.zeros((n, 14), np.float32)
    Sx = np.zeros((n, 9), np.float32)
    proj = np.full((n, npts, 2), np.nan, np.float32)
    B[:, 12] = np.asarray(frame_times, np.float32)
    B[0, 0:3] = np.asarray(t0, np.float32)
    sub = np.ones(npts, bool) if subset is None else np.asarray(subset, bool)
    r = np.float32(0)
    iters_capped = []
    for i in range(n):
        vp = alive[i] & sub
        if i == 0:
            proj[0, vp] = tracks[0, vp]                                      # p_ = p[vp]                      :125
            dt, dr, res = np.float32(np.nan), np.float32(0), np.float32(0)
        else:
            p = tracks[i, vp]
            t, ok = S.solve_translation(K, p.astype(float), p3[vp], np.array([0.0, 0.0, 1.0]))   # :139, default t (utils/NLS.py:9)
            if not ok:
                iters_capped.append(i)
            cam = np.concatenate([np.eye(3), t[None]]) @ K                    # world2image, utils/common.py:58-64 (float32 t)
            q = np.concatenate([p3[vp], np.ones((int(vp.sum()), 1))], 1) @ cam
            pp = q[:, :2] / q[:, 2:3]
            res = np.sqrt(np.mean((p - pp) ** 2))                             # rms(p - p_proj)                 utils/NLS.py:32
            proj[i, vp] = pp
            dt = B[i, 12] - B[i - 1, 12]                                      # :142
            d = t + B[0, 0:3] - B[i - 1, 0:3]                                 # :143 (float32 arrays)
            dr = np.sqrt((d * d).sum())
            r = r + dr
            B[i, 3:6] = t
            B[i, 0:3] = B[0, 0:3] + t
        Sx[i] = (i, 0.0, alive[i].sum(), res, dt, B[i, 12] - B[0, 12], dr, r, dr / dt * np.float32(3.6))   # :164
    return B, Sx, proj, iters_capped


def triangulate_rays_vec(A, U):
    """fcnNvintercept without the per-point Python loop: S1_i = sum_f (I - u u^T), S2_i = sum_f (I - u u^T) A_f."""
    M = np.eye(3)[:, :, None, None] - U[:, None] * U[None]      # [3,3,nf,nv]
    S1 = M.sum(2).transpose(2, 0, 1)
    S2 = np.einsum("abfv,fb->va", M, np.asarray(A, float))
    return np.einsum("vab,vb->va", np.linalg.inv(S1), S2)


def unit_rays_all(K, tracks):
    """U [3][n][nv] = pixel2uvec per frame (utils/MSV.py:13-15)."""
    return np.stack([S.unit_rays(np.asarray(K, float), tracks[j].astype(float)).T for j in range(tracks.shape[0])], 1)


def ba_inputs(tracks, alive):
    """P [5, nsel, n] of the full-length tracks (utils/NLS.py:190-191) and their indices."""
    idx = np.flatnonzero(alive[-1])
    n = tracks.shape[0]
    P = np.full((5, idx.size, n), np.nan, np.float32)
    P[0:2] = tracks[:, idx].transpose(2, 1, 0)
    P[4] = np.arange(n, dtype=np.float32)[None]
    return P, idx


def speeds_from_cameras(cw, B):
    """The S columns that follow from camera positions cw [n,3] (float32 arithmetic of vidExample.py:142-146)."""
    n = cw.shape[0]
    Bb = B.copy()
    Bb[:, 3:6] = cw.astype(np.float32)
    Bb[:, 0:3] = Bb[0:1, 0:3] + Bb[:, 3:6]
    sp = np.full(n, np.nan, np.float32)
    for i in range(1, n):
        d = Bb[i, 3:6] + Bb[0, 0:3] - Bb[i - 1, 0:3]
        sp[i] = np.sqrt((d * d).sum()) / (Bb[i, 12] - Bb[i - 1, 12]) * np.float32(3.6)
    return sp


def run_sequence(K, frames, p0, p3, frame_times, t0=(0.0, 0.0, 0.0), fbt=1.0, ba_iters=10, lk_fn=None, nthreads=0, **lk):
    tracks, alive = track_sequence(frames, p0, fbt=fbt, lk_fn=lk_fn, nthreads=nthreads, **lk)
    B, Sx, proj, capped = pose_table(K, tracks, alive, np.asarray(p3, float), frame_times, t0)
    P, idx = ba_inputs(tracks, alive)
    U = unit_rays_all(K, tracks[:, idx])
    A = (B[0, 0:3] - B[:, 0:3]).astype(float)
    C0 = triangulate_rays_vec(A, U)
    cw, pw, hist = S.bundle_sparse(K, P, C0, B[:, 3:6].astype(float), max_iter=ba_iters)
    return dict(tracks=tracks, alive=alive, B=B, S=Sx, proj=proj, idx=idx, C0=C0, cw=cw, pw=pw, hist=hist,
                speed_ba=speeds_from_cameras(cw, B), capped=capped)
