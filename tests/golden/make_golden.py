"""Generate the committed golden vectors by RUNNING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py            # needs /root/reference and cv2 4.13.0

Every .npz written next to this file holds seeded inputs plus the outputs of the reference's own
functions (utils/KLT.py, utils/NLS.py, utils/MSV.py, utils/transforms.py, vidExample.py through
oracle/ref_shim.py, which repairs the three one-line defects of SURVEY.md 0.2 at load time).
The reference has no tests or golden vectors of its own; these are the pins.
/root/reference does not exist on the GPU box, so the fixtures -- not this script -- travel.
"""
import contextlib
import io
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cv2  # noqa: E402

from oracle import ref_shim  # noqa: E402
from velocity_b200 import synth  # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
from util import ba_c3_inputs  # noqa: E402

EPS, COUNT = cv2.TERM_CRITERIA_EPS, cv2.TERM_CRITERIA_COUNT
LK_COARSE = dict(winSize=(15, 15), maxLevel=4, criteria=(EPS | COUNT, 10, 0.1))  # utils/KLT.py:106
LK_FINE = dict(winSize=(51, 51), maxLevel=0, criteria=(EPS | COUNT, 30, 0.001))  # utils/KLT.py:107
LK_C2 = dict(winSize=(15, 15), maxLevel=2, criteria=(EPS | COUNT, 10, 0.1))


def save(name, **kw):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **kw)
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def affine_pair(h, w, seed, M):
    im0 = synth.texture(h, w, seed)
    im1 = cv2.warpAffine(im0, np.asarray(M, np.float32), (w, h), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    return im0, im1


def lk_params_arrays(lk):
    return dict(win=np.array(lk["winSize"]), max_level=np.array(lk["maxLevel"]), max_count=np.array(lk["criteria"][1]),
                eps=np.array(lk["criteria"][2], np.float64))


def gen_lk(ref):
    rng = np.random.default_rng(100)
    cases = [
        ("lk_coarse", 240, 320, LK_COARSE, None, 300, [[1.01, 0.015, 2.2], [-0.012, 0.995, -1.4]]),
        ("lk_coarse_fb", 240, 320, LK_COARSE, 1.0, 300, [[1.0, 0.0, 3.6], [0.0, 1.0, -2.7]]),
        ("lk_c2_fb", 360, 480, LK_C2, 1.0, 500, [[1.004, 0.003, 1.7], [-0.002, 1.003, 0.9]]),
        ("lk_fine_fb", 200, 260, LK_FINE, 0.3, 160, [[1.0, 0.002, 0.6], [-0.002, 1.0, -0.4]]),
        ("lk_odd_21", 241, 323, dict(winSize=(21, 21), maxLevel=3, criteria=(EPS | COUNT, 30, 0.01)), None, 300,
         [[0.99, -0.01, 4.0], [0.01, 1.01, 1.0]]),
        ("lk_rect_win", 150, 210, dict(winSize=(9, 13), maxLevel=3, criteria=(EPS | COUNT, 20, 0.03)), 0.5, 200,
         [[1.0, 0.0, 1.3], [0.0, 1.0, 0.8]]),
    ]
    for name, h, w, lk, fbt, n, M in cases:
        im0, im1 = affine_pair(h, w, zlib.crc32(name.encode()) % 1000, M)
        # interior points plus a band of border / out-of-frame points (status=0 and REFLECT_101 paths)
        p = np.stack([rng.uniform(-12, w + 12, n), rng.uniform(-12, h + 12, n)], 1).astype(np.float32)
        p[: n // 2] = np.stack([rng.uniform(20, w - 20, n // 2), rng.uniform(20, h - 20, n // 2)], 1)
        p[-4:] = [[0, 0], [w - 1, h - 1], [-7.5, 10], [w + 6.25, h - 3]]
        p2, v, err = ref.KLT.cv2calcOpticalFlowPyrLK(im0, im1, p, None, fbt=fbt, **lk)
        save(name, im0=im0, im1=im1, p=p, p2=p2, v=v, err=err, fbt=np.array(-1.0 if fbt is None else fbt),
             **lk_params_arrays(lk))

    # numpy ROI views (non-contiguous) exactly as KLTregional passes them (utils/KLT.py:62,68)
    im0, im1 = affine_pair(300, 400, 77, [[1.0, 0.0, 2.0], [0.0, 1.0, 1.0]])
    roi0, roi1 = im0[21:260, 33:377], im1[23:262, 31:375]
    p = np.stack([rng.uniform(5, 339, 150), rng.uniform(5, 234, 150)], 1).astype(np.float32)
    p2, v, err = ref.KLT.cv2calcOpticalFlowPyrLK(roi0, roi1, p, None, fbt=1.0, **LK_COARSE)
    save("lk_roi_views", im0=im0, im1=im1, roi0=np.array([21, 260, 33, 377]), roi1=np.array([23, 262, 31, 375]), p=p, p2=p2,
         v=v, err=err, fbt=np.array(1.0), **lk_params_arrays(LK_COARSE))


def gen_primitives():
    for name, h, w in [("prim_even", 128, 192), ("prim_odd", 131, 197), ("prim_tiny", 37, 23)]:
        im = synth.texture(h, w, h + w)
        T = np.array([[1.013, 0.021], [-0.017, 0.992], [3.37, -2.81]], np.float32)  # 3x2 row-vector affine
        x0, x1, y0, y1 = 3, w - 2, 1, h - 5
        x, y = np.meshgrid(np.arange(x0, x1, dtype=np.float32), np.arange(y0, y1, dtype=np.float32), copy=False)
        mx = x * T[0, 0] + y * T[1, 0] + T[2, 0]  # utils/KLT.py:71-72
        my = x * T[0, 1] + y * T[1, 1] + T[2, 1]
        save(name, im=im, pyrdown=cv2.pyrDown(im), scharr_x=cv2.Scharr(im, cv2.CV_16S, 1, 0),
             scharr_y=cv2.Scharr(im, cv2.CV_16S, 0, 1),
             quarter=cv2.resize(im, (0, 0), fx=0.25, fy=0.25, interpolation=cv2.INTER_NEAREST), T=T,
             roi=np.array([x0, x1, y0, y1]), remap=cv2.remap(im, mx, my, cv2.INTER_LINEAR))


def gen_regional(ref):
    rng = np.random.default_rng(5)
    h, w = 360, 480
    M = [[1.006, 0.004, 5.3], [-0.003, 1.005, -3.6]]
    im0, im1 = affine_pair(h, w, 31, M)
    p0 = np.stack([rng.uniform(90, 380, 180), rng.uniform(70, 290, 180)], 1).astype(np.float32)
    # translate-only pass (utils/KLT.py:121-124)
    T = np.eye(3, 2)
    T[2] = [5.3 + 1.4, -3.6 + 0.9]
    p, v = ref.KLT.KLTregional(im0, im1, p0, T, LK_COARSE, fbt=1, translateFlag=True)
    save("regional_translate", im0=im0, im1=im1, p0=p0, T=T, p=p, v=v, fbt=np.array(1.0), **lk_params_arrays(LK_COARSE))
    # affine fine pass (utils/KLT.py:133)
    T23 = np.array(M, np.float64) + rng.normal(0, 1e-4, (2, 3))
    p, v = ref.KLT.KLTregional(im0, im1, p0, T23.T, LK_FINE, fbt=0.3)
    save("regional_affine", im0=im0, im1=im1, p0=p0, T=T23.T, p=p, v=v, fbt=np.array(0.3), **lk_params_arrays(LK_FINE))


def gen_kltmain(ref):
    h, w = 480, 640
    M = [[1.012, 0.003, 9.4], [-0.002, 1.011, -6.2]]
    im0, im1 = affine_pair(h, w, 404, M)
    roi = im0[120:360, 160:480]
    p0 = cv2.goodFeaturesToTrack(roi, 220, 0.01, 0, blockSize=5, useHarrisDetector=True).reshape(-1, 2) + np.float32([160, 120])
    p0 = p0.astype(np.float32)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        p, v, im_small = ref.KLT.KLTmain(im1, im0, None, p0)
    save("kltmain_pair", im0=im0, im1=im1, p0=p0, p=p, v=v, im_small=im_small)


def gen_nls(ref):
    rng = np.random.default_rng(9)
    K = synth.K_1080P.copy()
    # fcnNLS_t / estimateWorldCameraPose(findR=False)  (utils/NLS.py:102-129, :9-33)
    pw = synth.scene_points(151, seed=3)
    t_true = np.array([0.11, -0.07, -0.37])
    uv = (pw + t_true) @ K
    p = (uv[:, :2] / uv[:, 2:3] + rng.normal(0, 0.3, (151, 2))).astype(np.float32)
    x0 = np.array([0.0, 0.0, 1.0])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        t = ref.NLS.fcnNLS_t(K, p.astype(float), pw, x0.copy())
        tt, RR, res, pproj = ref.NLS.estimateWorldCameraPose(K, p, pw, t=x0.copy(), R=np.eye(3), findR=False)
    save("nls_t", K=K, p=p, pw=pw, x0=x0, t=t, est_t=tt, est_R=RR, est_res=np.array(res), est_pproj=pproj,
         stdout=np.array(buf.getvalue()))
    # fcnNLS_Rt on the plate corners (vidExample.py:118)
    plate = ref.common.worldPointsLicensePlate("Chile")
    rpy = np.array([0.05, -0.22, 0.03])
    R_true = ref.transforms.rpy2dcm(rpy)
    t_true = np.array([-0.4, 0.9, 7.5])
    uv = (plate.astype(float) @ R_true + t_true) @ K
    q = (uv[:, :2] / uv[:, 2:3]).astype(np.float32) + rng.normal(0, 0.05, (4, 2)).astype(np.float32)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        t6, R6, res6, pproj6 = ref.NLS.estimateWorldCameraPose(K, q, plate, findR=True)
    save("nls_rt", K=K, q=q, plate=plate, t=t6, R=R6, res=np.array(res6), pproj=pproj6, stdout=np.array(buf.getvalue()))
    # larger 6-dof problem
    pw2 = synth.scene_points(64, seed=12) - np.array([0, 0, 10.0])
    uv = (pw2 @ R_true + t_true) @ K
    p2 = uv[:, :2] / uv[:, 2:3] + rng.normal(0, 0.2, (64, 2))
    x06 = np.concatenate([np.zeros(3), [0, 0, 6.0]])
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        R6b, t6b = ref.NLS.fcnNLS_Rt(K, p2, pw2, x06.copy())
    save("nls_rt64", K=K, p=p2, pw=pw2, x0=x06, R=R6b, t=t6b, stdout=np.array(buf.getvalue()))
    # projection helpers
    a = synth.scene_points(32, seed=5)
    Rr = ref.transforms.rpy2dcm([0.1, -0.05, 0.2])
    save("projection", K=K, a=a, fzK=ref.NLS.fzK(a, K), fzC=ref.NLS.fzC(a, K, Rr, np.array([[0.1, 0.2, 0.3]])),
         R=Rr, world2image=ref.common.world2image(K, Rr, np.array([0.1, 0.2, 0.3]), a),
         pixel2uvec=ref.common.pixel2uvec(K, ref.NLS.fzK(a, K)))


def gen_transforms(ref):
    rng = np.random.default_rng(21)
    rpy = rng.uniform(-1.2, 1.2, (16, 3))
    dcm = np.stack([ref.transforms.rpy2dcm(r) for r in rpy])
    back = np.stack([ref.transforms.dcm2rpy(c) for c in dcm])
    X = rng.normal(0, 3, (40, 3))
    t = np.array([0.5, -1.5, 2.5])
    xf = np.stack([ref.transforms.transform(X, r, t) for r in rpy[:4]])
    quat = np.stack([ref.transforms.dcm2quat(c) for c in dcm])          # the 3-element pseudo-quaternion helpers (:60-73)
    quat_dcm = np.stack([ref.transforms.quat2dcm(q) for q in quat])
    save("transforms", rpy=rpy, dcm=dcm, back=back, X=X, t=t, xf=xf, quat=quat, quat_dcm=quat_dcm)


def gen_msv(ref):
    K = synth.K_1080P.copy()
    pw = synth.scene_points(40, seed=8)
    nf = 6
    P, cw = synth.scene_observations(pw, nf, noise=0.05, seed=4)
    vg = np.ones(40, bool)
    vg[[3, 17]] = False
    U = np.zeros((3, nf, int(vg.sum())))
    for j in range(nf):
        U[:, j] = ref.common.pixel2uvec(K, P[0:2, vg, j].T).T
    A = -cw  # camera origins in the frame of camera 0 (p_cam = p_w + cw  =>  origin = -cw)
    save("triangulate", A=A, U=U, c2v=ref.MSV.fcn2vintercept(A, U), cnv=ref.MSV.fcnNvintercept(A, U), pw=pw[vg])
    # fcnMSV1_t as vidExample.py:158 drives it: B[:,0:3] holds camera translations.
    # fcnMSV2_t (utils/MSV.py:52-94) cannot be pinned: its zero Jacobian blocks are turned into
    # -zhat/dx by `JT = (JT - zhat) / dx` (:84), JTJ is numerically singular and np.linalg.inv raises
    # LinAlgError for any realistic input (verified here); it also only reshapes for i == 2 (:71).
    B = np.zeros((nf, 14), np.float32)
    t0 = np.array([0.3, -0.2, 9.0])
    for j in range(nf):
        B[j, 0:3] = t0 + cw[j] + (0 if j < nf - 1 else np.array([0.02, -0.01, 0.03]))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        x1, b0 = ref.MSV.fcnMSV1_t(K, P, B, vg, nf - 1)
    save("msv_t", K=K, P=P, B=B, vg=vg, ii=np.array(nf - 1), x1=x1, b0=b0, stdout=np.array(buf.getvalue()))


def gen_ba(ref):
    K = synth.K_1080P.copy()
    rng = np.random.default_rng(33)
    for name, nt, nf in [("ba_small", 24, 5), ("ba_medium", 96, 8)]:
        pw = synth.scene_points(nt, seed=nt)
        P, cw = synth.scene_observations(pw, nf, noise=0.1, seed=nf)
        if name == "ba_small":
            P[:, 5, 2] = np.nan  # one short track: must be filtered out (utils/NLS.py:190)
        pw0 = pw + rng.normal(0, 0.05, pw.shape)
        cw0 = cw + rng.normal(0, 0.02, cw.shape)
        cw0[0] = 0
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            cw1, pw1 = ref.NLS.fcnNLS_batch(K, P.copy(), pw0.copy(), cw0.copy())
        out = dict(K=K, P=P, pw0=pw0, cw0=cw0, cw=cw1, pw=pw1, stdout=np.array(buf.getvalue()))
        if name == "ba_small":
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                cw2, pw2 = ref.NLS.fcnNLS_batch2(K, P.copy(), pw0.copy(), cw0.copy())
            out.update(cw_b2=cw2, pw_b2=pw2, stdout_b2=np.array(buf.getvalue()))
        save(name, **out)


def gen_ba_large(ref):
    """SURVEY 8(d) anchors: the reference's own DENSE fcnNLS_batch at nt=256/nf=10 and nt=512/nf=20 (1.6 s per
    iteration, 270 MB Jacobian) -- the multi-chunk path of K7 (>= 16 cameras) and a K8 system of 114 unknown cameras."""
    K = synth.K_1080P.copy()
    rng = np.random.default_rng(77)
    for name, nt, nf in [("ba_256x10", 256, 10), ("ba_512x20", 512, 20)]:
        pw = synth.scene_points(nt, seed=nt)
        P, cw = synth.scene_observations(pw, nf, noise=0.1, seed=nf)
        pw0 = pw + rng.normal(0, 0.05, pw.shape)
        cw0 = cw + rng.normal(0, 0.02, cw.shape)
        cw0[0] = 0
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            cw1, pw1 = ref.NLS.fcnNLS_batch(K, P.copy(), pw0.copy(), cw0.copy())
        save(name, K=K, P=P, pw0=pw0, cw0=cw0, cw=cw1, pw=pw1, stdout=np.array(buf.getvalue()))


def gen_ba_c3(ref):
    """The C3-size problem cannot run in the reference (277 GB dense Jacobian): the fixture holds the result of the
    ORACLE's block-sparse form (oracle/sfm_oracle.bundle_sparse, the same iteration solved through the Schur complement;
    checked against the reference's dense results on ba_256x10 / ba_512x20 in tests/test_oracle_sfm.py) after the
    reference's 10 iterations.  Inputs are regenerated from seeds (ba_c3_inputs)."""
    from oracle import sfm_oracle as S

    K, P, pw0, cw0 = ba_c3_inputs()
    cw1, pw1, hist = S.bundle_sparse(K, P, pw0, cw0, max_iter=10)
    save("ba_c3_sparse", cw=cw1, pw=pw1, hist=np.array(hist), pw0_crc=np.array(zlib.crc32(pw0.tobytes())), P_crc=np.array(zlib.crc32(P.tobytes())))


def gen_match():
    rng = np.random.default_rng(55)
    nq, nt = 300, 400
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    for i in range(0, nq, 2):  # planted near-duplicates (<= 20 flipped bits) and exact ties
        src = t[rng.integers(0, nt)].copy()
        bits = rng.choice(256, rng.integers(0, 21), replace=False)
        for b in bits:
            src[b // 8] ^= 1 << (b % 8)
        q[i] = src
    t[7] = t[3]  # duplicate train rows: lowest trainIdx must win
    q[1] = t[3]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, k=2)
    idx = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    dist = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    qf = rng.normal(0, 1, (nq, 64)).astype(np.float32)
    tf = rng.normal(0, 1, (nt, 64)).astype(np.float32)
    qf[::3] = tf[rng.integers(0, nt, len(qf[::3]))] + rng.normal(0, 0.05, (len(qf[::3]), 64)).astype(np.float32)
    m = cv2.BFMatcher().knnMatch(qf, tf, k=2)  # utils/KLT.py:16,25 (default NORM_L2)
    idxf = np.array([[a.trainIdx, b.trainIdx] for a, b in m], np.int32)
    distf = np.array([[a.distance, b.distance] for a, b in m], np.float32)
    save("match_knn2", q=q, t=t, idx=idx, dist=dist, qf=qf, tf=tf, idxf=idxf, distf=distf)


def gen_ingest():
    """cv2.cvtColor(BGR2GRAY) (vidExample.py:91) on random colours plus the extreme / rounding-edge triples."""
    rng = np.random.default_rng(91)
    bgr = rng.integers(0, 256, (61, 83, 3), dtype=np.uint8)
    bgr[0, :8] = [[0, 0, 0], [255, 255, 255], [255, 0, 0], [0, 255, 0], [0, 0, 255], [1, 1, 1], [254, 255, 254], [128, 127, 129]]
    save("ingest_bgr", bgr=bgr, gray=cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))


def gen_gftt():
    """cv2.goodFeaturesToTrack(roi, n, q, 0, blockSize=5, useHarrisDetector=True) (vidExample.py:110) and the Harris
    response behind it, on widths that are / are not multiples of 32 and 16 (cv2's SIMD tails)."""
    out = {}
    for tag, (h, w, seed) in {"a": (120, 160, 21), "b": (97, 203, 22), "c": (150, 77, 23)}.items():
        im = synth.texture(h, w, seed)
        out["im_" + tag] = im
        out["resp_" + tag] = cv2.cornerHarris(im, 5, 3, 0.04)
        for n, q in ((1000, 0.01), (64, 0.05), (4096, 0.001)):
            c = cv2.goodFeaturesToTrack(im, n, q, 0, blockSize=5, useHarrisDetector=True)
            out["xy_%s_%d" % (tag, n)] = np.zeros((0, 1, 2), np.float32) if c is None else c
    save("gftt", **out)


def gen_subpix():
    """cv2.cornerSubPix(im, p, (5,5), (-1,-1), criteria) (vidExample.py:113-115) on detector corners and on random points,
    many of them within the 7-px rim where the 13x13 sampling window hangs over the frame border."""
    out = {}
    crits = {"ref": (EPS | COUNT, 100, 0.001), "loose": (EPS | COUNT, 20, 0.03), "count": (COUNT, 7, 0.0)}
    for tag, (h, w, seed) in {"a": (120, 160, 41), "b": (97, 203, 42), "c": (31, 40, 43)}.items():
        im = synth.texture(h, w, seed)
        rng = np.random.default_rng(seed)
        p = np.stack([rng.uniform(0, w - 1, 400), rng.uniform(0, h - 1, 400)], 1).astype(np.float32)
        rim = np.stack([rng.choice([rng.uniform(0, 7), rng.uniform(w - 8, w - 1)], 1) for _ in range(100)] , 0)
        rim = np.concatenate([rim, rng.uniform(0, h - 1, (100, 1))], 1).astype(np.float32)
        g = cv2.goodFeaturesToTrack(im, 200, 0.01, 0, blockSize=5, useHarrisDetector=True).reshape(-1, 2)
        p = np.concatenate([p, rim, g]).astype(np.float32)
        out["im_" + tag], out["p_" + tag] = im, p
        for cn, c in crits.items():
            out["q_%s_%s" % (tag, cn)] = cv2.cornerSubPix(im, p.copy().reshape(-1, 1, 2), (5, 5), (-1, -1), c).reshape(-1, 2)
    save("subpix", **out)


def gen_ransac():
    """cv2.estimateAffine2D(from, to, method=cv2.RANSAC) (utils/KLT.py:116,127) on seeded correspondence sets: clean,
    10-60 % gross outliers, few points, exactly 3 points, all-collinear (no model)."""
    rng = np.random.default_rng(116)
    out, k = {}, 0
    for n, noise, frac in [(600, 0.2, 0.0), (350, 0.5, 0.1), (900, 1.5, 0.3), (120, 0.5, 0.6), (4096, 0.3, 0.25), (12, 0.1, 0.2), (4, 0.0, 0.0),
                           (3, 0.0, 0.0)]:
        fr = rng.uniform(0, 1900, (n, 2)).astype(np.float32)
        A = np.array([[1 + rng.normal() * 0.02, rng.normal() * 0.02], [rng.normal() * 0.02, 1 + rng.normal() * 0.02]])
        to = (fr @ A.T + rng.normal(size=2) * 20 + rng.normal(size=(n, 2)) * noise).astype(np.float32)
        nout = int(n * frac)
        if nout:
            to[rng.choice(n, nout, replace=False)] += rng.uniform(-80, 80, (nout, 2)).astype(np.float32)
        T, inl = cv2.estimateAffine2D(fr, to, method=cv2.RANSAC)
        out["from_%d" % k], out["to_%d" % k], out["T_%d" % k], out["inl_%d" % k] = fr, to, T, inl
        k += 1
    line = np.stack([np.arange(9, dtype=np.float32) * 7, np.arange(9, dtype=np.float32) * 3 + 1], 1)
    T, inl = cv2.estimateAffine2D(line, line + 2, method=cv2.RANSAC)
    assert T is None
    out["from_%d" % k], out["to_%d" % k], out["T_%d" % k], out["inl_%d" % k] = line, line + 2, np.zeros((0, 3)), inl
    out["ncases"] = np.array(k + 1)
    save("ransac", **out)


def gen_e2e():
    """vidExample.py end to end on the two clips that have plate fixtures."""
    mod, ns = ref_shim.load_vid_example()
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    try:
        out = {}
        for clip, start in [("IMG_4134.MOV", 19), ("IMG_4119.MOV", 41)]:
            src = open(os.path.join(ref_shim.REF_ROOT, "vidExample.py")).read()
            assert 'filename, startframe = f"{patha}IMG_4134.MOV", 19' in src
            m2, _ = ref_shim.load_vid_example()
            code = src.replace('filename, startframe = f"{patha}IMG_4134.MOV", 19',
                               'filename, startframe = f"{patha}%s", %d' % (clip, start))
            code = code.replace("S[i, :] = (i, proc_dt[i],", "S[i, :] = (i, proc_dt[i, 0],")
            marker = "        im_gaussian = cv2.GaussianBlur(im, (3, 3), 0)"
            code = code.replace(marker, "        im0 = im\n" + marker).replace("            del im0\n", "")
            g = {"__name__": "vid_shim"}
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                exec(compile(code, "<vidExample %s>" % clip, "exec"), g)
                g["vidExamplefcn"]()
            text = buf.getvalue()
            rows = [ln.split() for ln in text.splitlines() if ln.strip() and ln.split()[0].isdigit() and len(ln.split()) == 9]
            tab = np.array([[float(v) for v in r] for r in rows])
            summary = [ln for ln in text.splitlines() if ln.startswith("Speed") or ln.startswith("Res")]
            print(clip, summary)
            out[clip.split(".")[0] + "_table"] = tab
            out[clip.split(".")[0] + "_summary"] = np.array("\n".join(summary))
        save("e2e_vidExample", **out)
    finally:
        os.chdir(cwd)


def main():
    ref = ref_shim.load()
    print("cv2", cv2.__version__, "numpy", np.__version__)
    only = sys.argv[1:]
    if only:    # e.g. `make_golden.py ba_large ba_c3`: regenerate selected fixtures only
        import inspect

        for name in only:
            fn = globals()["gen_" + name]
            fn(ref) if inspect.signature(fn).parameters else fn()
        return
    gen_ba_large(ref)
    gen_ba_c3(ref)
    gen_primitives()
    gen_lk(ref)
    gen_regional(ref)
    gen_kltmain(ref)
    gen_nls(ref)
    gen_transforms(ref)
    gen_msv(ref)
    gen_ba(ref)
    gen_match()
    gen_ingest()
    gen_gftt()
    gen_subpix()
    gen_ransac()
    gen_e2e()


if __name__ == "__main__":
    main()
