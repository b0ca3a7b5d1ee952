"""Pins the C oracle (oracle/velocity_oracle.c) to the reference's own outputs (tests/golden, made by
running /root/reference with cv2 4.13.0): primitives bit-exact, LK status masks bit-exact, points
within the stated tolerance.  CPU only."""
import numpy as np
import pytest

from oracle import cv_oracle as O
from oracle import klt_oracle as KO
from util import LK_CASES, LK_ERR_TOL, LK_POINT_TOL_PX, PRIM_CASES, fbt_of, golden, lk_images, lk_kwargs


@pytest.mark.parametrize("name", PRIM_CASES)
def test_primitives_bit_exact(name):
    g = golden(name)
    im = g["im"]
    assert np.array_equal(O.pyrDown(im), g["pyrdown"])
    sc = O.scharr(im)
    assert np.array_equal(sc[..., 0], g["scharr_x"]) and np.array_equal(sc[..., 1], g["scharr_y"])
    assert np.array_equal(O.decimate4(im), g["quarter"])
    x0, x1, y0, y1 = (int(v) for v in g["roi"])
    assert np.array_equal(O.remap_affine(im, g["T"], x0, x1, y0, y1), g["remap"])


@pytest.mark.parametrize("name", LK_CASES)
def test_lk_against_reference(name):
    g = golden(name)
    im0, im1 = lk_images(g)
    p2, v, err = KO.lk_forward_backward(im0, im1, g["p"], fbt=fbt_of(g), **lk_kwargs(g))
    assert np.array_equal(v, g["v"])                                   # masks: bit-exact
    assert np.abs(p2 - g["p2"])[g["v"]].max() <= LK_POINT_TOL_PX       # points: <= 2e-3 px
    assert np.abs(err - g["err"])[g["v"]].max() <= LK_ERR_TOL


def test_effective_max_level_rule():
    assert O.effective_max_level(320, 240, (15, 15), 4) == 3      # 15x20 at level 4 is not > 15 high
    assert O.effective_max_level(1920, 1080, (15, 15), 4) == 4
    assert O.effective_max_level(480, 270, (15, 15), 4) == 4
    assert O.effective_max_level(100, 100, (51, 51), 3) == 0


@pytest.mark.parametrize("name,translate", [("regional_translate", True), ("regional_affine", False)])
def test_regional_against_reference(name, translate):
    g = golden(name)
    p, v = KO.klt_regional(g["im0"], g["im1"], g["p0"], g["T"], lk_kwargs(g), fbt=float(g["fbt"]), translate=translate)
    assert np.array_equal(v, g["v"]) and p.dtype == g["p"].dtype
    assert np.abs(p - g["p"])[g["v"]].max() <= LK_POINT_TOL_PX


def test_kltmain_against_reference():
    g = golden("kltmain_pair")
    p, v, small = KO.klt_main(g["im1"], g["im0"], None, g["p0"])
    assert np.array_equal(v, g["v"]) and np.array_equal(small, g["im_small"])
    assert np.abs(p - g["p"]).max() <= LK_POINT_TOL_PX


def test_matcher_against_reference():
    g = golden("match_knn2")
    idx, dist = O.knn2_hamming(g["q"], g["t"])
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"].astype(np.int32))
    idx, dist = O.knn2_l2(g["qf"], g["tf"])
    assert np.array_equal(idx, g["idxf"]) and np.abs(dist - g["distf"]).max() < 1e-5


def test_empty_and_single_point():
    im = golden("prim_even")["im"]
    p2, st, err = O.calcOpticalFlowPyrLK(im, im, np.zeros((0, 2), np.float32), (15, 15), 2, (3, 10, 0.1))
    assert p2.shape == (0, 2) and st.shape == (0, 1)
    p2, st, err = O.calcOpticalFlowPyrLK(im, im, np.float32([[64.25, 40.75]]), (15, 15), 2, (3, 10, 0.1))
    assert st[0, 0] == 1 and np.abs(p2 - [[64.25, 40.75]]).max() < 1e-3 and err[0, 0] == 0


def test_bgr2gray_against_reference():
    g = golden("ingest_bgr")
    assert np.array_equal(O.bgr2gray(g["bgr"]), g["gray"])


def test_oracle_parameter_sweep_against_cv2_itself():
    """The oracle against the reference's arithmetic provider run on THIS host (opencv-python, when importable) over a
    seeded parameter sweep: forward-backward masks identical, points within the stated tolerance."""
    cv2 = pytest.importorskip("cv2")
    from util import lk_sweep_cases

    for k, im0, im1, pts, lk, fbt in lk_sweep_cases():
        o2, ov, _ = KO.lk_forward_backward(im0, im1, pts, fbt=fbt, **lk)
        q2, st, _ = cv2.calcOpticalFlowPyrLK(im0, im1, pts, None, **lk)
        qv = st.ravel().astype(bool)
        if fbt is not None:
            q1, st1, _ = cv2.calcOpticalFlowPyrLK(im1, im0, q2, None, **lk)
            d = pts - q1
            qv = qv & st1.ravel().astype(bool) & (np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < np.float32(fbt))
        assert np.array_equal(ov, qv), (k, lk, fbt)
        both = ov & qv
        if both.any():
            assert np.abs(o2 - q2)[both].max() <= LK_POINT_TOL_PX, (k, lk, fbt)
