"""Pins oracle/ransac_oracle.py (restatement of cv2.estimateAffine2D(method=RANSAC), utils/KLT.py:116,127) to cv2's own
outputs in tests/golden/ransac.npz: inlier masks bit-exact, T within 1e-9 and identical after the float32 cast the
tracker applies (utils/KLT.py:58).  CPU only."""
import numpy as np
import pytest

from oracle import ransac_oracle as R
from util import golden

RANSAC_T_TOL = 1e-9


def ransac_cases():
    g = golden("ransac")
    for k in range(int(g["ncases"])):
        yield k, g["from_%d" % k], g["to_%d" % k], (g["T_%d" % k] if g["T_%d" % k].size else None), g["inl_%d" % k]


def check_against_cv2(k, T, inl, T_ref, inl_ref):
    assert np.array_equal(inl, inl_ref), k
    if T_ref is None:
        assert T is None, k
    else:
        assert T.shape == (2, 3) and T.dtype == np.float64
        assert np.abs(T - T_ref).max() <= RANSAC_T_TOL * max(1.0, np.abs(T_ref).max()), k
        assert np.array_equal(T.astype(np.float32), T_ref.astype(np.float32)), k


def test_oracle_matches_cv2_golden():
    for k, fr, to, T_ref, inl_ref in ransac_cases():
        T, inl = R.estimate_affine_2d(fr, to)
        check_against_cv2(k, T, inl, T_ref, inl_ref)


def test_oracle_against_cv2_itself_random_trials():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(8)
    for trial in range(40):
        n = int(rng.integers(4, 700))
        fr = rng.uniform(0, 1900, (n, 2)).astype(np.float32)
        A = np.array([[1 + rng.normal() * 0.02, rng.normal() * 0.02], [rng.normal() * 0.02, 1 + rng.normal() * 0.02]])
        to = (fr @ A.T + rng.normal(size=2) * 20 + rng.normal(size=(n, 2)) * rng.choice([0.1, 0.5, 1.5])).astype(np.float32)
        nout = int(n * rng.choice([0, 0.1, 0.3, 0.6]))
        if nout:
            to[rng.choice(n, nout, replace=False)] += rng.uniform(-80, 80, (nout, 2)).astype(np.float32)
        T_ref, inl_ref = cv2.estimateAffine2D(fr, to, method=cv2.RANSAC)
        T, inl = R.estimate_affine_2d(fr, to)
        check_against_cv2(trial, T, inl, T_ref, inl_ref)
