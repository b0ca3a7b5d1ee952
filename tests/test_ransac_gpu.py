"""Parity of K10 (GPU RANSAC affine fit, csrc/ransac.cu) against cv2.estimateAffine2D's golden outputs: inlier masks
bit-exact, T within 1e-9 and identical after the float32 cast of utils/KLT.py:58."""
import numpy as np
import pytest

from test_oracle_ransac import check_against_cv2, ransac_cases

pytestmark = pytest.mark.gpu


def test_gpu_ransac_matches_cv2_golden():
    from velocity_b200 import ransac

    for k, fr, to, T_ref, inl_ref in ransac_cases():
        T, inl = ransac.estimateAffine2D(fr, to)
        assert inl.dtype == np.uint8 and inl.shape == inl_ref.shape
        check_against_cv2(k, T, inl, T_ref, inl_ref)
    assert ransac.estimateAffine2D(np.zeros((2, 2), np.float32), np.zeros((2, 2), np.float32)) == (None, None)
    with pytest.raises(NotImplementedError):
        ransac.estimateAffine2D(fr, to, method=4)


def test_gpu_ransac_against_cv2_itself_random_trials():
    cv2 = pytest.importorskip("cv2")
    from velocity_b200 import ransac

    rng = np.random.default_rng(9)
    for trial in range(60):
        n = int(rng.integers(4, 3000))
        fr = rng.uniform(0, 1900, (n, 2)).astype(np.float32)
        A = np.array([[1 + rng.normal() * 0.02, rng.normal() * 0.02], [rng.normal() * 0.02, 1 + rng.normal() * 0.02]])
        to = (fr @ A.T + rng.normal(size=2) * 20 + rng.normal(size=(n, 2)) * rng.choice([0.1, 0.5, 1.5])).astype(np.float32)
        nout = int(n * rng.choice([0, 0.1, 0.3, 0.6]))
        if nout:
            to[rng.choice(n, nout, replace=False)] += rng.uniform(-80, 80, (nout, 2)).astype(np.float32)
        T_ref, inl_ref = cv2.estimateAffine2D(fr, to, method=cv2.RANSAC)
        T, inl = ransac.estimateAffine2D(fr, to)
        check_against_cv2(trial, T, inl, T_ref, inl_ref)
