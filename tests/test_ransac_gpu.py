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


def test_masked_device_fit_equals_the_compacted_call():
    """vel_estimate_affine2d_ransac_masked (the tracker's `T, inl = estimateAffine2D(p0[v], p[v]); v[v] = inl` as one launch on device
    data) against the plain call on the host-compacted rows: same mask bit for bit, same T bit for bit (same kernel body), the mapped
    `to` points equal numpy's float32 arithmetic; fewer than three kept rows = not found."""
    import torch

    from velocity_b200 import ransac

    rng = np.random.default_rng(21)
    for trial, (scale, off) in enumerate([(1.0, (0.0, 0.0)), (4.0, (0.0, 0.0)), (1.0, (137.0, -52.0)), (1.0, (0.0, 0.0))]):
        n = int(rng.integers(50, 4000)) if trial < 3 else 8192
        fr = rng.uniform(0, 1900, (n, 2)).astype(np.float32)
        to_full = (fr * np.float32(1.01) + np.float32(3.0) + rng.normal(size=(n, 2)).astype(np.float32) * np.float32(0.4)).astype(np.float32)
        to_full[rng.choice(n, n // 5, replace=False)] += rng.uniform(-60, 60, (n // 5, 2)).astype(np.float32)
        to_raw = ((to_full - np.float32(off)) / np.float32(scale)).astype(np.float32)
        v = rng.random(n) < 0.7
        mask_out, to_mapped, tail = ransac.estimateAffine2D_masked_device(torch.from_numpy(fr).cuda(), torch.from_numpy(to_raw).cuda(),
                                                                          torch.from_numpy(v.astype(np.uint8)).cuda(), to_scale=scale, to_off=off)
        to_np = (to_raw * np.float32(scale) + np.float32(off)).astype(np.float32)
        assert np.array_equal(to_mapped.cpu().numpy(), to_np)
        T_ref, inl_ref = ransac.estimateAffine2D(fr[v], to_np[v])
        want = v.copy()
        want[v] = inl_ref.ravel().astype(bool)
        t = tail.cpu().numpy()
        found, ninl, _, kept = (int(x) for x in t[6:8].view(np.int32))
        assert found == 1 and kept == int(v.sum()) and ninl == int(want.sum())
        assert np.array_equal(mask_out.cpu().numpy() != 0, want)
        assert np.array_equal(t[:6].reshape(2, 3), T_ref)
    # fewer than three rows kept: not found, mask cleared (cv2 returns None, None)
    v = np.zeros(100, np.uint8)
    v[[3, 50]] = 1
    fr = rng.uniform(0, 100, (100, 2)).astype(np.float32)
    mask_out, _, tail = ransac.estimateAffine2D_masked_device(torch.from_numpy(fr).cuda(), torch.from_numpy(fr).cuda(), torch.from_numpy(v).cuda())
    info = tail.cpu().numpy()[6:8].view(np.int32)
    assert info[0] == 0 and info[3] == 2 and not mask_out.cpu().numpy().any()
