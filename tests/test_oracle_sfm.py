"""Pins the numpy solver oracle (oracle/sfm_oracle.py) to the reference's outputs (tests/golden).
Tolerances: float32 outputs of the LM solvers 1e-5 relative; float64 triangulation
1e-9; bundle adjustment 1e-6 after the same number of iterations.  CPU only."""
import numpy as np

from oracle import sfm_oracle as S
from util import golden


def test_transforms():
    g = golden("transforms")
    for r, c, b in zip(g["rpy"], g["dcm"], g["back"]):
        assert np.allclose(S.euler_dcm(r), c, rtol=0, atol=1e-15)
        assert np.allclose(S.dcm_euler(c), b, rtol=0, atol=1e-15)


def test_translation_solver():
    g = golden("nls_t")
    t, ok = S.solve_translation(g["K"], g["p"].astype(float), g["pw"], g["x0"])
    assert ok and t.dtype == np.float32
    assert np.allclose(t, g["t"], rtol=1e-5, atol=1e-7)


def test_pose_solver():
    g = golden("nls_rt64")
    R, t, ok = S.solve_pose(g["K"], g["p"], g["pw"], g["x0"])
    assert np.allclose(R, g["R"], rtol=1e-5, atol=1e-6) and np.allclose(t, g["t"], rtol=1e-5, atol=1e-6)
    g = golden("nls_rt")
    x0 = np.concatenate((S.dcm_euler(np.eye(3)), [0, 0, 1.0]))
    R, t, ok = S.solve_pose(g["K"], g["q"].astype(float), g["plate"], x0)
    assert np.allclose(R, g["R"], rtol=1e-5, atol=1e-6) and np.allclose(t, g["t"], rtol=1e-5, atol=1e-6)


def test_triangulation():
    g = golden("triangulate")
    assert np.allclose(S.triangulate_pairs(g["A"], g["U"]), g["c2v"], rtol=1e-10, atol=1e-10)
    assert np.allclose(S.triangulate_rays(g["A"], g["U"]), g["cnv"], rtol=1e-9, atol=1e-9)


def test_msv1():
    g = golden("msv_t")
    x, b0, capped = S.solve_last_translation(g["K"], g["P"], g["B"], g["vg"], int(g["ii"]))
    assert np.allclose(x, g["x1"], rtol=1e-5, atol=1e-6)
    assert np.allclose(b0, g["b0"], rtol=1e-6, atol=1e-7)


def test_bundle_dense_and_sparse_match_reference():
    # ba_256x10 / ba_512x20: the SURVEY 8(d) anchors, produced by the reference's DENSE fcnNLS_batch (270 MB Jacobian)
    for name in ("ba_small", "ba_medium", "ba_256x10", "ba_512x20"):
        g = golden(name)
        cw, pw, hist = S.bundle_sparse(g["K"], g["P"], g["pw0"], g["cw0"])
        assert np.allclose(cw, g["cw"], rtol=1e-6, atol=1e-7), name
        assert np.allclose(pw, g["pw"], rtol=1e-6, atol=1e-7), name
    g = golden("ba_small")
    cw, pw, hist_d = S.bundle_dense(g["K"], g["P"], g["pw0"], g["cw0"])
    assert np.allclose(cw, g["cw"], rtol=1e-7, atol=1e-8) and np.allclose(pw, g["pw"], rtol=1e-7, atol=1e-8)
    # the reference prints "i: ..s, f=.., x=.." per iteration: same iteration count
    n_ref = sum(1 for ln in str(g["stdout"]).splitlines() if ": " in ln and "f=" in ln and "done" not in ln)
    assert len(hist_d) == n_ref


def test_sparse_form_iteration_count_and_residual_match_reference_at_anchor_sizes():
    import re

    for name in ("ba_256x10", "ba_512x20"):
        g = golden(name)
        _, _, hist = S.bundle_sparse(g["K"], g["P"], g["pw0"], g["cw0"])
        lines = [ln for ln in str(g["stdout"]).splitlines() if re.match(r"^\d+: ", ln)]
        assert len(hist) == len(lines), name
        f_ref = [float(re.search(r"f=([^,]+),", ln).group(1)) for ln in lines]
        assert np.allclose([h[0] for h in hist], f_ref, rtol=2e-5), name     # printed with 6 significant digits
