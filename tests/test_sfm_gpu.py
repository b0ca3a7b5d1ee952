"""Parity of the solver / triangulation / bundle-adjustment / matcher kernels (K4-K8) against the
numpy oracle and the reference-generated golden vectors.  Tolerances follow SURVEY.md 8(c):
LM outputs (float32) 1e-5 relative; float64 triangulation 1e-9; BA 1e-6 (same iteration count);
match indices and Hamming distances bit-exact."""
import contextlib
import io

import numpy as np
import pytest

from util import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch

    assert torch.cuda.is_available()
    from velocity_b200 import _lib

    _lib.lib()
    return torch


def test_fcnNLS_t_and_estimateWorldCameraPose(cuda):
    from oracle import sfm_oracle as S
    from velocity_b200 import NLS

    g = golden("nls_t")
    t = NLS.fcnNLS_t(g["K"], g["p"].astype(float), g["pw"], g["x0"].copy())
    assert t.dtype == np.float32 and t.shape == (3,)
    assert np.allclose(t, g["t"], rtol=1e-5, atol=1e-7)
    to, _ = S.solve_translation(g["K"], g["p"].astype(float), g["pw"], g["x0"])
    assert np.allclose(t, to, rtol=1e-6, atol=1e-7)
    tt, RR, res, pproj = NLS.estimateWorldCameraPose(g["K"], g["p"], g["pw"], t=g["x0"].copy(), R=np.eye(3), findR=False)
    assert np.allclose(tt, g["est_t"], rtol=1e-5, atol=1e-7) and np.allclose(RR, g["est_R"])
    assert abs(res - float(g["est_res"])) < 1e-5 and np.allclose(pproj, g["est_pproj"], atol=2e-3)


def test_fcnNLS_Rt(cuda):
    from velocity_b200 import NLS

    g = golden("nls_rt64")
    R, t = NLS.fcnNLS_Rt(g["K"], g["p"], g["pw"], g["x0"].copy())
    assert R.dtype == np.float32 and t.dtype == np.float32
    assert np.allclose(R, g["R"], rtol=1e-5, atol=1e-6) and np.allclose(t, g["t"], rtol=1e-5, atol=1e-6)
    g = golden("nls_rt")
    t6, R6, res6, pproj6 = NLS.estimateWorldCameraPose(g["K"], g["q"], g["plate"], findR=True)
    assert np.allclose(R6, g["R"], rtol=1e-5, atol=1e-6) and np.allclose(t6, g["t"], rtol=1e-5, atol=1e-6)
    assert abs(res6 - float(g["res"])) < 1e-4


def test_nls_max_iteration_warning_and_batch(cuda):
    """Batched launch: every problem of a ragged batch equals its single-problem solve; a problem
    that cannot converge in 30 iterations reports it (the reference prints a WARNING)."""
    import torch

    from velocity_b200 import NLS, synth

    rng = np.random.default_rng(0)
    K = synth.K_1080P
    sizes = [4, 151, 1, 4096, 37]
    ps, pws, x0s = [], [], []
    for n in sizes:
        pw = synth.scene_points(n, seed=n)
        t = rng.normal(0, 0.1, 3)
        uv = (pw + t) @ K
        ps.append(uv[:, :2] / uv[:, 2:3] + rng.normal(0, 0.2, (n, 2)))
        pws.append(pw)
        x0s.append(np.array([0.0, 0.0, 0.5]))
    first = np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int32)
    x, iters = NLS.nls_batch_device(NLS._dev64(K), NLS._dev64(np.concatenate(ps)), NLS._dev64(np.concatenate(pws)),
                                    torch.from_numpy(first).cuda(), torch.tensor(sizes, dtype=torch.int32).cuda(),
                                    NLS._dev64(np.stack(x0s)), 3)
    x, iters = x.cpu().numpy(), iters.cpu().numpy()
    for k, n in enumerate(sizes):
        xs, it = NLS._single(K, ps[k], pws[k], x0s[k], 3)
        assert np.array_equal(xs, x[k]) and it == iters[k]
    # unreachable tolerance in 30 iterations: wildly inconsistent measurements
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        NLS.fcnNLS_t(K, rng.uniform(0, 1900, (50, 2)), synth.scene_points(50, seed=1) * [1, 1, 0.001], np.array([0.0, 0, 1]))
    assert "WARNING: fcnNLS_t() reaching max iterations!" in buf.getvalue()


def test_nls_iteration_cap_is_reported_per_problem(cuda):
    """utils/NLS.py:126-127 (the for/else WARNING): the oracle and the GPU agree on WHICH problems hit the 30-iteration cap."""
    import torch

    from oracle import sfm_oracle as S
    from velocity_b200 import NLS, synth

    rng = np.random.default_rng(5)
    K = synth.K_1080P
    good_pw = synth.scene_points(64, seed=3)
    uv = (good_pw + [0.1, 0.0, 0.2]) @ K
    cases = [(uv[:, :2] / uv[:, 2:3], good_pw), (rng.uniform(0, 1900, (50, 2)), synth.scene_points(50, seed=1) * [1, 1, 0.001])]
    x0 = np.array([0.0, 0.0, 1.0])
    for p, pw in cases:
        _, ok = S.solve_translation(K, p, pw, x0)
        _, it = NLS._single(K, p, pw, x0, 3)
        assert (it > 0) == ok, (it, ok)
    assert S.solve_translation(K, *cases[0], x0)[1] and not S.solve_translation(K, *cases[1], x0)[1]


def test_triangulation(cuda):
    from oracle import sfm_oracle as S
    from velocity_b200 import MSV

    g = golden("triangulate")
    c2 = MSV.fcn2vintercept(g["A"], g["U"])
    cn = MSV.fcnNvintercept(g["A"], g["U"])
    assert c2.dtype == np.float64 and c2.shape == g["c2v"].shape
    assert np.allclose(c2, g["c2v"], rtol=1e-9, atol=1e-9) and np.allclose(cn, g["cnv"], rtol=1e-9, atol=1e-9)
    assert np.allclose(c2, S.triangulate_pairs(g["A"], g["U"]), rtol=1e-10, atol=1e-10)


def test_triangulation_long_sequence_recovers_scene(cuda):
    """BASELINE config 3 shape (300 frames x 4096 tracks): noise-free rays intersect at the scene points."""
    from velocity_b200 import MSV, synth
    from velocity_b200.common import pixel2uvec

    nt, nf = 4096, 300
    pw = synth.scene_points(nt, seed=2)
    K = synth.K_1080P
    A = np.stack([np.array([0.002 * j, -0.001 * j, 0.02 * j]) for j in range(nf)])  # camera origins
    U = np.zeros((3, nf, nt))
    for j in range(nf):
        uv = (pw - A[j]) @ K
        U[:, j] = pixel2uvec(K, uv[:, :2] / uv[:, 2:3]).T
    for fn in (MSV.fcnNvintercept, MSV.fcn2vintercept):
        c = fn(A, U)
        assert np.abs(c - pw).max() < 1e-6, fn.__name__


def test_fcnMSV1_t(cuda):
    from velocity_b200 import MSV

    g = golden("msv_t")
    x, b0 = MSV.fcnMSV1_t(g["K"], g["P"], g["B"], g["vg"], int(g["ii"]))
    assert x.dtype == np.float32 and b0.shape == g["b0"].shape
    assert np.allclose(x, g["x1"], rtol=1e-5, atol=1e-6)
    assert np.allclose(b0, g["b0"], rtol=1e-6, atol=1e-7)
    with pytest.raises(NotImplementedError):
        MSV.fcnMSV2_t(g["K"], g["P"], g["B"], g["vg"], 2)


@pytest.mark.parametrize("name", ["ba_small", "ba_medium", "ba_256x10", "ba_512x20"])
def test_fcnNLS_batch(cuda, name):
    from velocity_b200 import NLS

    g = golden(name)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        cw, pw = NLS.fcnNLS_batch(g["K"], g["P"].copy(), g["pw0"].copy(), g["cw0"].copy())
    assert cw.shape == g["cw"].shape and pw.shape == g["pw"].shape and cw.dtype == np.float64
    assert np.allclose(cw, g["cw"], rtol=1e-6, atol=1e-7)
    assert np.allclose(pw, g["pw"], rtol=1e-6, atol=1e-7)

    def iterations(text):
        return sum(1 for ln in text.splitlines() if "f=" in ln and "x=" in ln)

    assert iterations(buf.getvalue()) == iterations(str(g["stdout"]))
    assert "fcnNLS_batch done in" in buf.getvalue()


def test_fcnNLS_batch2(cuda):
    from velocity_b200 import NLS

    g = golden("ba_small")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        cw, pw = NLS.fcnNLS_batch2(g["K"], g["P"].copy(), g["pw0"].copy(), g["cw0"].copy())
    assert cw.shape == g["cw_b2"].shape and pw.shape == g["pw_b2"].shape
    assert np.allclose(cw, g["cw_b2"], rtol=1e-6, atol=1e-7)
    assert np.allclose(pw, g["pw_b2"], rtol=1e-6, atol=1e-7)
    ref_steps = [ln for ln in str(g["stdout_b2"]).splitlines() if "fcnNLS_batch2 done in" in ln][0].split("done in")[1].split("steps")[0]
    our_steps = [ln for ln in buf.getvalue().splitlines() if "fcnNLS_batch2 done in" in ln][0].split("done in")[1].split("steps")[0]
    assert ref_steps.strip() == our_steps.strip()


@pytest.mark.parametrize("name", ["ba_small", "ba_512x20"])
def test_ba_device_loop_equals_stepwise_loop(cuda, name):
    """vel_ba_iterate (the whole loop enqueued at once, convergence test on the device, cross blocks only in scaled form)
    against the stage-by-stage entry points with the host-side test: same number of iterations, same history, same result.
    The two differ in rounding only (one reciprocal per projection, W' formed directly)."""
    from oracle import sfm_oracle as S
    from velocity_b200 import NLS

    g = golden(name)
    z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
    a = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
    hist_a = []
    for _ in range(10):
        hist_a.append(a.step())
        if hist_a[-1][1] < 1e-7:
            break
    b = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
    hist_b = b.iterate(10, 1e-7)
    assert len(hist_b) == len(hist_a)
    assert int(b._iters.item()) == len(hist_b)
    ha, hb = np.array(hist_a), np.array(hist_b)
    assert np.allclose(hb[:, 0], ha[:, 0], rtol=1e-9) and np.allclose(hb[:, 1], ha[:, 1], rtol=1e-3, atol=1e-10)
    xa, xb = a.x.cpu().numpy(), b.x.cpu().numpy()
    # the weakly determined directions (point depth) amplify the forward-difference noise of the two roundings to ~2e-9 relative
    assert np.abs(xa - xb).max() <= 1e-7 * np.abs(xa).max()
    # a loop that is cut short by max_iter reports exactly max_iter rows
    c = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
    assert len(c.iterate(2, 1e-7)) == 2 and int(c._iters.item()) == 2
    # tolerance that the first iteration already meets: one iteration runs, the other nine return at the gate (their rows stay NaN)
    d = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
    assert len(d.iterate(10, 1e30)) == 1 and int(d._iters.item()) == 1
    assert np.isnan(d._hist[1:10].cpu().numpy()).all()
    assert np.abs(d.x.cpu().numpy() - c.x.cpu().numpy()).max() > 0          # c ran two iterations, d one


def test_ba_blocks_match_oracle(cuda):
    """K7 output blocks vs the numpy block oracle on one linearisation (tight: same forward differences)."""
    from oracle import sfm_oracle as S
    from velocity_b200 import NLS

    g = golden("ba_medium")
    z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
    V, U, W, gg, cost = S.ba_blocks(g["K"], x, z, nt, nc)
    ba = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
    ba.accumulate()
    iu3, iu6 = np.triu_indices(3), np.triu_indices(6)

    def close(a, b):  # forward-difference noise scales with the largest entry of the block set
        return np.abs(a - b).max() <= 1e-8 * np.abs(b).max()

    assert close(ba.V.cpu().numpy(), V[:, iu3[0], iu3[1]])
    assert close(ba.U.cpu().numpy(), U[:, iu6[0], iu6[1]])
    Wm = W.transpose(0, 2, 1, 3).reshape(6 * nc, 3 * nt)
    assert close(ba.W.cpu().numpy(), Wm)
    assert close(ba.g.cpu().numpy(), gg)
    assert abs(ba.cost.item() - cost) <= 1e-9 * cost


@pytest.mark.parametrize("name,first,count", [("ba_512x20", 0, 20), ("ba_512x20", 3, 9), ("ba_256x10", 0, 10)])
def test_ba_blocks_multi_chunk_and_camera_slices(cuda, name, first, count):
    """K7 with >= 16 cameras takes the chunked point path (ba_point_kernel chunk indexing + ba_point_reduce_kernel);
    a camera slice is what one rank of the sharded form computes.  Both against the numpy block oracle."""
    from oracle import sfm_oracle as S
    from velocity_b200 import NLS, _lib
    from velocity_b200.device import ptr, stream_ptr

    g = golden(name)
    z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
    V, U, W, gg, cost = S.ba_blocks(g["K"], x, z, nt, nc, first, count)
    ba = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
    _lib.check(_lib.lib().vel_ba_accumulate(ptr(ba.K), ptr(ba.x), ptr(ba.z), nt, nc, first, count, ptr(ba.V), ptr(ba.U), ptr(ba.W),
                                            ptr(ba.g), ptr(ba.cost), stream_ptr()), "vel_ba_accumulate")
    iu3, iu6 = np.triu_indices(3), np.triu_indices(6)

    def close(a, b):
        return np.abs(a - b).max() <= 1e-8 * max(np.abs(b).max(), 1e-300)

    lo, hi = max(first, 1) - 1, first + count - 1          # parameterised cameras of the slice own rows lo..hi-1
    assert close(ba.V.cpu().numpy(), V[:, iu3[0], iu3[1]])
    assert close(ba.U.cpu().numpy()[lo:hi], U[lo:hi][:, iu6[0], iu6[1]])
    Wm = W.transpose(0, 2, 1, 3).reshape(6 * nc, 3 * nt)
    assert close(ba.W.cpu().numpy()[6 * lo:6 * hi], Wm[6 * lo:6 * hi])
    gd = ba.g.cpu().numpy()
    assert close(gd[:3 * nt], gg[:3 * nt])
    assert close(gd[3 * nt + 3 * lo:3 * nt + 3 * hi], gg[3 * nt + 3 * lo:3 * nt + 3 * hi])
    assert close(gd[3 * nt + 3 * nc + 3 * lo:3 * nt + 3 * nc + 3 * hi], gg[3 * nt + 3 * nc + 3 * lo:3 * nt + 3 * nc + 3 * hi])
    assert abs(ba.cost.item() - cost) <= 1e-9 * cost


def test_bundle_adjustment_c3_size_matches_sparse_oracle(cuda):
    """BASELINE configs[2] size (nt=4096, nc=299, nx=14,082): the reference's dense form cannot run there (277 GB
    Jacobian); the fixture is the oracle's block-sparse form after the reference's 10-iteration loop (it stops at
    iteration 8 on rms(delta) < 1e-7).  Tolerance 1e-4 relative; measured agreement is ~1e-9."""
    import zlib

    from util import ba_c3_inputs
    from velocity_b200 import NLS

    g = golden("ba_c3_sparse")
    K, P, pw0, cw0 = ba_c3_inputs()
    assert zlib.crc32(pw0.tobytes()) == int(g["pw0_crc"]) and zlib.crc32(P.tobytes()) == int(g["P_crc"])   # same inputs
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        cw, pw = NLS.fcnNLS_batch(K, P, pw0, cw0)
    n_it = sum(1 for ln in buf.getvalue().splitlines() if "f=" in ln and "x=" in ln)
    assert n_it == len(g["hist"])
    assert np.abs(cw - g["cw"]).max() <= 1e-4 * np.abs(g["cw"]).max()
    assert np.abs(pw - g["pw"]).max() <= 1e-4 * np.abs(g["pw"]).max()
    assert np.abs(cw - g["cw"]).max() < 1e-7 and np.abs(pw - g["pw"]).max() < 1e-6       # what is actually achieved


def test_match_knn2(cuda):
    from oracle import cv_oracle as O
    from velocity_b200 import match

    g = golden("match_knn2")
    idx, dist = match.knn2_hamming256(g["q"], g["t"])
    assert idx.dtype == np.int32 and np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"].astype(np.int32))
    idx, dist = match.knn2_l2(g["qf"], g["tf"])
    oi, od = O.knn2_l2(g["qf"], g["tf"])
    assert np.array_equal(idx, g["idxf"]) and np.array_equal(idx, oi) and np.array_equal(dist, od)
    assert np.abs(dist - g["distf"]).max() < 1e-5


def test_match_full_size_and_ragged(cuda):
    """BASELINE config 4 size (8192 x 8192 x 256 bit) against the oracle; ragged / tiny train sets."""
    from oracle import cv_oracle as O
    from velocity_b200 import match

    rng = np.random.default_rng(4)
    t = rng.integers(0, 256, (8192, 32), dtype=np.uint8)
    q = t[rng.permutation(8192)].copy()
    flip = rng.integers(0, 256, (8192, 2))
    for k in range(2):
        q[np.arange(8192), flip[:, k] // 8] ^= (1 << (flip[:, k] % 8)).astype(np.uint8)
    idx, dist = match.knn2_hamming256(q, t)
    oi, od = O.knn2_hamming(q, t)
    assert np.array_equal(idx, oi) and np.array_equal(dist, od)
    for nq, nt in [(1, 1), (5, 2), (130, 129), (3, 1000)]:
        idx, dist = match.knn2_hamming256(q[:nq], t[:nt])
        oi, od = O.knn2_hamming(q[:nq], t[:nt])
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt)


def test_match_tensor_core_and_popcount_forms_agree(cuda, monkeypatch):
    """Both forms of K4 (tcgen05 int8 GEMM with fused top-2, and xor+popcount) against the oracle on
    ragged sizes (partial query / train tiles) with exact duplicates (lowest train index must win)."""
    from oracle import cv_oracle as O
    from velocity_b200 import match

    rng = np.random.default_rng(9)
    for nq, nt in [(128, 256), (200, 300), (1000, 777), (2049, 1025)]:
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        q = t[rng.integers(0, nt, nq)].copy()
        q[:, 3] ^= rng.integers(0, 4, nq).astype(np.uint8)
        t[17] = t[4]
        t[nt - 1] = t[nt // 2]
        oi, od = O.knn2_hamming(q, t)
        for mode in ("tc", "popc"):
            monkeypatch.setenv("VEL_MATCH_FORCE", mode)
            idx, dist = match.knn2_hamming256(q, t)
            assert np.array_equal(idx, oi) and np.array_equal(dist, od), (mode, nq, nt)
