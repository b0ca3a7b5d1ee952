"""Parity of the CUDA tracker path (K1 pyramid, K2 LK + forward-backward, K3 remap, KLTregional,
KLTmain) against the CPU oracle (bit-exact) and the reference-generated golden vectors."""
import numpy as np
import pytest

from util import LK_CASES, LK_ERR_TOL, LK_POINT_TOL_PX, PRIM_CASES, fbt_of, golden, lk_images, lk_kwargs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from velocity_b200 import _lib

    _lib.lib()
    return torch


@pytest.mark.parametrize("name", PRIM_CASES)
def test_pyramid_remap_decimate_bit_exact(cuda, name):
    from oracle import cv_oracle as O
    from velocity_b200 import KLT
    from velocity_b200.lk import FrameBatch

    g = golden(name)
    im = g["im"]
    d = cuda.from_numpy(im).cuda()
    fb = FrameBatch(d, (3, 3), 3).build()
    lvl = im
    for l in range(1, fb.layout.max_level + 1):
        lvl = O.pyrDown(lvl)
        assert np.array_equal(fb.level(0, l).cpu().numpy(), lvl), "pyramid level %d" % l
    assert np.array_equal(fb.level(0, 1).cpu().numpy(), g["pyrdown"])
    assert np.array_equal(KLT._decimate4_device(d).cpu().numpy(), g["quarter"])
    x0, x1, y0, y1 = (int(v) for v in g["roi"])
    out = KLT._remap_affine_device(d, g["T"], x0, x1, y0, y1).cpu().numpy()
    assert np.array_equal(out, g["remap"])
    assert np.array_equal(out, O.remap_affine(im, g["T"], x0, x1, y0, y1))


def test_pyramid_batch_1080p_matches_oracle(cuda):
    from oracle import cv_oracle as O
    from velocity_b200 import synth
    from velocity_b200.lk import FrameBatch

    frames = np.stack([synth.texture(1080, 1920, s) for s in (1, 2, 3)])
    fb = FrameBatch(cuda.from_numpy(frames).cuda(), (15, 15), 4).build()
    assert fb.layout.max_level == 4
    for i in range(3):
        lvl = frames[i]
        for l in range(1, 5):
            lvl = O.pyrDown(lvl)
            assert np.array_equal(fb.level(i, l).cpu().numpy(), lvl), (i, l)


@pytest.mark.parametrize("h,w,strip", [(1080, 1920, None), (270, 480, None), (64, 64, None), (33, 48, None), (17, 32, None), (542, 976, None),
                                       (135, 1936, None), (1080, 1920, 32), (542, 976, 32), (135, 1936, 32), (270, 480, 32), (2160, 3840, 32),
                                       (1080, 1920, 24)])
def test_pyramid_fused_two_level_kernel_equals_level_by_level(cuda, monkeypatch, h, w, strip):
    """K1's fused kernel (levels 1 and 2 from one pass over the frame; widths that are multiples of 16) against the
    level-by-level kernels and the oracle: even and odd level-1 heights (the bottom-edge reflection does not commute with
    the filter), strips that hang over the image, single-warp and multi-warp rows, several frames per launch."""
    from oracle import cv_oracle as O
    from velocity_b200 import synth
    from velocity_b200.lk import FrameBatch

    rng = np.random.default_rng(h * 10007 + w)
    frames = np.stack([synth.texture(h, w, 5) if h >= 64 and w >= 64 else rng.integers(0, 256, (h, w), dtype=np.uint8),
                       rng.integers(0, 256, (h, w), dtype=np.uint8)])
    d = cuda.from_numpy(frames).cuda()
    monkeypatch.setenv("VEL_PYR_FUSED", "1")
    if strip:
        monkeypatch.setenv("VEL_PYR_H2", str(strip))
    fused = FrameBatch(d, (3, 3), 3).build()
    monkeypatch.setenv("VEL_PYR_FUSED", "0")
    plain = FrameBatch(d, (3, 3), 3).build()
    assert fused.layout.max_level >= 2
    for i in range(2):
        lvl = frames[i]
        for l in range(1, fused.layout.max_level + 1):
            lvl = O.pyrDown(lvl)
            assert np.array_equal(fused.level(i, l).cpu().numpy(), lvl), (i, l)
            assert np.array_equal(plain.level(i, l).cpu().numpy(), lvl), (i, l)


@pytest.mark.parametrize("name", LK_CASES)
def test_lk_matches_oracle_bit_exact_and_reference_golden(cuda, name):
    from oracle import klt_oracle as KO
    from velocity_b200 import KLT

    g = golden(name)
    im0, im1 = lk_images(g)
    lk, fbt = lk_kwargs(g), fbt_of(g)
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, g["p"], None, fbt=fbt, **lk)
    assert p2.dtype == np.float32 and p2.shape == g["p2"].shape
    assert v.dtype == bool and v.shape == g["v"].shape
    assert err.dtype == np.float32 and err.shape == g["err"].shape
    # (1) CUDA == oracle, bit for bit (same integer accumulation, same float32 op order)
    o2, ov, oerr = KO.lk_forward_backward(im0, im1, g["p"], fbt=fbt, **lk)
    assert np.array_equal(v, ov)
    assert np.array_equal(p2, o2)
    fwd_ok = oerr.ravel() != 0
    assert np.array_equal(err[fwd_ok], oerr[fwd_ok])
    # (2) CUDA vs the reference's own output (cv2 4.13): masks exact, points within tolerance
    assert np.array_equal(v, g["v"])
    assert np.abs(p2 - g["p2"])[g["v"]].max() <= LK_POINT_TOL_PX
    assert np.abs(err - g["err"])[g["v"]].max() <= LK_ERR_TOL


@pytest.mark.parametrize("name,translate", [("regional_translate", True), ("regional_affine", False)])
def test_kltregional(cuda, name, translate):
    from oracle import klt_oracle as KO
    from velocity_b200 import KLT

    g = golden(name)
    lk = lk_kwargs(g)
    p, v = KLT.KLTregional(g["im0"], g["im1"], g["p0"], g["T"], lk, fbt=float(g["fbt"]), translateFlag=translate)
    po, vo = KO.klt_regional(g["im0"], g["im1"], g["p0"], g["T"], lk, fbt=float(g["fbt"]), translate=translate)
    assert np.array_equal(v, vo) and np.array_equal(p, po)
    assert p.dtype == g["p"].dtype and np.array_equal(v, g["v"])
    assert np.abs(p - g["p"])[g["v"]].max() <= LK_POINT_TOL_PX


def test_kltmain(cuda):
    from velocity_b200 import KLT

    g = golden("kltmain_pair")
    p, v, im_small = KLT.KLTmain(g["im1"], g["im0"], None, g["p0"])
    assert np.array_equal(v, g["v"])
    assert p.shape == g["p"].shape and np.abs(p - g["p"]).max() <= LK_POINT_TOL_PX
    assert np.array_equal(im_small.cpu().numpy(), g["im_small"])


def test_lk_full_size_properties(cuda):
    """BASELINE config 2 size (1080p, 4096 tracks, 15x15, 3 levels, FB 1.0): identity pair returns
    the input points with zero error; a pure integer shift is recovered; CUDA == oracle on a sample."""
    from oracle import klt_oracle as KO
    from velocity_b200 import KLT, synth

    im0 = synth.texture(1080, 1920, 1234)
    pts = synth.harris_tracks(im0, 4096)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im0, pts, None, fbt=1.0, **lk)
    assert v.all() and np.abs(p2 - pts).max() <= 1e-3 and err.max() == 0
    im1 = np.roll(im0, (2, 3), (0, 1))
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, pts, None, fbt=1.0, **lk)
    assert v.mean() > 0.99
    assert np.abs(p2[v] - pts[v] - np.float32([3, 2])).max() < 0.05
    sel = np.arange(0, 4096, 16)
    o2, ov, _ = KO.lk_forward_backward(im0, im1, pts[sel], fbt=1.0, **lk)
    assert np.array_equal(v[sel], ov) and np.array_equal(p2[sel], o2)


def test_lk_w15_word_kernel_equals_byte_kernel_and_oracle(cuda, monkeypatch):
    """The default 15x15 kernel (two points per warp) gathers aligned words and re-aligns them per
    lane; the byte-gather kernel (VEL_LK_W15=bytes) is the independent implementation.  Both must
    equal the oracle bit for bit on ROI views with every base misalignment, on points hugging / leaving the frame border,
    on a frame whose width is not a multiple of 4, and on the minimum 16-px-wide top level."""
    from oracle import klt_oracle as KO
    from velocity_b200 import KLT, synth

    rng = np.random.default_rng(7)
    big0 = synth.texture(300, 420, 11)
    big1 = np.roll(big0, (2, -3), (0, 1))
    d0, d1 = cuda.from_numpy(big0).cuda(), cuda.from_numpy(big1).cuda()
    lk = dict(winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.03))
    cases = []
    for x0 in (0, 1, 2, 3, 5):           # base pointer misalignment 0..3 (pitch stays 420)
        y0, hh, ww = 3 + x0, 240, 333 + x0
        cases.append((d0[y0:y0 + hh, x0:x0 + ww], d1[y0:y0 + hh, x0:x0 + ww], big0[y0:y0 + hh, x0:x0 + ww], big1[y0:y0 + hh, x0:x0 + ww]))
    odd0, odd1 = np.ascontiguousarray(big0[:131, :203]), np.ascontiguousarray(big1[:131, :203])   # host upload, width % 4 != 0
    cases.append((odd0, odd1, odd0, odd1))
    for a, b, ha, hb in cases:
        h, w = ha.shape
        inner = synth.harris_tracks(np.ascontiguousarray(ha), 150, border=2)
        edge = np.stack([rng.uniform(-12, w + 12, 120), rng.uniform(-12, h + 12, 120)], 1).astype(np.float32)
        rim = np.float32([[0, 0], [w - 1, h - 1], [7.5, 7.5], [w - 8.5, h - 8.5], [8, h / 2], [w - 9, h / 2], [w / 2, 8.01], [w / 2, h - 9.2],
                          [-14.99, 30], [w - 0.01, 40], [50, -14.9], [60, h - 0.5]])
        pts = np.concatenate([inner, edge, rim]).astype(np.float32)
        for fbt in (None, 1.0):
            o2, ov, oerr = KO.lk_forward_backward(np.ascontiguousarray(ha), np.ascontiguousarray(hb), pts, fbt=fbt, **lk)
            res = {}
            for impl in ("default", "bytes"):     # default = word-gathering kernel, two points per warp
                if impl == "default":
                    monkeypatch.delenv("VEL_LK_W15", raising=False)
                else:
                    monkeypatch.setenv("VEL_LK_W15", impl)
                res[impl] = KLT.cv2calcOpticalFlowPyrLK(a, b, pts, None, fbt=fbt, **lk)
            monkeypatch.delenv("VEL_LK_W15", raising=False)
            for impl, (p2, v, err) in res.items():
                assert np.array_equal(v, ov), (impl, ha.shape, fbt)
                assert np.array_equal(p2, o2), (impl, ha.shape, fbt)
                ok = oerr.ravel() != 0
                assert np.array_equal(err[ok], oerr[ok]), (impl, ha.shape, fbt)
            assert ov.any() and not ov.all()     # the case mixes tracked, lost and never-inside points
    # the mid-size-window kernel (51x51 lk_fine, cv2's default 21x21, a rectangular one): word path vs byte path vs oracle
    a, b, ha, hb = cases[2]
    h, w = ha.shape
    pts = np.concatenate([synth.harris_tracks(np.ascontiguousarray(ha), 60, border=2),
                          np.stack([rng.uniform(-20, w + 20, 40), rng.uniform(-20, h + 20, 40)], 1)]).astype(np.float32)
    for win, lvl in (((51, 51), 0), ((21, 21), 2), ((19, 33), 1)):
        lkw = dict(winSize=win, maxLevel=lvl, criteria=(3, 20, 0.003))
        o2, ov, oerr = KO.lk_forward_backward(np.ascontiguousarray(ha), np.ascontiguousarray(hb), pts, fbt=0.5, **lkw)
        for impl in ("default", "bytes"):
            if impl == "default":
                monkeypatch.delenv("VEL_LK_W15", raising=False)
            else:
                monkeypatch.setenv("VEL_LK_W15", impl)
            p2, v, err = KLT.cv2calcOpticalFlowPyrLK(a, b, pts, None, fbt=0.5, **lkw)
            ok = oerr.ravel() != 0
            assert np.array_equal(v, ov) and np.array_equal(p2, o2) and np.array_equal(err[ok], oerr[ok]), (win, impl)
        monkeypatch.delenv("VEL_LK_W15", raising=False)
        assert ov.any() and not ov.all()
    # smallest legal top level: 16 px wide / high
    tiny0 = synth.texture(64, 64, 3)
    tiny1 = np.roll(tiny0, (1, 1), (0, 1))
    pts = rng.uniform(0, 63, (64, 2)).astype(np.float32)
    lk2 = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.03))
    o2, ov, _ = KO.lk_forward_backward(tiny0, tiny1, pts, fbt=1.0, **lk2)
    p2, v, _ = KLT.cv2calcOpticalFlowPyrLK(tiny0, tiny1, pts, None, fbt=1.0, **lk2)
    assert np.array_equal(v, ov) and np.array_equal(p2, o2)


def test_lk_randomized_parameters_match_oracle(cuda):
    """Seeded sweep over the parameter space of cv2.calcOpticalFlowPyrLK as the wrapper exposes it (util.lk_sweep_cases):
    status masks, points and errors equal the oracle bit for bit (the oracle itself is checked against cv2 on the same
    sweep by tests/test_oracle_cv.py)."""
    from oracle import klt_oracle as KO
    from util import lk_sweep_cases
    from velocity_b200 import KLT

    for k, im0, im1, pts, lk, fbt in lk_sweep_cases():
        p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, pts, None, fbt=fbt, **lk)
        o2, ov, oerr = KO.lk_forward_backward(im0, im1, pts, fbt=fbt, **lk)
        ok = oerr.ravel() != 0
        assert np.array_equal(v, ov) and np.array_equal(p2, o2) and np.array_equal(err[ok], oerr[ok]), (k, lk, fbt)


def test_full_size_against_cv2_itself(cuda):
    """BASELINE config 2 at full size against the reference's own arithmetic provider run on THIS host
    (opencv-python, when importable): forward-backward masks identical except where cv2's float32-lane
    accumulation flips a stop criterion (<= 0.7 % of points), points of commonly valid
    tracks within the stated tolerance, on a rendered consecutive pair of the synthetic sequence."""
    cv2 = pytest.importorskip("cv2")
    from velocity_b200 import KLT, synth

    frames, _ = synth.plane_sequence(2, h=1080, w=1920, seed=1234, Z0=40.0)
    pts = synth.harris_tracks(frames[0], 4096)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(frames[0], frames[1], pts, None, fbt=1.0, **lk)
    q2, st, qerr = cv2.calcOpticalFlowPyrLK(frames[0], frames[1], pts, None, **lk)
    q1, st1, _ = cv2.calcOpticalFlowPyrLK(frames[1], frames[0], q2, None, **lk)
    d = pts - q1
    qv = st.ravel().astype(bool) & st1.ravel().astype(bool) & (np.sqrt(d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) < np.float32(1.0))
    assert (v != qv).mean() <= 0.007
    both = v & qv
    assert both.mean() > 0.98
    assert np.abs(p2 - q2)[both].max() <= LK_POINT_TOL_PX
    assert np.abs(err - qerr)[both].max() <= LK_ERR_TOL


def test_klt_regional_entry_point_equals_the_staged_calls(cuda):
    """vel_klt_regional (one C call: crop, shift / remap, pyramids of the crops, LK + backward gate) against the same steps issued
    one by one through vel_remap_affine_u8 / vel_pyramid_u8 / vel_lk_track on ROI views -- bit for bit, for the whole frame with a
    zero shift, a shifted ROI at an odd offset and an affine ROI; and the argument checks of the entry point."""
    import ctypes as C

    from velocity_b200 import KLT, _lib, synth
    from velocity_b200.device import ptr, stream_ptr
    from velocity_b200.lk import FrameBatch, lk_params, track_pairs

    im0 = synth.texture(300, 420, 9)
    im1 = np.roll(im0, (2, -3), (0, 1))
    d0, d1 = cuda.from_numpy(im0).cuda(), cuda.from_numpy(im1).cuda()
    pts = synth.harris_tracks(im0[60:240, 80:340], 200, border=10) + np.float32([80, 60])
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    params = lk_params(fbt=1.0, **lk)
    dp = cuda.from_numpy(pts).cuda()

    def staged(prev_roi, next_roi, p_roi):
        fa, fb = FrameBatch(prev_roi, (15, 15), 2).build(), FrameBatch(next_roi, (15, 15), 2).build()
        out, st, err, _ = track_pairs(fa, fb, p_roi, params)
        return out[0], st[0], err[0]

    T_id = np.float32([[1, 0], [0, 1], [0, 0]])
    a = KLT._regional_device(d0, d1, dp, 0, 420, 0, 300, T_id, True, 1.0, lk)
    b = staged(d0, d1, dp)
    assert all(cuda.equal(x, y) for x, y in zip(a, b)) and int(a[1].sum()) > 150
    x0, x1, y0, y1 = 31, 391, 11, 289                       # odd offsets: unaligned ROI views
    T_sh = np.float32([[1, 0], [0, 1], [-3.7, 2.2]])        # int() truncates toward zero: shift (-3, 2)
    a = KLT._regional_device(d0, d1, dp, x0, x1, y0, y1, T_sh, True, 1.0, lk)
    b = staged(d0[y0:y1, x0:x1], d1[y0 + 2:y1 + 2, x0 - 3:x1 - 3], dp - cuda.tensor([x0, y0], dtype=cuda.float32, device="cuda"))
    assert all(cuda.equal(x, y) for x, y in zip(a, b)) and int(a[1].sum()) > 150
    T_af = np.float32([[1.001, 0.002], [-0.002, 0.999], [-3.1, 2.3]])
    a = KLT._regional_device(d0, d1, dp, x0, x1, y0, y1, T_af, False, 1.0, lk)
    b = staged(d0[y0:y1, x0:x1], KLT._remap_affine_device(d1, T_af, x0, x1, y0, y1), dp - cuda.tensor([x0, y0], dtype=cuda.float32, device="cuda"))
    assert all(cuda.equal(x, y) for x, y in zip(a, b)) and int(a[1].sum()) > 150
    with pytest.raises(RuntimeError, match="leaves the frame"):
        KLT._regional_device(d0, d1, dp, x0, x1, y0, y1, np.float32([[1, 0], [0, 1], [-40, 0]]), True, 1.0, lk)
    with pytest.raises(RuntimeError, match="outside the"):
        KLT._regional_device(d0, d1, dp, x0, 500, y0, y1, T_id, True, 1.0, lk)


def test_4k_frames_and_the_10m_generator_where_most_tracks_fail(cuda):
    """VERDICT r1 weak point 3: (a) a C5-size pair (3840 x 2160, 4 levels) and (b) SURVEY 8(d)'s literal generator (plane at 10 m:
    up to 37 px of flow per frame, ~4 of 5 tracks fail the forward-backward gate, so the failure paths -- bounds exits at every
    level, the minimum-eigenvalue test, 10-iteration non-convergence -- dominate) against the oracle, bit for bit."""
    from oracle import klt_oracle as KO
    from velocity_b200 import KLT, synth

    lk = dict(winSize=(15, 15), maxLevel=3, criteria=(3, 10, 0.1))
    frames, _ = synth.plane_sequence(2, h=2160, w=3840, seed=77, Z0=40.0)
    pts = synth.harris_tracks(frames[0], 1536)
    pts = np.concatenate([pts, np.float32([[3.0, 4.0], [3836.5, 2157.25], [1920.0, 6.5], [-20.0, 800.0], [3839.9, 1000.0]])])
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(frames[0], frames[1], pts, None, fbt=1.0, **lk)
    o2, ov, oerr = KO.lk_forward_backward(frames[0], frames[1], pts, fbt=1.0, **lk)
    ok = oerr.ravel() != 0
    assert np.array_equal(v, ov) and np.array_equal(p2, o2) and np.array_equal(err[ok], oerr[ok])
    assert v.mean() > 0.9

    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    frames, _ = synth.plane_sequence(2, h=1080, w=1920, seed=1234, Z0=10.0)
    pts = synth.harris_tracks(frames[0], 2048)
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(frames[0], frames[1], pts, None, fbt=1.0, **lk)
    o2, ov, oerr = KO.lk_forward_backward(frames[0], frames[1], pts, fbt=1.0, **lk)
    ok = oerr.ravel() != 0
    assert np.array_equal(v, ov) and np.array_equal(p2, o2) and np.array_equal(err[ok], oerr[ok])
    assert 0.05 < v.mean() < 0.6          # the regime SURVEY describes: most tracks are lost


def test_edge_cases_empty_single_and_errors(cuda):
    """Empty and single-point inputs, points far outside the frame (status 0, no crash), argument
    errors surfaced as RuntimeError with the C ABI's message; a point set on a pure-constant image
    fails the minimum-eigenvalue test exactly like cv2 (status 0)."""
    from oracle import klt_oracle as KO
    from velocity_b200 import KLT, synth

    im0 = synth.texture(200, 260, 5)
    im1 = np.roll(im0, (1, 1), (0, 1))
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, np.zeros((0, 2), np.float32), None, fbt=1.0, **lk)
    assert p2.shape == (0, 2) and v.shape == (0,) and err.shape == (0, 1)
    pts = np.float32([[100.5, 80.25]])
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, pts, None, fbt=1.0, **lk)
    o2, ov, oerr = KO.lk_forward_backward(im0, im1, pts, fbt=1.0, **lk)
    assert np.array_equal(p2, o2) and np.array_equal(v, ov)
    far = np.float32([[-500, -500], [1e6, 40], [130, 1e7], [259.9, 199.9], [-14.9, -14.9]])
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, far, None, fbt=1.0, **lk)
    o2, ov, _ = KO.lk_forward_backward(im0, im1, far, fbt=1.0, **lk)
    assert np.array_equal(v, ov) and np.array_equal(p2, o2) and not v[:3].any()
    flat = np.full((120, 160), 77, np.uint8)
    p2, v, err = KLT.cv2calcOpticalFlowPyrLK(flat, flat, np.float32([[60, 60], [80, 40]]), None, **lk)
    assert not v.any()
    for win in [(51, 51), (21, 21), (9, 13)]:      # every kernel variant agrees with the oracle on the same pair
        lkw = dict(winSize=win, maxLevel=1, criteria=(3, 15, 0.01))
        pts = synth.harris_tracks(im0, 64, border=8)
        p2, v, err = KLT.cv2calcOpticalFlowPyrLK(im0, im1, pts, None, fbt=0.5, **lkw)
        o2, ov, _ = KO.lk_forward_backward(im0, im1, pts, fbt=0.5, **lkw)
        assert np.array_equal(v, ov) and np.array_equal(p2, o2), win
    with pytest.raises(ValueError):
        KLT.cv2calcOpticalFlowPyrLK(im0, im1[:-1], pts, None, **lk)
    with pytest.raises(RuntimeError, match="must be larger than the window"):
        KLT.cv2calcOpticalFlowPyrLK(im0[:12, :12], im1[:12, :12], pts, None, **lk)
    with pytest.raises(NotImplementedError):
        KLT.cv2calcOpticalFlowPyrLK(im0, im1, pts, None, flags=4, **lk)


def test_sequence_tracker_host_api_matches_per_pair_calls(cuda):
    """The public host-buffer entry point (what bench.py's e2e number times): chunked H2D pipeline,
    one pyramid per frame, results identical to calling cv2calcOpticalFlowPyrLK pair by pair."""
    from velocity_b200 import KLT, synth
    from velocity_b200.sequence import track_sequence

    frames, _ = synth.plane_sequence(7, h=270, w=480, seed=3, Z0=40.0)
    frames = np.stack(frames)
    pts = synth.harris_tracks(frames[0], 300, border=20)
    lk = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    p2, v, err = track_sequence(frames, pts, fbt=1.0, chunk=3, **lk)
    assert p2.shape == (6, 300, 2) and v.shape == (6, 300)
    for k in range(6):
        q2, qv, qerr = KLT.cv2calcOpticalFlowPyrLK(frames[k], frames[k + 1], pts, None, fbt=1.0, **lk)
        assert np.array_equal(p2[k], q2) and np.array_equal(v[k], qv) and np.array_equal(err[k], qerr.ravel())
    # the same pipeline replayed as one CUDA graph (what bench.py's e2e number times): identical results, also on re-use
    from velocity_b200.sequence import SequenceTracker

    tr = SequenceTracker(270, 480, 300, chunk=3, fbt=1.0, **lk)
    fh, ph = cuda.from_numpy(frames).pin_memory(), cuda.from_numpy(pts).pin_memory()
    op = cuda.empty((6, 300, 2), dtype=cuda.float32).pin_memory()
    os_ = cuda.empty((6, 300), dtype=cuda.uint8).pin_memory()
    oe = cuda.empty((6, 300), dtype=cuda.float32).pin_memory()
    for rep in range(3):
        op.zero_(); os_.zero_(); oe.zero_()
        tr.run(fh, ph, op, os_, oe, graph=True)
        assert np.array_equal(op.numpy(), p2) and np.array_equal(os_.numpy() != 0, v) and np.array_equal(oe.numpy(), err), rep


def test_bgr2gray_bit_exact(cuda):
    from oracle import cv_oracle as O
    from velocity_b200 import ingest

    g = golden("ingest_bgr")
    assert np.array_equal(ingest.bgr2gray(g["bgr"]), g["gray"])
    rng = np.random.default_rng(2)
    batch = rng.integers(0, 256, (3, 270, 480, 3), dtype=np.uint8)
    out = ingest.bgr2gray(cuda.from_numpy(batch).cuda()).cpu().numpy()
    for k in range(3):
        assert np.array_equal(out[k], O.bgr2gray(batch[k]))
