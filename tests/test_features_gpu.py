"""Parity of K9 (Harris response + goodFeaturesToTrack selection, csrc/features.cu) against the CPU oracle and the
cv2-generated golden vectors: same response bits, same corners, same order."""
import numpy as np
import pytest

from util import golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from velocity_b200 import _lib

    _lib.lib()
    return torch


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_harris_response_equals_oracle_bit_for_bit(cuda, tag):
    from oracle import gftt_oracle as G
    from velocity_b200 import features

    g = golden("gftt")
    im = g["im_" + tag]
    _, _, resp = features.harris_corners_device(im, 100, 0.01, want_response=True)
    R = resp.cpu().numpy()
    assert np.array_equal(R, G.harris_response(im))          # CUDA == oracle, every pixel
    ref, tail = g["resp_" + tag], (im.shape[1] // 16) * 16   # == cv2 except its last-row SIMD tail (oracle header)
    assert np.array_equal(R[:-1], ref[:-1]) and np.array_equal(R[-1, :tail], ref[-1, :tail])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("n,q", [(1000, 0.01), (64, 0.05), (4096, 0.001)])
def test_good_features_equal_cv2_golden(cuda, tag, n, q):
    from velocity_b200 import features

    g = golden("gftt")
    ref = g["xy_%s_%d" % (tag, n)]
    out = features.goodFeaturesToTrack(g["im_" + tag], n, q, 0, blockSize=5, useHarrisDetector=True)
    assert out.dtype == np.float32 and out.shape == ref.shape
    assert np.array_equal(out, ref)


def test_good_features_full_frame_and_views(cuda):
    """1080p frame (BASELINE size): equal to the oracle; a CUDA ROI view (pitch != width, as vidExample.py:109 cuts it)
    gives the same corners as the contiguous copy; unsupported configurations are refused, not approximated."""
    from oracle import gftt_oracle as G
    from velocity_b200 import features, synth

    im = synth.texture(1080, 1920, 77)
    out = features.goodFeaturesToTrack(im, 4096, 0.001, 0, blockSize=5, useHarrisDetector=True)
    assert np.array_equal(out.reshape(-1, 2), G.good_features_to_track(im, 4096, 0.001))
    d = cuda.from_numpy(im).cuda()
    roi = d[101:901, 333:1500]
    a = features.goodFeaturesToTrack(roi, 1000, 0.01, 0, blockSize=5, useHarrisDetector=True)
    b = features.goodFeaturesToTrack(np.ascontiguousarray(im[101:901, 333:1500]), 1000, 0.01, 0, blockSize=5, useHarrisDetector=True)
    assert np.array_equal(a, b) and len(a) == 1000
    assert np.array_equal(a.reshape(-1, 2), G.good_features_to_track(im[101:901, 333:1500], 1000, 0.01))
    flat = np.full((64, 64), 9, np.uint8)
    assert features.goodFeaturesToTrack(flat, 10, 0.01, 0, blockSize=5, useHarrisDetector=True) is None
    with pytest.raises(NotImplementedError):
        features.goodFeaturesToTrack(im, 10, 0.01, 5, blockSize=5, useHarrisDetector=True)
    with pytest.raises(NotImplementedError):
        features.goodFeaturesToTrack(im, 10, 0.01, 0, blockSize=3, useHarrisDetector=True)


def test_good_features_against_cv2_itself(cuda):
    """The detector against opencv-python run on THIS host (when importable), at the BASELINE frame size and on the
    plate ROI geometry of vidExample.py:109 (width not a multiple of 16): same corners, same order."""
    cv2 = pytest.importorskip("cv2")
    from velocity_b200 import features, synth

    frames, _ = synth.plane_sequence(1, h=1080, w=1920, seed=9, Z0=40.0)
    for im in (frames[0], np.ascontiguousarray(frames[0][40:1041, 260:1661])):
        for n, q in ((1000, 0.01), (4096, 0.001)):
            ref = cv2.goodFeaturesToTrack(im, n, q, 0, blockSize=5, useHarrisDetector=True)
            out = features.goodFeaturesToTrack(im, n, q, 0, blockSize=5, useHarrisDetector=True)
            assert out.shape == ref.shape and np.array_equal(out, ref), (im.shape, n, q)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("crit", ["ref", "loose", "count"])
def test_corner_subpix_equals_cv2_golden(cuda, tag, crit):
    from velocity_b200 import features

    crits = {"ref": (3, 100, 0.001), "loose": (3, 20, 0.03), "count": (1, 7, 0.0)}
    g = golden("subpix")
    p = g["p_" + tag]
    out = features.cornerSubPix(g["im_" + tag], p.reshape(-1, 1, 2), (5, 5), (-1, -1), crits[crit])
    assert out.shape == (len(p), 1, 2) and out.dtype == np.float32
    assert np.array_equal(out.reshape(-1, 2), g["q_%s_%s" % (tag, crit)])


def test_corner_subpix_full_frame_against_cv2_itself(cuda):
    cv2 = pytest.importorskip("cv2")
    from velocity_b200 import features, synth

    frames, _ = synth.plane_sequence(1, h=1080, w=1920, seed=9, Z0=40.0)
    im = frames[0]
    p = cv2.goodFeaturesToTrack(im, 1000, 0.01, 0, blockSize=5, useHarrisDetector=True)
    ref = cv2.cornerSubPix(im, p.copy(), (5, 5), (-1, -1), (3, 100, 0.001))
    assert np.array_equal(features.cornerSubPix(im, p, (5, 5), (-1, -1), (3, 100, 0.001)), ref)
    with pytest.raises(NotImplementedError):
        features.cornerSubPix(im, p, (5, 5), (2, 2), (3, 100, 0.001))
