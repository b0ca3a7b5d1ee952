"""Pins oracle/gftt_oracle.py (restatement of cv2.goodFeaturesToTrack as vidExample.py:110 calls it) to cv2's own
outputs stored in tests/golden/gftt.npz (made by tests/golden/make_golden.py with opencv-python 4.13.0).  CPU only."""
import numpy as np
import pytest

from oracle import gftt_oracle as G
from util import golden

CASES = [("a", 160), ("b", 203), ("c", 77)]


@pytest.mark.parametrize("tag,width", CASES)
def test_harris_response_bit_exact(tag, width):
    g = golden("gftt")
    im = g["im_" + tag]
    assert im.shape[1] == width
    R = G.harris_response(im)
    ref = g["resp_" + tag]
    # everything but the final (width mod 16) values of the LAST row is bit-identical (see the oracle's header)
    tail = (width // 16) * 16
    assert np.array_equal(R[:-1], ref[:-1])
    assert np.array_equal(R[-1, :tail], ref[-1, :tail])
    assert np.allclose(R[-1, tail:], ref[-1, tail:], rtol=1e-5, atol=1e-12)


@pytest.mark.parametrize("tag,width", CASES)
@pytest.mark.parametrize("n,q", [(1000, 0.01), (64, 0.05), (4096, 0.001)])
def test_good_features_same_corners_same_order(tag, width, n, q):
    g = golden("gftt")
    ref = g["xy_%s_%d" % (tag, n)].reshape(-1, 2)
    mine = G.good_features_to_track(g["im_" + tag], n, q)
    assert mine.dtype == np.float32 and np.array_equal(mine, ref)
    assert len(ref) > 10


SUBPIX_CRIT = {"ref": (3, 100, 0.001), "loose": (3, 20, 0.03), "count": (1, 7, 0.0)}


@pytest.mark.parametrize("tag", ["a", "b", "c"])
@pytest.mark.parametrize("crit", list(SUBPIX_CRIT))
def test_corner_subpix_bit_exact(tag, crit):
    """oracle/velocity_oracle.c::orc_corner_subpix_u8 against cv2.cornerSubPix (golden subpix.npz): every refined corner
    bit-identical, including those whose 13x13 sampling window hangs over the frame border."""
    from oracle import cv_oracle as O

    g = golden("subpix")
    out = O.cornerSubPix(g["im_" + tag], g["p_" + tag], (5, 5), (-1, -1), SUBPIX_CRIT[crit])
    assert out.dtype == np.float32 and np.array_equal(out, g["q_%s_%s" % (tag, crit)])
    assert (out != g["p_" + tag]).any(1).mean() > 0.2      # the rest diverged and were reset to the input, as cv2 does
