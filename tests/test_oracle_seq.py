"""CPU checks of the sequence oracle (oracle/seq_oracle.py): its vectorised pieces against the pinned per-function
oracles, and the recovered speed of a small synthetic approach sequence."""
import numpy as np

from oracle import seq_oracle as Q
from oracle import sfm_oracle as S
from util import golden
from velocity_b200 import synth


def test_vectorised_triangulation_matches_pinned_form():
    g = golden("triangulate")
    A, U = g["A"], g["U"]
    ref = S.triangulate_rays(A, U)
    np.testing.assert_allclose(Q.triangulate_rays_vec(A, U), ref, rtol=0, atol=1e-9)
    np.testing.assert_allclose(ref, g["cnv"], rtol=0, atol=1e-9)     # and the reference's own fcnNvintercept output


def small_scene(n=10, h=240, w=320, npts=96, z_start=60.0, seed=5):
    K = np.array([[400.0, 0, 0], [0, 400.0, 0], [w / 2 + 0.5, h / 2 + 0.5, 1]])
    frames, Z = synth.approach_sequence(n, h=h, w=w, seed=seed, z_start=z_start, K_row=K)
    p0 = synth.approach_tracks(frames[0], npts, Z[0] / Z[-1], K_row=K, border=24)
    p3 = np.concatenate([(p0 - K[2, 0:2]) / K[0, 0] * Z[0], np.full((npts, 1), Z[0])], 1).astype(float)
    times = np.arange(n) / 29.97
    return K, frames, p0, p3, times


def test_sequence_oracle_recovers_speed():
    K, frames, p0, p3, times = small_scene()
    out = Q.run_sequence(K, frames, p0, p3, times, ba_iters=3, winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))
    assert out["alive"][-1].mean() > 0.9
    assert out["alive"][0].all() and (np.diff(out["alive"].astype(int), axis=0) <= 0).all()     # tracks only ever die
    assert np.isnan(out["tracks"][~out["alive"]]).all()
    sp = out["S"][1:, 8]
    assert abs(sp.mean() - 40.0) / 40.0 < 0.02, sp
    assert np.isfinite(out["speed_ba"][1:]).all()
    assert out["S"][0, 2] == len(p0) and out["S"][-1, 2] == out["alive"][-1].sum()
    # the bundle adjustment lowers the reprojection rms it minimises
    assert out["hist"][-1][0] <= out["hist"][0][0]
