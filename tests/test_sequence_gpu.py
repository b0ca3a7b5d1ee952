"""GPU parity of the connected sequence pipeline (velocity_b200.sfm.SfmSequence, C-ABI vel_klt_sequence / vel_seq_*)
against oracle/seq_oracle.py, and the C3-size properties (recovered speed within 1 %, BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import seq_oracle as Q
from test_oracle_seq import small_scene
from velocity_b200 import synth

pytestmark = pytest.mark.gpu
LK = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))


def scene_with_dying_tracks():
    K, frames, p0, p3, times = small_scene(n=9, npts=80)
    h, w = frames[0].shape
    rng = np.random.default_rng(3)
    # tracks that leave the frame as the plane grows, flat-texture points, and seeds outside the frame
    extra = np.float32([[4.2, 5.1], [w - 6.5, h - 7.25], [w - 9.0, 30.0], [-30.0, 10.0], [w + 50.0, h / 2], [3.0, h - 4.0]])
    extra = np.concatenate([extra, np.stack([rng.uniform(8, w - 8, 10), rng.uniform(8, 14, 10)], 1).astype(np.float32)])
    p0 = np.concatenate([p0, extra]).astype(np.float32)
    Z0 = p3[0, 2]
    p3 = np.concatenate([(p0 - K[2, 0:2]) / K[0, 0] * Z0, np.full((len(p0), 1), Z0)], 1).astype(float)
    return K, frames, p0, p3, times


def run_gpu(K, frames, p0, p3, times, host_frames=False, chunk=4, ba_iters=4):
    from velocity_b200.sfm import SfmSequence

    n, h, w = frames.shape
    seq = SfmSequence(K, h, w, n, len(p0), fbt=1.0, ba_iters=ba_iters, chunk=chunk, **LK)
    fr = torch.from_numpy(frames)
    fr = fr.pin_memory() if host_frames else fr.cuda()
    hist = seq.run(fr, p0, p3, times)
    torch.cuda.synchronize()
    return seq, hist


def test_sequence_matches_oracle():
    K, frames, p0, p3, times = scene_with_dying_tracks()
    ref = Q.run_sequence(K, frames, p0, p3, times, ba_iters=4, **LK)
    seq, hist = run_gpu(K, frames, p0, p3, times)
    alive = seq.alive.cpu().numpy() != 0
    tracks = seq.tracks.cpu().numpy()
    assert np.array_equal(alive, ref["alive"])                                   # masks bit-exact
    assert 0 < alive[-1].sum() < alive[0].sum()                                  # the case exercises dying tracks
    assert np.array_equal(tracks[alive], ref["tracks"][alive])                   # propagated points bit-exact
    B, S = seq.B.cpu().numpy(), seq.S.cpu().numpy()
    np.testing.assert_allclose(B[1:, 3:6], ref["B"][1:, 3:6], rtol=1e-5, atol=1e-6)   # fcnNLS_t, float32 outputs
    np.testing.assert_allclose(B[:, 0:3], ref["B"][:, 0:3], rtol=1e-5, atol=1e-6)
    assert np.array_equal(S[:, 0], ref["S"][:, 0]) and np.array_equal(S[:, 2], ref["S"][:, 2])
    np.testing.assert_allclose(S[1:, 3], ref["S"][1:, 3], rtol=1e-4)                   # rms residual
    np.testing.assert_allclose(S[1:, 4:9], ref["S"][1:, 4:9], rtol=2e-4, atol=1e-6)    # dt, time, dr, distance, speed
    assert np.isnan(S[0, 4]) and np.isnan(S[0, 8]) and S[0, 6] == 0
    proj = seq.proj.cpu().numpy()
    assert np.array_equal(np.isnan(proj), np.isnan(ref["proj"]))
    np.testing.assert_allclose(proj[alive], ref["proj"][alive], rtol=0, atol=2e-3)
    assert (seq.iters.cpu().numpy()[1:] > 0).all() and not ref["capped"]
    # full-length selection, triangulation, bundle adjustment
    idx, pw, cw = seq.points_ba()
    assert np.array_equal(idx.cpu().numpy(), ref["idx"])
    np.testing.assert_allclose(seq.C0[:len(ref["idx"])].cpu().numpy(), ref["C0"], rtol=1e-6, atol=1e-6)
    assert len(hist) == len(ref["hist"])
    np.testing.assert_allclose([h[0] for h in hist], [h[0] for h in ref["hist"]], rtol=1e-6)
    np.testing.assert_allclose(cw.cpu().numpy(), ref["cw"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(pw.cpu().numpy(), ref["pw"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(seq.S_ba.cpu().numpy()[1:, 8], ref["speed_ba"][1:], rtol=1e-3)
    # the reference's P array
    P = seq.export_P().cpu().numpy()
    assert P.shape == (5, len(p0), len(frames))
    want = np.full_like(P, np.nan)
    want[0:2] = ref["tracks"].transpose(2, 1, 0)
    want[4] = np.where(ref["alive"].T, np.arange(len(frames), dtype=np.float32)[None], np.nan)
    assert np.array_equal(np.isnan(P[[0, 1, 4]]), np.isnan(want[[0, 1, 4]]))
    assert np.array_equal(np.nan_to_num(P[[0, 1, 4]]), np.nan_to_num(want[[0, 1, 4]]))
    np.testing.assert_allclose(np.nan_to_num(P[2:4]), np.nan_to_num(ref["proj"].transpose(2, 1, 0)), atol=2e-3)


def test_host_frames_path_is_identical():
    K, frames, p0, p3, times = scene_with_dying_tracks()
    a, _ = run_gpu(K, frames, p0, p3, times, host_frames=False)
    for chunk in (1, 4, 9):
        b, _ = run_gpu(K, frames, p0, p3, times, host_frames=True, chunk=chunk)
        assert torch.equal(a.alive, b.alive) and torch.equal(a.tracks, b.tracks)
        assert torch.equal(a.S.nan_to_num(), b.S.nan_to_num()) and torch.equal(a.ba.x, b.ba.x)
        assert b.h2d_bytes == frames.size


def test_prefetched_uploads_pipeline_across_sequences():
    """SfmSequence.prefetch: the next sequence's upload is started before the current one runs (two device frame buffers);
    three different recordings pushed through the pipeline give exactly what each gives on its own."""
    from velocity_b200.sfm import SfmSequence

    K, frames, p0, p3, times = scene_with_dying_tracks()
    n, h, w = frames.shape
    rng = np.random.default_rng(5)
    recs = [frames, np.clip(frames.astype(np.int16) + rng.integers(-3, 4, frames.shape), 0, 255).astype(np.uint8), frames[:, ::-1].copy()]
    want = []
    for fr in recs:
        a, _ = run_gpu(K, fr, p0, p3, times, host_frames=False)
        want.append((a.tracks.clone(), a.alive.clone(), a.S.clone()))
    seq = SfmSequence(K, h, w, n, len(p0), fbt=1.0, ba_iters=4, chunk=4, **LK)
    host = [torch.from_numpy(fr).pin_memory() for fr in recs]
    hp0, hp3, htm = torch.from_numpy(p0).pin_memory(), torch.from_numpy(p3).pin_memory(), torch.from_numpy(np.asarray(times, np.float32)).pin_memory()
    seq.prefetch(host[0], hp0, hp3, htm)
    for k in range(3):
        if k + 1 < 3:
            seq.prefetch(host[k + 1], hp0, hp3, htm)
        seq.run(host[k], hp0, hp3, htm)
        torch.cuda.synchronize()
        assert torch.equal(seq.tracks, want[k][0]) and torch.equal(seq.alive, want[k][1]), k
        assert torch.equal(seq.S.nan_to_num(), want[k][2].nan_to_num()), k
        assert seq.h2d_bytes == frames.size
    # a run() whose frames were not announced uploads by itself; a third unconsumed prefetch is refused
    seq.run(host[1], p0, p3, times)
    torch.cuda.synchronize()
    assert torch.equal(seq.tracks, want[1][0])
    seq.prefetch(host[0]); seq.prefetch(host[2])
    with pytest.raises(RuntimeError):
        seq.prefetch(host[1])


def test_sequence_results_to_host():
    from velocity_b200.sfm import SfmSequence

    K, frames, p0, p3, times = scene_with_dying_tracks()
    n, h, w = frames.shape
    seq = SfmSequence(K, h, w, n, len(p0), ba_iters=2, **LK)
    out = dict(S=torch.empty((n, 9)).pin_memory(), S_ba=torch.empty((n, 9)).pin_memory(), B=torch.empty((n, 14)).pin_memory(),
               P=torch.empty((5, len(p0), n)).pin_memory())
    seq.run(torch.from_numpy(frames).pin_memory(), p0, p3, times, out=out)
    assert torch.equal(out["S"].nan_to_num(), seq.S.cpu().nan_to_num()) and torch.equal(out["B"], seq.B.cpu())
    assert seq.d2h_bytes == 4 * (2 * n * 9 + n * 14 + 5 * len(p0) * n)
    # running the same object again gives the same answer (buffers are reused)
    S1 = out["S"].clone()
    seq.run(torch.from_numpy(frames).pin_memory(), p0, p3, times, out=out)
    assert torch.equal(S1.nan_to_num(), out["S"].nan_to_num())
    # sync=False: the small tables are there when the call returns, the P export lands behind the NEXT sequence's launches
    P1 = out["P"].clone()
    out2 = {k: torch.full_like(v, -7.0).pin_memory() for k, v in out.items()}
    pinned = torch.from_numpy(frames).pin_memory()
    for _ in range(2):
        seq.run(pinned, p0, p3, times, out=out2, sync=False)
        assert torch.equal(S1.nan_to_num(), out2["S"].nan_to_num())
    seq.wait_results()
    assert torch.equal(P1.nan_to_num(), out2["P"].nan_to_num())


def test_c3_full_size_speed_within_one_percent():
    """BASELINE configs[2]: 300 x 1080p frames, 4096 tracks propagated through the whole sequence, fcnNLS_t per frame,
    triangulation + 10 bundle-adjustment iterations; the 40 km/h of the generator is recovered within 1 %."""
    from oracle import klt_oracle
    from velocity_b200.sfm import SfmSequence

    n, npts = 300, 4096
    K = synth.K_1080P
    frames, Z = synth.approach_sequence(n, seed=2025, z_start=200.0)
    p0 = synth.approach_tracks(frames[0], npts, Z[0] / Z[-1])
    p3 = np.concatenate([(p0 - K[2, 0:2]) / K[0, 0] * Z[0], np.full((npts, 1), Z[0])], 1).astype(float)
    times = np.arange(n) / 29.97
    seq = SfmSequence(K, 1080, 1920, n, npts, **LK)
    hist = seq.run(torch.from_numpy(frames).cuda(), p0, p3, times)
    S, S_ba = seq.S.cpu().numpy(), seq.S_ba.cpu().numpy()
    alive = seq.alive.cpu().numpy() != 0
    assert alive[-1].mean() > 0.99
    assert abs(S[1:, 8].mean() - 40.0) / 40.0 < 0.01, S[1:, 8].mean()
    assert np.abs(S[1:, 8] - 40.0).max() / 40.0 < 0.03                     # every single frame within 3 %
    assert abs(S[-1, 7] - (Z[0] - Z[-1])) / (Z[0] - Z[-1]) < 0.01           # distance travelled
    assert len(hist) == 10 and hist[-1][0] <= hist[0][0]
    assert abs(S_ba[1:, 8].mean() - 40.0) / 40.0 < 0.01, S_ba[1:, 8].mean()
    # the first propagated frames against the CPU oracle, bit for bit, at full size
    tr = seq.tracks[:3].cpu().numpy()
    p = p0
    for i in (1, 2):
        p2, v, _ = klt_oracle.lk_forward_backward(frames[i - 1], frames[i], p, fbt=1.0, **LK)
        assert np.array_equal(v, alive[i]) and np.array_equal(p2[v], tr[i][v])
        p = p2


def test_in_kernel_frame_loop_equals_per_pair_launches(monkeypatch):
    """vel_klt_sequence has four forms: the 15x15 sequence kernel (frame loop inside, ONE template per frame serving the
    backward pass of pair j-1 and the forward pass of pair j), the same kernel with the search neighbourhood staged in shared memory
    by TMA (VEL_LK_SEQ=tma, the A/B variant), the batch kernel with the frame loop inside (VEL_LK_SEQ=twice) and one K2 launch per
    pair (any window; VEL_LK_SEQ=pairs).  Same tracks, same masks, same errors, bit for bit."""
    K, frames, p0, p3, times = scene_with_dying_tracks()
    a, _ = run_gpu(K, frames, p0, p3, times)
    for mode in ("pairs", "twice", "tma"):   # one K2 launch per pair / frame loop in the batch kernel (template per pass) / TMA boxes
        monkeypatch.setenv("VEL_LK_SEQ", mode)
        b, _ = run_gpu(K, frames, p0, p3, times)
        assert torch.equal(a.alive, b.alive), mode
        al = a.alive != 0
        assert torch.equal(a.tracks[al], b.tracks[al]), mode
        assert torch.equal(a.err[al[:-1]], b.err[al[:-1]]), mode        # err[k] belongs to pair k, defined where the track was alive in frame k
        assert (a.tracks[~al] == -1.0e5).all() and (b.tracks[~al] == -1.0e5).all()
        assert torch.equal(a.S.nan_to_num(), b.S.nan_to_num()), mode
    monkeypatch.setenv("VEL_LK_SEQ", "pairs")
    # a 21x21 window takes the per-pair path by construction and must agree with the oracle too
    from velocity_b200.sfm import SfmSequence

    lk = dict(winSize=(21, 21), maxLevel=2, criteria=(3, 10, 0.03))
    n, h, w = frames.shape
    seq = SfmSequence(K, h, w, n, len(p0), fbt=0.5, ba_iters=1, **lk)
    seq.run(torch.from_numpy(frames).cuda(), p0, p3, times, bundle=False)
    tr, al = Q.track_sequence(frames, p0, fbt=0.5, **lk)
    assert np.array_equal(seq.alive.cpu().numpy() != 0, al) and np.array_equal(seq.tracks.cpu().numpy()[al], tr[al])
