"""K8 building blocks (csrc/dense_f64.cu) against numpy: the FP64 tensor-core SYRK with deterministic split-K and the
cooperative blocked Cholesky + solve -- the hand-written replacements for cuBLAS DSYRK / cuSOLVER DPOTRF+DPOTRS behind
`inv(JJ^T + I) @ ...` (utils/NLS.py:236).  Also the bundle adjustment with VEL_BA_SOLVER=native end to end."""
import contextlib
import io

import numpy as np
import pytest
import torch

from util import ba_c3_inputs, golden

pytestmark = pytest.mark.gpu


def _lib():
    from velocity_b200 import _lib as L

    return L


@pytest.mark.parametrize("m,k", [(6, 16), (30, 48), (130, 100), (257, 1000), (594, 1536), (1794, 2048)])
def test_syrk_lower_sub(m, k):
    from velocity_b200.device import ptr, stream_ptr

    L = _lib()
    rng = np.random.default_rng(m)
    ld = (k + 31) // 32 * 32
    E = np.zeros((m, ld))
    E[:, :k] = rng.normal(0, 1, (m, k))
    S0 = rng.normal(0, 1, (m, m))
    dE = torch.from_numpy(E).cuda()
    work = torch.empty(max(L.lib().vel_syrk_lower_sub_workspace(m, k), 16), dtype=torch.uint8, device="cuda")
    outs = []
    for _ in range(2):
        dS = torch.from_numpy(S0).cuda()
        L.check(L.lib().vel_syrk_lower_sub(ptr(dE), ld, m, k, ptr(dS), m, ptr(work), work.numel(), stream_ptr()), "syrk")
        outs.append(dS)
    got, want = outs[0].cpu().numpy(), S0 - E @ E.T
    il, iu = np.tril_indices(m), np.triu_indices(m, 1)
    assert np.abs(got[il] - want[il]).max() <= 1e-12 * np.abs(want[il]).max()
    assert np.array_equal(got[iu], S0[iu])                               # the strict upper triangle is never touched
    assert torch.equal(outs[0], outs[1])                                 # split-K partials are applied in a fixed order
    # argument checks: unpadded rows are refused, not silently misread
    with pytest.raises(RuntimeError):
        L.check(L.lib().vel_syrk_lower_sub(ptr(dE), ld - 2, m, ld, ptr(outs[0]), m, ptr(work), work.numel(), stream_ptr()), "syrk")


@pytest.mark.parametrize("n", [1, 6, 30, 64, 65, 200, 300, 594, 1794])
def test_spd_solve(n):
    from velocity_b200.device import ptr, stream_ptr

    L = _lib()
    rng = np.random.default_rng(n)
    A = rng.normal(0, 1, (n, n + 8))
    S = A @ A.T + np.eye(n)
    b = rng.normal(0, 1, n)
    dS, db = torch.from_numpy(S).cuda(), torch.from_numpy(b).cuda()
    info = torch.full((1,), 7, dtype=torch.int32, device="cuda")
    L.check(L.lib().vel_spd_solve(ptr(dS), n, n, ptr(db), ptr(info), stream_ptr()), "spd_solve")
    x, Lg = db.cpu().numpy(), np.tril(dS.cpu().numpy())
    want, Lw = np.linalg.solve(S, b), np.linalg.cholesky(S)
    assert info.item() == 0
    assert np.abs(x - want).max() <= 1e-10 * np.abs(want).max()
    assert np.abs(Lg - Lw).max() <= 1e-12 * np.abs(Lw).max()
    assert np.array_equal(np.triu(dS.cpu().numpy(), 1), np.triu(S, 1))   # upper triangle untouched
    # a matrix that is not positive definite is reported, not silently "solved"
    bad = S.copy()
    bad[n // 2, n // 2] = -1.0
    dB = torch.from_numpy(bad).cuda()
    L.check(L.lib().vel_spd_solve(ptr(dB), n, n, ptr(db), ptr(info), stream_ptr()), "spd_solve")
    assert info.item() == 1


@pytest.mark.parametrize("grid", ["32", "8"])
def test_spd_solve_on_fewer_ctas_than_the_device_has(monkeypatch, grid):
    """The task-graph factorisation on a grid a smaller device would give it (VEL_CHOL_GRID): 32 CTAs for 29 panels -- the panel
    CTAs keep their two critical tiles resident AND share in the other tiles (an ownership test that only held when they did not
    share was found here) -- and 8 CTAs, where a CTA owns several diagonal tiles and nothing stays resident."""
    monkeypatch.setenv("VEL_CHOL_GRID", grid)
    test_spd_solve(1794)
    test_spd_solve(300)


@pytest.mark.parametrize("name", ["ba_small", "ba_medium", "ba_256x10", "ba_512x20"])
def test_bundle_adjustment_with_native_solver(monkeypatch, name):
    """fcnNLS_batch with this library's own SYRK + Cholesky (VEL_BA_SOLVER=native) against the reference's golden results,
    and bit-identical to itself across runs (the vendor path uses a split-K reduction kernel too; both are deterministic)."""
    from velocity_b200 import NLS

    monkeypatch.setenv("VEL_BA_SOLVER", "native")
    g = golden(name)
    res = []
    for _ in range(2):
        with contextlib.redirect_stdout(io.StringIO()):
            res.append(NLS.fcnNLS_batch(g["K"], g["P"].copy(), g["pw0"].copy(), g["cw0"].copy()))
    cw, pw = res[0]
    assert np.allclose(cw, g["cw"], rtol=1e-6, atol=1e-7) and np.allclose(pw, g["pw"], rtol=1e-6, atol=1e-7)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    if name == "ba_small":
        with contextlib.redirect_stdout(io.StringIO()):
            cw2, pw2 = NLS.fcnNLS_batch2(g["K"], g["P"].copy(), g["pw0"].copy(), g["cw0"].copy())
        assert np.allclose(cw2, g["cw_b2"], rtol=1e-6, atol=1e-7) and np.allclose(pw2, g["pw_b2"], rtol=1e-6, atol=1e-7)


def test_native_solver_at_c3_size(monkeypatch):
    from velocity_b200 import NLS

    monkeypatch.setenv("VEL_BA_SOLVER", "native")
    g = golden("ba_c3_sparse")
    K, P, pw0, cw0 = ba_c3_inputs()
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        cw, pw = NLS.fcnNLS_batch(K, P, pw0, cw0)
    assert sum(1 for ln in buf.getvalue().splitlines() if "f=" in ln and "x=" in ln) == len(g["hist"])
    assert np.abs(cw - g["cw"]).max() < 1e-7 and np.abs(pw - g["pw"]).max() < 1e-6


def test_failed_cholesky_poisons_rms_delta(monkeypatch):
    """a reduced system that is not positive definite must not pass for a converged step: rms(delta) is NaN."""
    from velocity_b200 import NLS, _lib as L
    from velocity_b200.device import ptr, stream_ptr

    for mode in ("native", "vendor"):      # "vendor" only differs in a build with --vendor-solver (the A/B build)
        monkeypatch.setenv("VEL_BA_SOLVER", mode)
        g = golden("ba_small")
        from oracle import sfm_oracle as S

        z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
        ba = NLS.BundleAdjuster(g["K"], z, x, nt, nc)
        ba.accumulate()
        ba.U[:, 0] = -1e12                                               # wreck the first diagonal entry of every camera block
        ba.solve()
        assert np.isnan(ba.rms_delta.item()), mode


@pytest.mark.parametrize("blocked", ["", "0"])
def test_more_panels_than_ctas_takes_the_shared_ownership_forms(monkeypatch, blocked):
    """n = 9,664 (151 panels on 148 SMs; the 8-GPU global BA factors n = 14,394 on its owner): a CTA then owns several diagonal tiles,
    so the resident-tile form and the task-graph backward substitution step aside for the global-memory forms.  Both ways of
    factoring it: the BLOCKED form (groups of 16 panels by the task graph, the trailing matrix by the SYRK kernel; the default above
    48 panels) and, with VEL_CHOL_BLOCKED=0, the pure task graph.  Against torch's FP64 Cholesky on the same device; the SYRK at the
    same order; the queue-fed and the statically assigned SYRK agree bit for bit."""
    if blocked:
        monkeypatch.setenv("VEL_CHOL_BLOCKED", blocked)
    import os

    from velocity_b200.device import ptr, stream_ptr

    L = _lib()
    n, k = 9664, 1024
    g = torch.Generator(device="cuda").manual_seed(1)
    E = torch.randn((n, k), dtype=torch.float64, device="cuda", generator=g) / 30.0
    S0 = torch.eye(n, dtype=torch.float64, device="cuda") * 3.0
    work = torch.empty((max(int(L.lib().vel_syrk_lower_sub_workspace(n, k)), 16),), dtype=torch.uint8, device="cuda")
    outs = []
    for q in ("1", "0"):
        os.environ["VEL_SYRK_QUEUE"] = q
        S = S0.clone()
        L.check(L.lib().vel_syrk_lower_sub(ptr(E), k, n, k, ptr(S), n, ptr(work), work.numel(), stream_ptr()), "syrk")
        outs.append(torch.tril(S))
    os.environ.pop("VEL_SYRK_QUEUE")
    want = S0 - E @ E.T
    assert torch.equal(outs[0], outs[1])
    assert (outs[0] - torch.tril(want)).abs().max().item() <= 1e-12 * want.abs().max().item()
    A = (S0 + E @ E.T).contiguous()
    b = torch.randn((n,), dtype=torch.float64, device="cuda", generator=g)
    info = torch.zeros((1,), dtype=torch.int32, device="cuda")
    Aw, bw = A.clone(), b.clone()
    L.check(L.lib().vel_spd_solve(ptr(Aw), n, n, ptr(bw), ptr(info), stream_ptr()), "spd_solve")
    Lw = torch.linalg.cholesky(A)
    xw = torch.cholesky_solve(b[:, None], Lw)[:, 0]
    assert info.item() == 0
    assert (bw - xw).abs().max().item() <= 1e-10 * xw.abs().max().item()
    assert (torch.tril(Aw) - Lw).abs().max().item() <= 1e-11 * Lw.abs().max().item()


def test_blocked_form_reports_a_matrix_that_is_not_positive_definite():
    """The blocked factorisation (n = 3,264: 51 panels > 48) keeps the contract of vel_spd_solve: info = 1 when a pivot is not
    positive (here deep inside the fourth group of panels), 0 for the repaired matrix, whose solution then matches torch."""
    from velocity_b200.device import ptr, stream_ptr

    L = _lib()
    n = 3264
    g = torch.Generator(device="cuda").manual_seed(3)
    E = torch.randn((n, 256), dtype=torch.float64, device="cuda", generator=g) / 16.0
    A = (torch.eye(n, dtype=torch.float64, device="cuda") * 2.0 + E @ E.T).contiguous()
    b = torch.randn((n,), dtype=torch.float64, device="cuda", generator=g)
    info = torch.zeros((1,), dtype=torch.int32, device="cuda")
    bad = A.clone()
    bad[3100, 3100] = -5.0
    bw = b.clone()
    L.check(L.lib().vel_spd_solve(ptr(bad), n, n, ptr(bw), ptr(info), stream_ptr()), "spd_solve")
    assert info.item() == 1
    Aw, bw = A.clone(), b.clone()
    L.check(L.lib().vel_spd_solve(ptr(Aw), n, n, ptr(bw), ptr(info), stream_ptr()), "spd_solve")
    assert info.item() == 0
    xw = torch.cholesky_solve(b[:, None], torch.linalg.cholesky(A))[:, 0]
    assert (bw - xw).abs().max().item() <= 1e-11 * xw.abs().max().item()
