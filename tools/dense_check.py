"""Checks and times the hand-written FP64 building blocks of K8 (csrc/dense_f64.cu) against numpy on the GPU box:
vel_syrk_lower_sub (DMMA SYRK with turnstile split-K) and vel_spd_solve (cooperative blocked Cholesky + solves)."""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from velocity_b200 import _lib
from velocity_b200.device import ptr, stream_ptr

L = _lib.lib()
rng = np.random.default_rng(0)


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def syrk_case(m, k, reps=1):
    ld = (k + 31) // 32 * 32
    E = np.zeros((m, ld))
    E[:, :k] = rng.normal(0, 1, (m, k))
    S0 = rng.normal(0, 1, (m, m))
    dE, dS = torch.from_numpy(E).cuda(), torch.from_numpy(S0).cuda()
    nbytes = L.vel_syrk_lower_sub_workspace(m, k)
    work = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")

    def run():
        _lib.check(L.vel_syrk_lower_sub(ptr(dE), ld, m, k, ptr(dS), m, ptr(work), work.numel(), stream_ptr()), "syrk")
    run()
    torch.cuda.synchronize()
    got = dS.cpu().numpy()
    want = S0 - E @ E.T
    il = np.tril_indices(m)
    err = np.abs(got[il] - want[il]).max() / max(np.abs(want[il]).max(), 1e-300)
    # second call on the same S must be bit-reproducible relative to a fresh run
    dS2 = torch.from_numpy(S0).cuda()
    _lib.check(L.vel_syrk_lower_sub(ptr(dE), ld, m, k, ptr(dS2), m, ptr(work), work.numel(), stream_ptr()), "syrk")
    same = bool(torch.equal(torch.tril(dS2), torch.tril(dS)))
    ms = timed(run, reps) if reps > 1 else float("nan")
    fl = float(m) * (m + 1) * k
    print("syrk m=%5d k=%6d: rel err %.2e  reproducible %s  %.3f ms  %.1f TFLOP/s" % (m, k, err, same, ms, fl / ms / 1e9 if reps > 1 else 0))
    return err < 1e-12 and same


def chol_case(n, reps=1):
    A = rng.normal(0, 1, (n, n + 8))
    S = A @ A.T + np.eye(n)
    b = rng.normal(0, 1, n)
    dS0 = torch.from_numpy(S).cuda()
    db0 = torch.from_numpy(b).cuda()
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    dS, db = dS0.clone(), db0.clone()

    def run():
        dS.copy_(dS0); db.copy_(db0)
        _lib.check(L.vel_spd_solve(ptr(dS), n, n, ptr(db), ptr(info), stream_ptr()), "spd_solve")
    run()
    torch.cuda.synchronize()
    x = db.cpu().numpy()
    Lg = np.tril(dS.cpu().numpy())
    want = np.linalg.solve(S, b)
    Lw = np.linalg.cholesky(S)
    ex = np.abs(x - want).max() / np.abs(want).max()
    el = np.abs(Lg - Lw).max() / np.abs(Lw).max()
    ms = timed(run, reps) if reps > 1 else float("nan")
    print("chol n=%5d: x rel err %.2e  L rel err %.2e  info %d  %.3f ms (incl. two small copies)" % (n, ex, el, info.item(), ms))
    # not positive definite -> info = 1
    bad = S.copy(); bad[n // 2, n // 2] = -1.0
    dB = torch.from_numpy(bad).cuda()
    _lib.check(L.vel_spd_solve(ptr(dB), n, n, ptr(db), ptr(info), stream_ptr()), "spd_solve")
    flagged = info.item() == 1
    return ex < 1e-9 and el < 1e-10 and flagged


ok = True
for m, k in [(6, 16), (30, 48), (130, 100), (257, 1000), (594, 1536), (1794, 12288)]:
    ok &= syrk_case(m, k, reps=5 if m >= 594 else 1)
for n in [1, 6, 30, 64, 65, 200, 594, 1794]:
    ok &= chol_case(n, reps=5 if n >= 594 else 1)
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
