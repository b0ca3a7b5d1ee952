"""K8 building blocks at the C5 global-BA size (8 GPUs x 300 cameras: n = 6 * 2399 = 14,394) and just above the CTA count
(n = 9,664: 151 panels > 148 CTAs), checked against torch's FP64 GEMM / Cholesky on the same GPU.  These sizes take the
non-resident tile ownership (a CTA owns several diagonal tiles) and the barrier form of the backward substitution."""
import sys
import time

import torch

sys.path.insert(0, ".")
from velocity_b200 import _lib
from velocity_b200.device import ptr, stream_ptr

L = _lib.lib()
ok = True
g = torch.Generator(device="cuda").manual_seed(1)
for n in (int(a) for a in (sys.argv[1:] or ["9664", "14394"])):
    k = 3072
    E = torch.randn((n, k), dtype=torch.float64, device="cuda", generator=g) / 50.0
    S0 = torch.eye(n, dtype=torch.float64, device="cuda") * 3.0
    S = S0.clone()
    work = torch.empty((max(int(L.vel_syrk_lower_sub_workspace(n, k)), 16),), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(L.vel_syrk_lower_sub(ptr(E), k, n, k, ptr(S), n, ptr(work), work.numel(), stream_ptr()), "syrk")
    torch.cuda.synchronize(); t_syrk = time.perf_counter() - t0
    want = S0 - E @ E.T
    err = (torch.tril(S) - torch.tril(want)).abs().max().item() / want.abs().max().item()
    print("syrk m=%d k=%d: rel err %.2e  %.1f ms (%.1f TFLOP/s)" % (n, k, err, t_syrk * 1e3, n * (n + 1.0) * k / t_syrk / 1e12), flush=True)
    ok &= err < 1e-12
    del want
    # SPD system: S_spd = 3 I + E E^T
    A = (S0 + E @ E.T).contiguous()
    b = torch.randn((n,), dtype=torch.float64, device="cuda", generator=g)
    info = torch.zeros((1,), dtype=torch.int32, device="cuda")
    Aw, bw = A.clone(), b.clone()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(L.vel_spd_solve(ptr(Aw), n, n, ptr(bw), ptr(info), stream_ptr()), "spd_solve")
    torch.cuda.synchronize(); t_chol = time.perf_counter() - t0
    Lw = torch.linalg.cholesky(A)
    xw = torch.cholesky_solve(b[:, None], Lw)[:, 0]
    ex = (bw - xw).abs().max().item() / xw.abs().max().item()
    el = (torch.tril(Aw) - Lw).abs().max().item() / Lw.abs().max().item()
    print("chol n=%d: x rel err %.2e  L rel err %.2e  info %d  %.1f ms" % (n, ex, el, info.item(), t_chol * 1e3), flush=True)
    ok &= ex < 1e-9 and el < 1e-10 and info.item() == 0
    del A, Aw, Lw, E, S, S0
    torch.cuda.empty_cache()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
