"""Drop-in for the reference's utils/MSV.py.

    fcn2vintercept  utils/MSV.py:98-142   -> K6 vel_triangulate_2v
    fcnNvintercept  utils/MSV.py:146-175  -> K6 vel_triangulate_nv
    fcnMSV1_t       utils/MSV.py:8-49     -> vel_msv1_t (whole LM loop in one persistent kernel)
    fcnMSV2_t       utils/MSV.py:52-94    -> refused: the reference implementation cannot run
                                             (its zero Jacobian blocks become -zhat/dx at :84 and
                                             np.linalg.inv raises LinAlgError for every input).
"""
import numpy as np
import torch

from . import _lib
from .common import pixel2uvec
from .device import ptr, require_cuda, stream_ptr


def _dev64(a):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64))).cuda()


def _triangulate(A, U, which):
    require_cuda()
    on_dev = isinstance(U, torch.Tensor) and U.is_cuda
    dA = A if on_dev else _dev64(A)
    dU = U if on_dev else _dev64(U)
    _, nf, nv = dU.shape
    out = torch.empty((nv, 3), dtype=torch.float64, device=dU.device)
    fn = getattr(_lib.lib(), which)
    _lib.check(fn(ptr(dA.contiguous()), ptr(dU.contiguous()), nf, nv, ptr(out), stream_ptr()), which)
    return out if on_dev else out.cpu().numpy()


def fcn2vintercept(A, U):
    """A [nf,3] ray origins, U [3,nf,nv] unit rays -> tie points [nv,3] (mean of pairwise closest approaches)."""
    return _triangulate(A, U, "vel_triangulate_2v")


def fcnNvintercept(A, U):
    """N-ray least-squares intersection -> [nv,3]."""
    return _triangulate(A, U, "vel_triangulate_nv")


def fcnMSV1_t(K, P, B, vg, ii):
    """Solves the last camera's translation with re-triangulation inside the LM loop.
    Returns (x float32 [3], b0 float64 [ng,3]) like the reference."""
    require_cuda()
    nf = ii + 1
    ng = int(vg.sum())
    U = np.zeros((3, nf, ng))
    for j in range(nf):
        U[:, j] = pixel2uvec(K, P[0:2, vg, j].T).T
    u0 = B[0, 0:3] - B[:nf, 0:3]
    x0 = np.array([0, 0, 1]) - u0[nf - 2]
    z = np.ascontiguousarray(P[0:2, vg, ii].T, dtype=np.float64)  # [ng, 2] == ravel("F") pairs
    max_iter = 1000
    dK, dA, dU, dz, dx0 = _dev64(K), _dev64(u0[:-1]), _dev64(U), _dev64(z), _dev64(x0)
    x = torch.empty(3, dtype=torch.float64, device="cuda")
    b0 = torch.empty((ng, 3), dtype=torch.float64, device="cuda")
    iters = torch.empty(1, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().vel_msv1_t(ptr(dK), ptr(dA), ptr(dU), nf, ng, ptr(dz), ptr(dx0), max_iter, ptr(x), ptr(b0),
                                     ptr(iters), stream_ptr()), "vel_msv1_t")
    # the reference warns when the LAST allowed iteration was reached, converged or not (utils/MSV.py:43)
    it = int(iters.item())
    if it < 0 or it == max_iter:
        print("WARNING: fcnMSV1_t() reaching max iterations!")
    return x.cpu().numpy().astype(np.float32), b0.cpu().numpy()


def fcnMSV2_t(K, P, B, vg, i):
    raise NotImplementedError(
        "fcnMSV2_t is not runnable in the reference either: `JT = (JT - zhat) / dx` (utils/MSV.py:84) turns the zero "
        "off-diagonal Jacobian blocks into -zhat/dx, JtJ becomes numerically singular and np.linalg.inv raises "
        "LinAlgError for every input (tests/golden/make_golden.py documents the check)")
