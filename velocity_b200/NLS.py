"""Drop-in for the reference's utils/NLS.py: same names, arguments, dtypes and printed messages.

    estimateWorldCameraPose  utils/NLS.py:9-33
    fzK / fzC                utils/NLS.py:71-86
    fcnNLS_t                 utils/NLS.py:102-129   -> K5 vel_nls_t
    fcnNLS_Rt                utils/NLS.py:133-183   -> K5 vel_nls_rt
    fcnNLS_batch             utils/NLS.py:186-250   -> K7 vel_ba_accumulate + K8 vel_ba_solve

The solvers run on the GPU in float64 with the reference's forward-difference Jacobians; only
argument packing, the iteration printouts and the final float32 casts happen on the host.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from .ba_exchange import camera_slices, exchange_blocks, gather_rows_to_owner, small_buffer, tile_row_ranges
from .common import addcol1, cam2ned, cc2sc, pscale, rms, sc2cc, world2image
from .device import ptr, require_cuda, stream_ptr
from .transforms import dcm2rpy, rpy2dcm


def fzK(a, K):
    return pscale(a @ K)


def fzC(a, K, R, t=np.zeros((1, 3))):
    return pscale(addcol1(a) @ (np.concatenate([R, t]) @ K))


def _dev64(a):
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64))).cuda()


def nls_batch_device(K, p, pw, first, count, x0, dof):
    """Batched K5 call on CUDA float64 tensors.  Returns (x [nprob,dof], iters [nprob] int32)."""
    require_cuda()
    nprob = first.shape[0]
    x = torch.empty((nprob, dof), dtype=torch.float64, device=p.device)
    iters = torch.empty((nprob,), dtype=torch.int32, device=p.device)
    fn = _lib.lib().vel_nls_t if dof == 3 else _lib.lib().vel_nls_rt
    _lib.check(fn(ptr(K), ptr(p), ptr(pw), ptr(first), ptr(count), nprob, ptr(x0), ptr(x), ptr(iters), stream_ptr()),
               "vel_nls_t" if dof == 3 else "vel_nls_rt")
    return x, iters


def _single(K, p, pw, x, dof):
    """One pose problem from host arrays: ONE upload (K | x0 | p | pw | first,count packed in a float64 buffer) and ONE read-back
    (x | iters) -- the per-frame call of the drop-in loop (vidExample.py:139)."""
    require_cuda()
    pw = np.asarray(pw, np.float64).reshape(-1, 3)
    p = np.asarray(p, np.float64).reshape(-1, 2)
    n = pw.shape[0]
    host = np.empty(9 + dof + 5 * n + 1, np.float64)
    host[0:9] = np.asarray(K, np.float64).ravel()
    host[9:9 + dof] = np.asarray(x, np.float64).ravel()[:dof]
    o_p, o_pw, o_int = 9 + dof, 9 + dof + 2 * n, 9 + dof + 5 * n
    host[o_p:o_pw] = p.ravel()
    host[o_pw:o_int] = pw.ravel()
    host[o_int:o_int + 1].view(np.int32)[:] = (0, n)                      # first, count
    dev = torch.from_numpy(host).cuda()
    out = torch.empty((dof + 1,), dtype=torch.float64, device=dev.device)   # x | iters (int32 in the last 8 bytes)
    base, obase = dev.data_ptr(), out.data_ptr()
    fn = _lib.lib().vel_nls_t if dof == 3 else _lib.lib().vel_nls_rt
    _lib.check(fn(C.c_void_p(base), C.c_void_p(base + 8 * o_p), C.c_void_p(base + 8 * o_pw), C.c_void_p(base + 8 * o_int), C.c_void_p(base + 8 * o_int + 4),
                  1, C.c_void_p(base + 72), C.c_void_p(obase), C.c_void_p(obase + 8 * dof), stream_ptr()), "vel_nls_t" if dof == 3 else "vel_nls_rt")
    res = out.cpu().numpy()
    return res[:dof].copy(), int(res[dof:dof + 1].view(np.int32)[0])


def fcnNLS_t(K, p, pw, x):
    """3-dof translation fit; returns float32 [3]."""
    xs, it = _single(K, p, pw, x, 3)
    if it < 0:
        print("WARNING: fcnNLS_t() reaching max iterations!")
    return xs.astype(np.float32)


def fcnNLS_Rt(K, p, pw, x):
    """6-dof fit; returns (R float32 3x3, t float32 [3])."""
    xs, it = _single(K, p, pw, x, 6)
    if it < 0:
        print("WARNING: fcnNLS_Rt() reaching max iterations!")
    return rpy2dcm(xs[:3]).astype(np.float32), xs[3:6].astype(np.float32)


def estimateWorldCameraPose(K, p, p3, t=np.array([0, 0, 1]), R=np.eye(3), findR=False):
    x0 = np.concatenate((dcm2rpy(R), t))
    if findR is True:
        R, t = fcnNLS_Rt(K.astype(float), p.astype(float), p3, x0)
    else:
        t = fcnNLS_t(K.astype(float), p.astype(float), p3, t)
    p_proj = world2image(K, R, t, p3)
    residuals = rms(p - p_proj)
    return t, R, residuals, p_proj


class BundleAdjuster:
    """Device-resident state of one fcnNLS_batch problem (K7 + K8).  With torch.distributed initialised and `shard=True`,
    cameras are split across ranks (ba_exchange.py): each rank accumulates the blocks of its own cameras, ONE all-reduce
    carries V / g_p / cost (partial sums) together with U / g_c (owner entries), ONE all-gather carries the cross blocks W;
    each rank then forms its tile rows of the reduced camera system, the owner (rank 0) factors it once and broadcasts
    delta_c, and every rank applies the same update."""

    launches_per_step = 15   # this library's kernels per LM iteration (K7: 5; K8: prep, scale, init, SYRK, 2+1 GEMV, Cholesky, update, rms)

    def __init__(self, K, z, x0, nt, nc, shard=False, group=None):
        require_cuda()
        self.nt, self.nc = nt, nc
        dev = torch.device("cuda", torch.cuda.current_device())
        self.K = _dev64(K)
        self.z = _dev64(z)
        self.x = _dev64(x0).clone()
        self.group = group
        self.rank, self.world = 0, 1
        if shard and torch.distributed.is_available() and torch.distributed.is_initialized():
            self.rank, self.world = torch.distributed.get_rank(group), torch.distributed.get_world_size(group)
        self.per, self.slices = camera_slices(nc, self.world)
        self.rms_delta = torch.zeros((1,), dtype=torch.float64, device=dev)
        self.timing = None
        self._staged = False
        if self.world > 1:
            self._stage_buffers()

    _STAGED = ("small", "cost", "V", "g", "U", "W_ext", "W", "work", "S", "rhs", "blocks", "row_ranges")

    def __getattr__(self, name):            # only reached when normal lookup fails: the lazily allocated stage buffers
        if name in BundleAdjuster._STAGED and not self.__dict__.get("_staged", True):
            self._stage_buffers()
            if name in self.__dict__:
                return self.__dict__[name]
        raise AttributeError(name)

    def _stage_buffers(self):
        """Buffers of the stage-by-stage form (accumulate / solve / step): the K7 blocks and the K8 workspace.  iterate() keeps
        its own single workspace (the cross blocks exist only in scaled form there), so these are allocated on first use."""
        if self._staged:
            return
        self._staged = True
        nt, nc, dev, L = self.nt, self.nc, self.x.device, _lib.lib()
        self.small, self.cost, self.V, self.g, self.U = small_buffer(nt, nc, dev)
        # one 6-row block per camera 0..per*world-1 (camera 0's block is padding); the solver's W starts at camera 1
        self.W_ext = torch.zeros((6 * self.per * self.world, 3 * nt), dtype=torch.float64, device=dev)
        self.W = self.W_ext[6:]
        nbytes = int(L.vel_ba_solve_workspace(nt, nc))
        if nbytes == 0:
            raise RuntimeError("vel_ba_solve_workspace failed: %s" % L.vel_last_error().decode())
        self.work = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        if self.world > 1:
            off_S, off_rhs = C.c_int64(0), C.c_int64(0)
            _lib.check(L.vel_ba_solve_layout(nt, nc, C.byref(off_S), C.byref(off_rhs)), "vel_ba_solve_layout")
            n6 = 6 * nc
            self.S = self.work[off_S.value:off_S.value + 8 * n6 * n6].view(torch.float64).view(n6, n6)
            self.rhs = self.work[off_rhs.value:off_rhs.value + 8 * n6].view(torch.float64)
            nb, bme = C.c_int32(0), C.c_int32(0)
            _lib.check(L.vel_syrk_tile_rows(n6, C.byref(nb), C.byref(bme)), "vel_syrk_tile_rows")
            self.blocks = tile_row_ranges(nb.value, self.world)
            self.row_ranges = [(min(n6, lo * bme.value), min(n6, hi * bme.value)) for lo, hi in self.blocks]

    def reset(self, z, x0):
        """New measurements / start values for a problem of the same size (buffers are reused)."""
        self.z = _dev64(z)
        self.x.copy_(_dev64(x0))

    def _t(self, name):
        """stage marks for bench.py: self.timing = {} switches them on (CUDA events on the launching stream)"""
        if self.timing is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.timing.setdefault("_marks", []).append((name, ev))

    def stage_ms(self):
        """Sum per stage name over the recorded marks (call after a synchronise)."""
        out = {}
        marks = self.timing.get("_marks", []) if self.timing else []
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            if name != "begin":
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
        return out

    def accumulate(self):
        L = _lib.lib()
        self._stage_buffers()
        first, count = self.slices[self.rank]
        self._t("begin")
        if self.world > 1:
            self.small.zero_()               # foreign cameras' U / g_c entries must be exact zeros for the gathering all-reduce
        _lib.check(L.vel_ba_accumulate(ptr(self.K), ptr(self.x), ptr(self.z), self.nt, self.nc, first, count, ptr(self.V),
                                       ptr(self.U), ptr(self.W), ptr(self.g), ptr(self.cost), stream_ptr()), "vel_ba_accumulate")
        self._t("accumulate")
        if self.world > 1:
            exchange_blocks(self.small, self.W_ext, self.per, self.rank, self.world, self.group)
            self._t("exchange_allreduce_allgather")

    def solve(self):
        L = _lib.lib()
        self._stage_buffers()
        if self.world == 1:
            _lib.check(L.vel_ba_solve(ptr(self.V), ptr(self.U), ptr(self.W), ptr(self.g), self.nt, self.nc, ptr(self.x),
                                      ptr(self.rms_delta), ptr(self.work), self.work.numel(), stream_ptr()), "vel_ba_solve")
            self._t("solve")
            return
        lo, hi = self.blocks[self.rank]
        _lib.check(L.vel_ba_reduce(ptr(self.V), ptr(self.U), ptr(self.W), ptr(self.g), self.nt, self.nc, lo, hi, ptr(self.work),
                                   self.work.numel(), stream_ptr()), "vel_ba_reduce")
        self._t("reduce_tile_rows")
        gather_rows_to_owner(self.S, self.row_ranges, self.rank, 0, self.group)
        self._t("rows_to_owner")
        if self.rank == 0:
            _lib.check(L.vel_ba_factor(self.nt, self.nc, ptr(self.work), self.work.numel(), stream_ptr()), "vel_ba_factor")
        self._t("factor_on_owner")
        torch.distributed.broadcast(self.rhs, src=0 if self.group is None else torch.distributed.get_global_rank(self.group, 0),
                                    group=self.group)
        self._t("broadcast_delta_c")
        _lib.check(L.vel_ba_update(ptr(self.W), self.nt, self.nc, ptr(self.x), ptr(self.rms_delta), ptr(self.work), self.work.numel(),
                                   stream_ptr()), "vel_ba_update")
        self._t("update")

    def iterate(self, max_iter=10, tol=1e-7):
        """The reference's whole iteration loop (utils/NLS.py:222-242) enqueued at once: the `rms(delta) < tol: break` test runs
        on the device (vel_ba_iterate), the host reads the history back afterwards.  Returns [(f, rms(delta)), ...] for the
        iterations that ran -- the same ones a step-by-step loop with a host-side test would run.  Single rank only."""
        if self.world != 1:
            raise RuntimeError("BundleAdjuster.iterate: the sharded form exchanges blocks between stages -- use step()")
        L = _lib.lib()
        assert max_iter <= 64
        self._t("begin")
        _lib.check(L.vel_ba_iterate(ptr(self.K), ptr(self.z), self.nt, self.nc, ptr(self.x), max_iter, float(tol), *self.loop_buffers(),
                                    stream_ptr()), "vel_ba_iterate")
        self._t("iterate")
        return self.history(max_iter)

    def loop_buffers(self):
        """(hist, iters_run, work, work_bytes) arguments of vel_ba_iterate, allocated on first use"""
        if getattr(self, "_loop_work", None) is None:
            nbytes = int(_lib.lib().vel_ba_iterate_workspace(self.nt, self.nc))
            if nbytes == 0:
                raise RuntimeError("vel_ba_iterate_workspace failed: %s" % _lib.lib().vel_last_error().decode())
            self._loop_work = torch.empty((nbytes,), dtype=torch.uint8, device=self.x.device)
            self._hist = torch.empty((64, 2), dtype=torch.float64, device=self.x.device)
            self._iters = torch.zeros((1,), dtype=torch.int32, device=self.x.device)
        return ptr(self._hist), ptr(self._iters), ptr(self._loop_work), self._loop_work.numel()

    launches_per_loop_iteration = 10   # cam setup, point, reduce+prep, camera, camera reduce (+ clears S), SYRK, Cholesky, GEMV, update, finalize

    def history(self, max_iter):
        """[(f, rms(delta))] of the last iterate() call (synchronises on the readback)."""
        h = self._hist[:max_iter].cpu().numpy()
        nz = 2 * self.nt * (self.nc + 1)
        ran = int(np.isfinite(h[:, 0]).sum())
        return [(float(np.sqrt(h[i, 0] / nz)), float(h[i, 1])) for i in range(ran)]

    def step(self):
        """One LM iteration.  Returns (f = rms(z - zhat) before the update, rms(delta))."""
        self.accumulate()
        self.solve()
        out = torch.stack([self.cost[0], self.rms_delta[0]]).cpu().numpy()
        nz = 2 * self.nt * (self.nc + 1)
        return float(np.sqrt(out[0] / nz)), float(out[1])


def fcnNLS_batch(K, P, pw, cw):
    """Bundle adjustment over tie points, camera positions and camera roll/pitch/yaw (camera 0
    fixed).  Same update rule, iteration cap (10), stopping test and printouts as the reference;
    returns (cw [nc+1,3], pw [nt,3]) float64."""
    v = np.isfinite(P[4]).sum(1) == P.shape[2]
    P, pw = P[:, v], pw[v]
    _, nt, ncam = P.shape
    nc = ncam - 1
    if np.isnan(P[:2]).any():
        raise ValueError("fcnNLS_batch: NaN pixel in a full-length track (the reference zeroes the residual but not the "
                         "Jacobian row, utils/NLS.py:200-201,233 -- undefined behaviour, refused here)")
    z = np.concatenate((P[0].T.ravel(), P[1].T.ravel())).astype(np.float64)  # [2][nc+1][nt]
    x0 = np.concatenate((np.asarray(pw, np.float64), np.asarray(cw, np.float64)[1:], np.zeros((nc, 3)))).ravel()
    ba = BundleAdjuster(np.asarray(K, float), z, x0, nt, nc)
    max_iter = 10
    tic = time.time()
    hist = ba.iterate(max_iter, 1e-7)              # the loop of utils/NLS.py:222-242, convergence test on the device
    dt = (time.time() - tic) / max(len(hist), 1)
    i, f = len(hist) - 1, float("nan")
    for i, (f, xr) in enumerate(hist):
        print(f"{i:g}: {dt:.3f}s, f={f:g}, x={xr}")
    if not hist or not hist[-1][1] < 1e-7:
        print("WARNING: fcnNLS_batch() reaching max iterations!")
    print(f"fcnNLS_batch done in {i:g} steps, {dt:.3f}s, f={f:g}")
    x = ba.x.cpu().numpy()
    j = nt * 3
    pw_out = x[:j].reshape(nt, 3)
    cw_out = np.concatenate((np.zeros((1, 3)), x[j:j + nc * 3].reshape(nc, 3)), 0)
    return cw_out, pw_out


def fcnNLS_batch2(K, P, pw, cw):
    """Bundle adjustment with the camera track parametrised as one joint rotation, one direction
    (elevation, azimuth) and a range per camera (utils/NLS.py:253-328).  Same update rule as
    fcnNLS_batch, <= 20 iterations; returns (cw [nc+1,3], pw [nt,3]) float64."""
    require_cuda()
    v = np.isfinite(P[4]).sum(1) == P.shape[2]
    P, pw = P[:, v], pw[v]
    _, nt, ncam = P.shape
    nc = ncam - 1
    if np.isnan(P[:2]).any():
        raise ValueError("fcnNLS_batch2: NaN pixel in a full-length track (undefined in the reference, refused here)")
    Cn = cam2ned()
    z = np.concatenate((P[0].T.ravel(), P[1].T.ravel())).astype(np.float64)
    sc = cc2sc(Cn @ (np.asarray(cw, float)[1] - np.asarray(cw, float)[0]))
    ranges = np.arange(1, nc + 1) * sc[0]
    x0 = np.concatenate((np.asarray(pw, np.float64).ravel(), np.zeros(3), sc[1:3], ranges))
    nq = 5 + nc
    dev = torch.device("cuda", torch.cuda.current_device())
    Kd, zd, xd = _dev64(np.asarray(K, float)), _dev64(z), _dev64(x0).clone()
    V = torch.zeros((nt, 6), dtype=torch.float64, device=dev)
    G = torch.zeros((nq, nq), dtype=torch.float64, device=dev)
    W = torch.zeros((nq, 3 * nt), dtype=torch.float64, device=dev)
    g = torch.zeros((3 * nt + nq,), dtype=torch.float64, device=dev)
    stats = torch.zeros((2,), dtype=torch.float64, device=dev)
    L = _lib.lib()
    nbytes = int(L.vel_ba_solve_workspace(nt, (nq + 5) // 6))
    if nbytes == 0:
        raise RuntimeError("vel_ba_solve_workspace failed: %s" % L.vel_last_error().decode())
    work = torch.empty((nbytes + (1 << 20),), dtype=torch.uint8, device=dev)
    max_iter = 20
    i, f, tic = 0, float("nan"), time.time()
    for i in range(max_iter):
        tic = time.time()
        _lib.check(L.vel_ba2_accumulate(ptr(Kd), ptr(xd), ptr(zd), nt, nc, ptr(V), ptr(G), ptr(W), ptr(g), ptr(stats[0:1]),
                                        stream_ptr()), "vel_ba2_accumulate")
        _lib.check(L.vel_ba2_solve(ptr(V), ptr(G), ptr(W), ptr(g), nt, nq, ptr(xd), ptr(stats[1:2]), ptr(work), work.numel(),
                                   stream_ptr()), "vel_ba2_solve")
        cost, xr = stats.cpu().numpy()
        f = float(np.sqrt(cost / z.size))
        if xr < 1e-7:
            break
    else:
        print("WARNING: fcnNLS_batch() reaching max iterations!")
    print(f"fcnNLS_batch2 done in {i:g} steps, {time.time() - tic:.3f}s, f={f:g}")
    x = xd.cpu().numpy()
    j = nt * 3
    scm = np.zeros((nc, 3))
    scm[:, 0] = x[j + 5:j + 5 + nc]
    scm[:, 1] = x[j + 3]
    scm[:, 2] = x[j + 4]
    cw_out = np.concatenate((np.zeros((1, 3)), sc2cc(scm) @ Cn), 0)
    return cw_out, x[:j].reshape(nt, 3)
