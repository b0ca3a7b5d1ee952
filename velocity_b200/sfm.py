"""The connected SFM speed-estimation sequence on the GPU (BASELINE.json configs[2], "C3").

The reference's frame loop (vidExample.py:75-166) with its state kept in HBM:

    frame 0   p = corners, p3 = planar 3-D points from the plate pose          vidExample.py:107-130
    frame i   p, v = KLT(im, im0, p); vg[vg] = v                               :134-135   K1 + K2 (vel_klt_sequence)
              t = estimateWorldCameraPose(K, p[vp], p3[vp], findR=False)       :139       fcnNLS_t  (vel_seq_pose_t)
              dr = |t_i - t_(i-1)|, speed = dr / dt * 3.6                       :142-146   (vel_seq_stats)
              P[0:2, vg, i] = p, P[2:4, vp, i] = p_proj, B[i], S[i]             :151-164   device arrays (f4)
    end       pw = fcnNvintercept(origins, rays over all frames)               utils/MSV.py:146-175   (K6)
              cw, pw = fcnNLS_batch(K, P, pw, B[:, 3:6])                        :157, utils/NLS.py:186-250   (K7 + K8)

Tracking is the reference's cv2calcOpticalFlowPyrLK wrapper with the forward-backward gate (utils/KLT.py:37-51), applied
frame to frame with the tracked points PROPAGATED (SURVEY.md 8(d) C3 allows "C2-style plain LK"; the three-stage KLTmain
is available as velocity_b200.KLT.KLTmain).  Nothing is computed on the host: the per-frame loop runs inside
libvelocity_b200.so, the only synchronisations are the track count before the bundle adjustment and its convergence test.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_ptr
from .lk import FrameBatch, lk_params
from .NLS import BundleAdjuster

LK_C3 = dict(winSize=(15, 15), maxLevel=2, criteria=(3, 10, 0.1))


class SfmSequence:
    """Buffers and launch plan for sequences of `nframes` frames x `npts` tracks.  `run()` may be called repeatedly."""

    def __init__(self, K, height, width, nframes, npts, fbt=1.0, ba_iters=10, chunk=25, device=None, **lk_param):
        require_cuda()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.h, self.w, self.n, self.npts = height, width, nframes, npts
        self.ba_iters, self.chunk = ba_iters, max(1, min(chunk, nframes))
        lk_param = lk_param or LK_C3
        self.params = lk_params(fbt=fbt, **lk_param)
        self.win = (self.params.win_w, self.params.win_h)
        dev = self.dev
        self.K = torch.from_numpy(np.ascontiguousarray(np.asarray(K, np.float64))).to(dev)
        self.frames = torch.empty((nframes, height, width), dtype=torch.uint8, device=dev)     # filled per run (e2e) or aliased
        self._bufs = [self.frames, None]     # host frames land in one of two device buffers (the second is allocated by the first prefetch)
        self._buf_free = [None, None]        # event: the tracking launches that last read the buffer have been issued and completed
        self._upload_buf = 0
        self._pending = []                   # FIFO of uploads in flight: (host data_ptr, buffer index, per-chunk copy events, staged small inputs)
        self._small = [{}, {}]               # per frame buffer: device copies of the small per-sequence inputs (p0, p3, frame_times)
        self._cur_small = {}                 # the staged small inputs of the sequence being run
        self._in = {}                        # their per-sequence device copies (taken out of the staging slot when tracking starts)
        self.batch = FrameBatch(self.frames, self.win, self.params.max_level)
        self.tracks = torch.empty((nframes, npts, 2), dtype=torch.float32, device=dev)
        self.alive = torch.empty((nframes, npts), dtype=torch.uint8, device=dev)
        self.err = torch.empty((max(nframes - 1, 1), npts), dtype=torch.float32, device=dev)
        self.status = torch.empty((npts,), dtype=torch.uint8, device=dev)
        self.proj = torch.empty((nframes, npts, 2), dtype=torch.float32, device=dev)
        self.B = torch.zeros((nframes, 14), dtype=torch.float32, device=dev)
        self.S = torch.zeros((nframes, 9), dtype=torch.float32, device=dev)
        self.B_ba = torch.zeros((nframes, 14), dtype=torch.float32, device=dev)
        self.S_ba = torch.zeros((nframes, 9), dtype=torch.float32, device=dev)
        self.iters = torch.zeros((nframes,), dtype=torch.int32, device=dev)
        self.p3 = torch.empty((npts, 3), dtype=torch.float64, device=dev)
        self.idx = torch.empty((npts,), dtype=torch.int32, device=dev)
        self.count = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.U = torch.empty((3, nframes, npts), dtype=torch.float64, device=dev)
        self.A = torch.empty((nframes, 3), dtype=torch.float64, device=dev)
        self.C0 = torch.empty((npts, 3), dtype=torch.float64, device=dev)
        self.z = torch.empty((2 * nframes * npts,), dtype=torch.float64, device=dev)
        self.x0 = torch.empty((3 * npts + 6 * (nframes - 1),), dtype=torch.float64, device=dev)
        self.P = None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._back_stream, self._back = None, None              # read-back of run(..., sync=False)
        self.x0_host = (C.c_double * 3)(0.0, 0.0, 1.0)       # estimateWorldCameraPose's default t (utils/NLS.py:9)
        self.ba = None
        self.launches = 0
        self.h2d_bytes = self.d2h_bytes = 0
        self.marks = None          # set to [] to collect (stage name, CUDA event) pairs during run()

    def _mark(self, name):
        if self.marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(torch.cuda.current_stream(self.dev))
            self.marks.append((name, ev))

    # ---- stage 1: pyramids of frames [pyr_lo, hi] + propagated tracking over frames lo..hi (tracks row lo given) ------
    def _track_range(self, lo, hi, pyr_lo):
        L = _lib.lib()
        fr, fb = self.frames, self.batch
        if fb.layout.max_level > 0 and hi >= pyr_lo:
            _lib.check(L.vel_pyramid_u8(C.c_void_p(fr.data_ptr() + pyr_lo * fr.stride(0)), fr.stride(0), fr.stride(1), hi - pyr_lo + 1,
                                        C.byref(fb.layout), C.c_void_p(fb.pyr.data_ptr() + pyr_lo * fb.pyr.stride(0)), fb.pyr.stride(0),
                                        stream_ptr()), "vel_pyramid_u8")
            self.launches += fb.pyramid_launches()
        nfr = hi - lo + 1
        if nfr >= 2:
            _lib.check(L.vel_klt_sequence(C.c_void_p(fr.data_ptr() + lo * fr.stride(0)), fr.stride(0), fr.stride(1),
                                          C.c_void_p(fb.pyr.data_ptr() + lo * fb.pyr.stride(0)), fb.pyr.stride(0), C.byref(fb.layout), nfr,
                                          self.npts, C.byref(self.params), ptr(self.tracks[lo]), ptr(self.alive[lo]), ptr(self.err[lo]),
                                          ptr(self.status), stream_ptr()), "vel_klt_sequence")
            self.launches += self._sequence_launches(nfr)

    def _sequence_launches(self, nfr):
        """Kernel launches of one vel_klt_sequence call over nfr frames: seed + ONE kernel that walks all pairs for the 15x15
        window on word-aligned pitches (csrc/lk_track.cu: vel_lk_sequence_w15h), else seed + (track, propagate) per pair."""
        import os

        fb = self.batch
        one = (self.win == (15, 15) and fb.pitch % 4 == 0 and (fb.frames.stride(0) % 4 == 0)
               and all(fb.layout.pitch[l] % 4 == 0 for l in range(1, fb.layout.max_level + 1))
               and os.environ.get("VEL_LK_W15") != "bytes" and os.environ.get("VEL_LK_SEQ") != "pairs")
        return 2 if one else 1 + 2 * (nfr - 1)

    def prefetch(self, frames, p0=None, p3=None, frame_times=None):
        """Start uploading the NEXT sequence's frames (pinned host tensor) into the idle one of two device buffers, chunk by
        chunk on the copy stream.  Call it before run() of the CURRENT sequence: the upload then overlaps the current
        sequence's tracking and bundle adjustment (neither reads the idle buffer), and the run() that later receives the same
        host tensor finds its frames already on their way.  Optional -- run() uploads by itself otherwise.
        The small per-sequence inputs (p0, p3, frame_times: host tensors) should be handed over as well: they are uploaded AHEAD
        of the frames.  Left to run(), their copies would queue in the host-to-device DMA engine behind whatever frame upload is
        in flight (the engine serves copies in issue order across streams) and the tracking would wait ~12 ms for 100 KB.
        run() uses the staged copies when it is given the same objects."""
        assert not frames.is_cuda and tuple(frames.shape) == (self.n, self.h, self.w) and frames.dtype == torch.uint8
        if len(self._pending) >= 2:
            raise RuntimeError("SfmSequence.prefetch: both device buffers already hold uploads that no run() has consumed")
        if self._bufs[1] is None:
            self._bufs[1] = torch.empty_like(self._bufs[0])
        b = self._upload_buf
        self._upload_buf = 1 - b
        dst = self._bufs[b]
        if self._buf_free[b] is not None:
            self.copy_stream.wait_event(self._buf_free[b])      # the tracking launches that last read this buffer
        else:
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
        events, small = [], {}
        with torch.cuda.stream(self.copy_stream):
            for name, src in (("p0", p0), ("p3", p3), ("frame_times", frame_times)):
                if isinstance(src, torch.Tensor) and not src.is_cuda:
                    if self._small[b].get(name) is None or self._small[b][name].shape != src.shape or self._small[b][name].dtype != src.dtype:
                        self._small[b][name] = torch.empty(src.shape, dtype=src.dtype, device=self.dev)
                    self._small[b][name].copy_(src, non_blocking=True)
                    small[name] = (src, self._small[b][name])
            for lo in range(0, self.n, self.chunk):
                hi = min(self.n, lo + self.chunk)
                dst[lo:hi].copy_(frames[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                events.append((lo, hi, ev))
        self._pending.append((frames.data_ptr(), b, events, small))

    def _staged(self, name, given):
        """The device copy prefetch() made of a small input, if `given` is the object it was made from."""
        ent = self._cur_small.get(name)
        return ent[1] if ent is not None and ent[0] is given else given

    def track(self, frames, p0, alive0=None):
        """Stage 1 alone.  frames: CUDA uint8 [n,H,W] (used in place) or a pinned host tensor (uploaded chunk by chunk on
        a copy stream while the previous chunk is being tracked; already in flight if prefetch() was given the same tensor)."""
        n = self.n
        assert tuple(frames.shape) == (n, self.h, self.w) and frames.dtype == torch.uint8
        compute = torch.cuda.current_stream(self.dev)
        self._cur_small = {}
        host_path = not frames.is_cuda
        p0_done = False
        if host_path:
            if not self._pending or self._pending[0][0] != frames.data_ptr():
                # not announced: the small input goes first (behind the frame chunks it would wait for the whole upload)
                self.tracks[0].copy_(p0 if isinstance(p0, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p0, np.float32)), non_blocking=True)
                p0_done = True
                self.prefetch(frames)
            _, b, events, self._cur_small = self._pending.pop(0) if self._pending[0][0] == frames.data_ptr() else self._pending.pop()
            compute.wait_event(events[0][2])                   # staged small inputs precede the first frame chunk on the copy stream
            # move them out of the per-buffer staging slots now (the slot is handed back together with the frame buffer below)
            for name, (src, dev_t) in list(self._cur_small.items()):
                if self._in.get(name) is None or self._in[name].shape != dev_t.shape or self._in[name].dtype != dev_t.dtype:
                    self._in[name] = torch.empty_like(dev_t)
                self._in[name].copy_(dev_t, non_blocking=True)
                self._cur_small[name] = (src, self._in[name])
            p0 = self._staged("p0", p0)
        if not p0_done:
            self.tracks[0].copy_(p0 if isinstance(p0, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p0, np.float32)), non_blocking=True)
        if alive0 is None:
            self.alive[0].fill_(1)
        else:
            self.alive[0].copy_(alive0)
        if frames.is_cuda:
            if frames.data_ptr() != self.frames.data_ptr():
                self.frames = frames if frames.is_contiguous() else frames.contiguous()
                self.batch.frames = self.frames
            self._track_range(0, n - 1, 0)
            return
        self.frames = self._bufs[b]
        self.batch.frames = self.frames
        self.h2d_bytes += n * self.h * self.w
        # chunk by chunk behind the upload -- but chunks that have already landed (a prefetched sequence usually has, all of it) are
        # taken together: one pyramid launch and one sequence kernel over their whole frame range instead of one pair per chunk
        i = 0
        while i < len(events):
            j = i
            while j + 1 < len(events) and events[j + 1][2].query():
                j += 1
            lo, hi = events[i][0], events[j][1]
            compute.wait_event(events[j][2])
            self._track_range(max(lo - 1, 0), hi - 1, lo)
            i = j + 1
        done = torch.cuda.Event()
        done.record(compute)
        self._buf_free[b] = done

    # ---- stages 2-4 -----------------------------------------------------------------------------------------------------
    def solve(self, p3, frame_times, t0=(0.0, 0.0, 0.0), subset=None, bundle=True, verbose=False, before_sync=None):
        """Per-frame translation + speed table, then triangulation over all frames and the bundle adjustment.
        before_sync: called once everything is enqueued and before the history is read back (the one synchronisation at the end):
        work enqueued there -- run()'s read-backs -- needs no synchronisation of its own."""
        L = _lib.lib()
        n, npts = self.n, self.npts
        p3, frame_times = self._staged("p3", p3), self._staged("frame_times", frame_times)
        self.p3.copy_(p3 if isinstance(p3, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p3, np.float64)), non_blocking=True)
        ft = frame_times if isinstance(frame_times, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(frame_times, np.float32))
        self.B[:, 12].copy_(ft, non_blocking=True)
        if getattr(self, "_t0_key", None) != tuple(t0):
            self._t0_key, self._t0_dev = tuple(t0), torch.tensor(t0, dtype=torch.float32).to(self.dev)
        self.B[0, 0:3].copy_(self._t0_dev, non_blocking=True)
        sub = ptr(subset) if subset is not None else C.c_void_p(0)
        _lib.check(L.vel_seq_pose_t(ptr(self.K), ptr(self.tracks), ptr(self.alive), sub, ptr(self.p3), n, npts, self.x0_host, ptr(self.B),
                                    ptr(self.S), ptr(self.proj), ptr(self.iters), stream_ptr()), "vel_seq_pose_t")
        _lib.check(L.vel_seq_stats(ptr(self.B), ptr(self.alive), n, npts, ptr(self.S), stream_ptr()), "vel_seq_stats")
        self.launches += 3
        self._mark("pose")
        if not bundle or n < 2:
            if before_sync is not None:
                before_sync()
            return None
        _lib.check(L.vel_seq_select(ptr(self.alive[n - 1]), sub, npts, ptr(self.idx), ptr(self.count), stream_ptr()), "vel_seq_select")
        nsel = int(self.count.item())                      # the one data-dependent size: full-length tracks (utils/NLS.py:190)
        self.nsel = nsel
        if nsel < 1:
            raise RuntimeError("SfmSequence: no track survived the whole sequence")
        U = self.U.view(-1)[:3 * n * nsel].view(3, n, nsel)
        _lib.check(L.vel_seq_rays(ptr(self.K), ptr(self.tracks), ptr(self.idx), n, npts, nsel, ptr(self.B), ptr(U), ptr(self.A), stream_ptr()),
                   "vel_seq_rays")
        C0 = self.C0[:nsel]
        _lib.check(L.vel_triangulate_nv(ptr(self.A), ptr(U), n, nsel, ptr(C0), stream_ptr()), "vel_triangulate_nv")
        z = self.z[:2 * n * nsel]
        x0 = self.x0[:3 * nsel + 6 * (n - 1)]
        _lib.check(L.vel_seq_pack_ba(ptr(self.tracks), ptr(self.idx), n, npts, nsel, ptr(C0), ptr(self.B), ptr(z), ptr(x0), stream_ptr()),
                   "vel_seq_pack_ba")
        self.launches += 7                                 # select, rays, triangulation (partial + solve), pack ...
        self._mark("triangulate")
        if self.ba is None or (self.ba.nt, self.ba.nc) != (nsel, n - 1):
            self.ba = BundleAdjuster(self.K, z, x0, nsel, n - 1)
        else:
            self.ba.reset(z, x0)
        # the whole iteration loop is enqueued at once (convergence test on the device); what follows does not depend on how many
        # iterations ran, so it is enqueued behind it and the history is read back last
        self.ba.timing = None
        _lib.check(L.vel_ba_iterate(ptr(self.ba.K), ptr(self.ba.z), nsel, n - 1, ptr(self.ba.x), self.ba_iters, 1e-7, *self.ba.loop_buffers(),
                                    stream_ptr()), "vel_ba_iterate")
        # B[:, 3:6] = cw (the commented call site, vidExample.py:157) on a copy, and the speed table that follows from it
        _lib.check(L.vel_seq_ba_cameras(ptr(self.ba.x), nsel, n, ptr(self.B), ptr(self.S), ptr(self.B_ba), ptr(self.S_ba), stream_ptr()),
                   "vel_seq_ba_cameras")
        _lib.check(L.vel_seq_stats(ptr(self.B_ba), ptr(self.alive), n, npts, ptr(self.S_ba), stream_ptr()), "vel_seq_stats")
        self._mark("bundle")
        if before_sync is not None:
            before_sync()
        hist = self.ba.history(self.ba_iters)
        if verbose:
            for it, (f, xr) in enumerate(hist):
                print(f"{it:g}: f={f:g}, x={xr}")
            if not hist or not hist[-1][1] < 1e-7:
                print("WARNING: fcnNLS_batch() reaching max iterations!")
        self.launches += 1 + self.ba.launches_per_loop_iteration * len(hist) + 2
        return hist

    def export_P(self):
        """The reference's P [5, npts, n] float32 array (device tensor)."""
        if self.P is None:
            self.P = torch.empty((5, self.npts, self.n), dtype=torch.float32, device=self.dev)
        _lib.check(_lib.lib().vel_seq_export_P(ptr(self.tracks), ptr(self.proj), ptr(self.alive), self.n, self.npts, ptr(self.P), stream_ptr()),
                   "vel_seq_export_P")
        self.launches += 1
        return self.P

    def points_ba(self):
        """(indices of the full-length tracks, their bundle-adjusted 3-D positions, bundle-adjusted camera positions)."""
        nsel, n = self.nsel, self.n
        x = self.ba.x
        return self.idx[:nsel], x[:3 * nsel].view(nsel, 3), torch.cat([torch.zeros((1, 3), dtype=torch.float64, device=self.dev),
                                                                       x[3 * nsel:3 * nsel + 3 * (n - 1)].view(n - 1, 3)])

    def run(self, frames, p0, p3, frame_times, t0=(0.0, 0.0, 0.0), out=None, bundle=True, sync=True):
        """One whole sequence.  With `out` (dict of pinned host tensors 'S', 'S_ba', 'B', and optionally 'P') the results
        are copied back and the call synchronises; otherwise everything stays on the device.
        sync=False (a caller that runs sequence after sequence): the small tables are on the host when the call returns, the
        large 'P' export (24.6 MB at C3) travels on a side stream while the NEXT sequence is already being tracked;
        wait_results() waits for it (copies into the same host buffer stay in order)."""
        self.h2d_bytes = self.d2h_bytes = 0
        self._mark("start")
        self.track(frames, p0)
        self._mark("track")
        compute = torch.cuda.current_stream(self.dev)
        if out is not None and not sync:
            def read_back():
                # behind the bundle adjustment on the compute stream: complete when solve() has read the history back
                for name, src in (("S", self.S), ("S_ba", self.S_ba), ("B", self.B)):
                    if name in out:
                        out[name].copy_(src, non_blocking=True)
                        self.d2h_bytes += src.numel() * 4
                if "P" in out:
                    if self._back is not None:
                        compute.wait_event(self._back[1])      # the export buffer is reused: the previous copy must have left it
                    exp = self.export_P()                      # nothing of the next sequence writes into it before its own export
                    ready = torch.cuda.Event()
                    ready.record(compute)
                    if self._back_stream is None:
                        self._back_stream = torch.cuda.Stream(device=self.dev)
                    self._back_stream.wait_event(ready)
                    with torch.cuda.stream(self._back_stream):
                        out["P"].copy_(exp, non_blocking=True)
                    done = torch.cuda.Event()
                    done.record(self._back_stream)
                    self._back = (exp, done)
                    self.d2h_bytes += self.P.numel() * 4

            hist = self.solve(p3, frame_times, t0=t0, bundle=bundle, before_sync=read_back)
            if hist is None:                                   # no bundle adjustment, hence no history read-back: synchronise here
                compute.synchronize()
            return hist
        hist = self.solve(p3, frame_times, t0=t0, bundle=bundle)
        if out is not None:
            self.wait_results()
            for name, src in (("S", self.S), ("S_ba", self.S_ba), ("B", self.B)):
                if name in out:
                    out[name].copy_(src, non_blocking=True)
                    self.d2h_bytes += src.numel() * 4
            if "P" in out:
                out["P"].copy_(self.export_P(), non_blocking=True)
                self.d2h_bytes += self.P.numel() * 4
            compute.synchronize()
        return hist

    def wait_results(self):
        """Block until the read-back a run(..., sync=False) left in flight has landed in its host buffers."""
        if self._back is not None:
            self._back[1].synchronize()
            self._back = None


def plane_points_from_pose(K, R, t, p):
    """vidExample.py:119: p3 = addcol0(image2world(K, R, t, p)) @ R + t -- the planar 3-D points in the frame-0 camera
    frame (one-shot initialisation on the host, like the reference)."""
    from .common import addcol0, image2world

    return addcol0(image2world(np.asarray(K, float), np.asarray(R, float), np.asarray(t, float), np.asarray(p)).astype(float)) @ np.asarray(R, float) + np.asarray(t, float)
