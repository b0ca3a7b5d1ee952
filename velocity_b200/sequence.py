"""Sequence-mode tracking through HOST buffers: the public end-to-end entry point.

`track_sequence(frames_host, pts, ...)` is what a caller holding decoded video frames in host
memory uses: frames are staged chunk by chunk from pinned memory on a copy stream while the
previous chunk's pyramids (K1) and tracks (K2) run on the compute stream; per-pair results come
back to the host.  Each frame is uploaded once and its pyramid is built once (it serves as `next`
of pair k-1 and `prev` of pair k); consecutive chunks overlap by one frame, which is copied
device-to-device rather than re-uploaded.

Semantics per pair are exactly utils/KLT.py:37-51 (cv2calcOpticalFlowPyrLK with fbt): every pair
tracks the SAME seed points `pts` from frame k to frame k+1 (independent pairs -- the C2 workload
of BASELINE.json).
"""
import numpy as np
import torch

from .device import require_cuda
from .lk import FrameBatch, lk_params, track_pairs


class SequenceTracker:
    def __init__(self, height, width, npts, chunk=32, device=None, fbt=1.0, **lk_param):
        require_cuda()
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.h, self.w, self.npts, self.chunk = height, width, npts, chunk
        self.params = lk_params(fbt=fbt, **lk_param)
        self.win = (self.params.win_w, self.params.win_h)
        # two device slots of chunk+1 frames: [0] is the carried-over last frame of the previous chunk
        self.slots = [torch.empty((chunk + 1, height, width), dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.batches = [FrameBatch(s, self.win, self.params.max_level) for s in self.slots]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.launches = 0
        self._graph = self._graph_key = self._graph_bytes = self._graph_keep = None

    def run(self, frames_host, pts_host, out_pts, out_status, out_err, graph=False):
        """frames_host: pinned uint8 [F,H,W]; pts_host: pinned float32 [npts,2];
        out_*: pinned host tensors [F-1,npts,2] f32 / [F-1,npts] u8 / [F-1,npts] f32.
        Returns (h2d_bytes, d2h_bytes).  Synchronises once at the end.

        graph=True captures the whole chunk pipeline (every H2D copy, K1/K2 launch and D2H copy on both streams) into
        one CUDA graph the first time it sees a set of buffers and replays it afterwards: a step then costs one launch
        from the host, so the copy engine is never left waiting for Python to issue the next chunk."""
        if graph:
            key = (frames_host.data_ptr(), tuple(frames_host.shape), pts_host.data_ptr(), out_pts.data_ptr(), out_status.data_ptr(),
                   out_err.data_ptr())
            if self._graph_key != key:
                g = torch.cuda.CUDAGraph()
                side = torch.cuda.Stream(device=self.dev)
                side.wait_stream(torch.cuda.current_stream(self.dev))
                with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):   # other threads (NCCL watchdog) may touch CUDA
                    self._graph_bytes = self._run_chunks(frames_host, pts_host, out_pts, out_status, out_err, capturing=True)
                self._graph, self._graph_key = g, key
            self._graph.replay()
            torch.cuda.current_stream(self.dev).synchronize()
            return self._graph_bytes
        return self._run_chunks(frames_host, pts_host, out_pts, out_status, out_err, capturing=False)

    def _run_chunks(self, frames_host, pts_host, out_pts, out_status, out_err, capturing):
        F = frames_host.shape[0]
        compute = torch.cuda.current_stream(self.dev)
        self.copy_stream.wait_stream(compute)          # the copy stream forks from the compute stream (required under capture)
        pts = pts_host.to(self.dev, non_blocking=True)
        h2d, d2h = pts_host.numel() * 4, 0
        nchunks = (F - 1 + self.chunk - 1) // self.chunk
        upload_done = [None, None]
        slot_free = [None, None]
        results = []

        def upload(c):
            s = c & 1
            first = 1 + c * self.chunk  # frames first..last go to slot rows 1..n
            last = min(F - 1, first + self.chunk - 1)
            n = last - first + 1
            with torch.cuda.stream(self.copy_stream):
                if slot_free[s] is not None:
                    self.copy_stream.wait_event(slot_free[s])
                self.slots[s][1:1 + n].copy_(frames_host[first:last + 1], non_blocking=True)
                if c == 0:
                    self.slots[s][0].copy_(frames_host[0], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            upload_done[s] = ev
            return n

        counts = {}
        if nchunks > 0:
            counts[0] = upload(0)
            h2d += (counts[0] + 1) * self.h * self.w
        for c in range(nchunks):
            s = c & 1
            if c + 1 < nchunks:
                counts[c + 1] = upload(c + 1)
                h2d += counts[c + 1] * self.h * self.w
            n = counts[c]
            compute.wait_event(upload_done[s])
            fb = self.batches[s]
            sub = fb if n == self.chunk else self._sub_batch(fb, n + 1)
            sub.build()
            self.launches += sub.pyramid_launches()
            out, st, err, _ = track_pairs(sub, sub, pts, self.params, 0, 1, n)
            self.launches += 1
            lo = c * self.chunk
            out_pts[lo:lo + n].copy_(out, non_blocking=True)
            out_status[lo:lo + n].copy_(st, non_blocking=True)
            out_err[lo:lo + n].copy_(err, non_blocking=True)
            d2h += out.numel() * 4 + st.numel() + err.numel() * 4
            results.append((out, st, err))  # keep alive until the copies are queued behind them
            if c + 1 < nchunks:  # carry this chunk's last frame into row 0 of the other slot (device to device)
                self.slots[s ^ 1][0].copy_(self.slots[s][n], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(compute)
            slot_free[s] = ev
        compute.wait_stream(self.copy_stream)           # join
        if capturing:
            self._graph_keep = (results, pts)            # tensors of the graph's private pool stay referenced
        else:
            compute.synchronize()
        return h2d, d2h

    def _sub_batch(self, fb, nframes):
        sub = FrameBatch.__new__(FrameBatch)
        sub.frames = fb.frames[:nframes]
        sub.n, sub.h, sub.w = nframes, fb.h, fb.w
        sub.pitch, sub.frame_stride = fb.pitch, fb.frames.stride(0)
        sub.layout, sub.pyr, sub.built = fb.layout, fb.pyr[:nframes], False
        return sub


def track_sequence(frames, pts, fbt=1.0, chunk=32, **lk_param):
    """Convenience wrapper: numpy in, numpy out.  frames uint8 [F,H,W], pts float32 [N,2].
    Returns (p2 [F-1,N,2] f32, v [F-1,N] bool, err [F-1,N] f32)."""
    frames = np.ascontiguousarray(frames)
    F, H, W = frames.shape
    pts = np.ascontiguousarray(np.asarray(pts, np.float32).reshape(-1, 2))
    tr = SequenceTracker(H, W, pts.shape[0], chunk=min(chunk, max(F - 1, 1)), fbt=fbt, **lk_param)
    fh = torch.from_numpy(frames).pin_memory()
    ph = torch.from_numpy(pts).pin_memory()
    op = torch.empty((F - 1, pts.shape[0], 2), dtype=torch.float32).pin_memory()
    os_ = torch.empty((F - 1, pts.shape[0]), dtype=torch.uint8).pin_memory()
    oe = torch.empty((F - 1, pts.shape[0]), dtype=torch.float32).pin_memory()
    tr.run(fh, ph, op, os_, oe)
    return op.numpy(), os_.numpy() != 0, oe.numpy()
