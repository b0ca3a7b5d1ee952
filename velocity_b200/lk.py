"""Batched pyramid + Lucas-Kanade engine over the C ABI (K1 + K2).

This is the device-resident layer the reference-facing functions in KLT.py are built on and the
one bench.py drives: frames stay in HBM, each frame's pyramid is built once and used in both the
`prev` and the `next` role of consecutive pairs (SURVEY.md 8(d) "sequence steady state").
"""
import ctypes as C

import torch

from . import _lib
from .device import ptr, require_cuda, stream_ptr

TERM_COUNT, TERM_EPS = 1, 2


def lk_params(winSize=(21, 21), maxLevel=3, criteria=(TERM_COUNT | TERM_EPS, 30, 0.01), minEigThreshold=1e-4, flags=0,
              fbt=None):
    """cv2.calcOpticalFlowPyrLK keyword arguments -> vel_lk_params (same defaults and clamping)."""
    if flags != 0:
        raise NotImplementedError("only flags=0 is supported (the reference never passes flags)")
    ctype, count, eps = criteria
    if not (ctype & TERM_COUNT):
        count = 30
    if not (ctype & TERM_EPS):
        eps = 0.01
    p = _lib.LkParams()
    p.win_w, p.win_h = int(winSize[0]), int(winSize[1])
    p.max_level = int(maxLevel)
    p.max_count = int(count)
    p.eps = float(eps)
    p.min_eig_threshold = float(minEigThreshold)
    p.fb_threshold = -1.0 if fbt is None else float(fbt)
    return p


class FrameBatch:
    """F frames [F, H, pitch] uint8 resident in HBM plus their pyramids (levels >= 1)."""

    def __init__(self, frames, win, max_level):
        require_cuda()
        if frames.dim() == 2:
            frames = frames.unsqueeze(0)
        assert frames.is_cuda and frames.dtype == torch.uint8 and frames.dim() == 3 and frames.stride(2) == 1
        self.frames = frames
        self.n, self.h, self.w = frames.shape
        self.pitch = frames.stride(1)
        self.frame_stride = frames.stride(0) if self.n > 1 else 0
        self.layout = _lib.pyr_layout(self.w, self.h, win, max_level)
        nbytes = max(int(self.layout.bytes), 16)
        self.pyr = torch.empty((self.n, nbytes), dtype=torch.uint8, device=frames.device)
        self.built = False

    def build(self):
        """K1 over all frames of the batch (one launch per level)."""
        L = _lib.lib()
        _lib.check(L.vel_pyramid_u8(ptr(self.frames), self.frame_stride, self.pitch, self.n, C.byref(self.layout),
                                    ptr(self.pyr), self.pyr.stride(0), stream_ptr()), "vel_pyramid_u8")
        self.built = True
        return self

    def pyramid_launches(self):
        """Kernel launches of one build(): levels 1 and 2 share one fused kernel when the frames are 16-byte aligned rows of a
        width that is a multiple of 16 (csrc/pyramid.cu), every further level is one launch."""
        import os

        lv = self.layout.max_level
        fused = (lv >= 2 and self.w % 16 == 0 and self.w >= 32 and self.h >= 16 and self.pitch % 16 == 0
                 and self.frames.data_ptr() % 16 == 0 and self.frame_stride % 16 == 0 and os.environ.get("VEL_PYR_FUSED", "1")[:1] != "0")
        return lv - 1 if fused else lv

    def level(self, i, l):
        """Level l of frame i as a [h, w] uint8 view (for tests)."""
        if l == 0:
            return self.frames[i]
        lay = self.layout
        off, pitch, w, h = lay.offset[l], lay.pitch[l], lay.width[l], lay.height[l]
        return self.pyr[i, off:off + pitch * h].view(h, pitch)[:, :w]


def track_pairs(prev, next_, pts, params, prev_first=0, next_first=0, npairs=None, want_back=False):
    """K2 on `npairs` pairs: prev frame prev_first+k -> next frame next_first+k.

    prev/next_ are FrameBatch objects with identical geometry (they may be the same object for a
    sequence: prev_first=0, next_first=1).  pts is a CUDA float32 tensor [npts, 2] (shared by all
    pairs) or [npairs, npts, 2].  Returns (next_pts [npairs,npts,2] f32, status [npairs,npts] u8,
    err [npairs,npts] f32, back_pts or None), all CUDA tensors; nothing is synchronised."""
    if not (prev.built and next_.built):
        raise RuntimeError("FrameBatch.build() must run before tracking")
    assert (prev.w, prev.h) == (next_.w, next_.h)
    assert prev.layout.max_level == next_.layout.max_level
    if npairs is None:
        npairs = min(prev.n - prev_first, next_.n - next_first)
    pts = pts.contiguous()
    assert pts.is_cuda and pts.dtype == torch.float32 and pts.shape[-1] == 2
    npts = pts.shape[-2]
    pts_stride = 0 if pts.dim() == 2 else npts * 2
    if pts.dim() == 3:
        assert pts.shape[0] == npairs
    dev = pts.device
    out = torch.empty((npairs, npts, 2), dtype=torch.float32, device=dev)
    status = torch.empty((npairs, npts), dtype=torch.uint8, device=dev)
    err = torch.empty((npairs, npts), dtype=torch.float32, device=dev)
    back = torch.empty((npairs, npts, 2), dtype=torch.float32, device=dev) if want_back else None
    L = _lib.lib()
    pf = prev.frames.data_ptr() + prev_first * prev.frames.stride(0)
    nf = next_.frames.data_ptr() + next_first * next_.frames.stride(0)
    pp = prev.pyr.data_ptr() + prev_first * prev.pyr.stride(0)
    npy = next_.pyr.data_ptr() + next_first * next_.pyr.stride(0)
    ps = prev.frames.stride(0) if npairs > 1 else 0
    ns = next_.frames.stride(0) if npairs > 1 else 0
    _lib.check(L.vel_lk_track(C.c_void_p(pf), ps, prev.pitch, C.c_void_p(pp), prev.pyr.stride(0), C.c_void_p(nf), ns,
                              next_.pitch, C.c_void_p(npy), next_.pyr.stride(0), C.byref(prev.layout), npairs, ptr(pts),
                              pts_stride, npts, C.byref(params), ptr(out), ptr(status), ptr(err), ptr(back), stream_ptr()),
               "vel_lk_track")
    return out, status, err, back
