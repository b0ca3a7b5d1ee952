"""Device plumbing (torch owns memory and streams; the kernels live in libvelocity_b200.so)."""
import ctypes as C

import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("velocity_b200: no CUDA device visible -- the accelerated path has no CPU fallback")
    _lib.lib()


def stream_ptr():
    # the raw handle of torch's current stream on the current device (torch.cuda.current_stream() builds a Stream object: ~15 us a call,
    # and the per-frame drop-in path makes a dozen C calls per frame)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def to_device(a, dtype=None):
    """numpy / torch(any device) -> contiguous CUDA tensor (no copy when already suitable)."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        a = np.asarray(a)
        if dtype is not None and a.dtype != np.dtype(dtype):
            a = a.astype(dtype)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        td = getattr(torch, np.dtype(dtype).name)
        if t.dtype != td:
            t = t.to(td)
    if not t.is_cuda:
        t = t.cuda(non_blocking=False)
    return t.contiguous()


def image_view(im):
    """2-D uint8 image -> (cuda tensor keeping it alive, data pointer, width, height, pitch).

    CUDA tensors are used in place when their last stride is 1 (ROI views of a resident frame cost
    nothing); host arrays are uploaded (only the ROI when a numpy slice is passed)."""
    if isinstance(im, torch.Tensor) and im.is_cuda:
        if im.dtype != torch.uint8 or im.dim() != 2:
            raise ValueError("expected a 2-D uint8 image")
        if im.stride(1) != 1 or im.stride(0) < im.shape[1]:
            im = im.contiguous()
        return im, im.data_ptr(), im.shape[1], im.shape[0], im.stride(0)
    a = np.asarray(im) if not isinstance(im, torch.Tensor) else im.numpy()
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("expected a 2-D uint8 image, got %s %s" % (a.dtype, a.shape))
    h, w = a.shape
    if w % 4 == 0:
        t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    else:
        # rows padded to a multiple of 16 bytes: the word-gathering 15x15 LK kernel needs pitch % 4 == 0
        buf = torch.empty((h, (w + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
        t = buf[:, :w]
        t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t, t.data_ptr(), t.shape[1], t.shape[0], t.stride(0)
