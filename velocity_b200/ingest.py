"""Frame ingest on the GPU: the colour conversion the reference's driver does
per frame on the host (cv2.cvtColor(imbgr, cv2.COLOR_BGR2GRAY), vidExample.py:91), bit-exact in
OpenCV's 15-bit fixed point, so decoded BGR frames can go straight to HBM and stay there."""
import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_ptr


def bgr2gray(bgr):
    """uint8 [H,W,3] or [F,H,W,3] (numpy or CUDA tensor) -> gray uint8 [H,W] / [F,H,W]; same kind out."""
    require_cuda()
    on_dev = isinstance(bgr, torch.Tensor) and bgr.is_cuda
    t = bgr if on_dev else torch.from_numpy(np.ascontiguousarray(np.asarray(bgr, np.uint8))).cuda()
    single = t.dim() == 3
    if single:
        t = t.unsqueeze(0)
    if t.dim() != 4 or t.shape[-1] != 3 or t.dtype != torch.uint8:
        raise ValueError("expected uint8 BGR frames [H,W,3] or [F,H,W,3]")
    t = t.contiguous()
    F, H, W, _ = t.shape
    out = torch.empty((F, H, W), dtype=torch.uint8, device=t.device)
    _lib.check(_lib.lib().vel_bgr2gray_u8(ptr(t), H * W * 3, W * 3, F, W, H, ptr(out), H * W, W, stream_ptr()), "vel_bgr2gray_u8")
    out = out[0] if single else out
    return out if on_dev else out.cpu().numpy()
