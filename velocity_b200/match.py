"""GPU all-pairs descriptor matcher (K4) -- cv2.BFMatcher(norm).knnMatch(q, t, k=2) semantics."""
import numpy as np
import torch

from . import _lib
from .device import ptr, require_cuda, stream_ptr, to_device


def knn2_hamming256(q, t):
    """q [nq,32] uint8, t [nt,32] uint8 -> (idx int32 [nq,2], dist int32 [nq,2]); numpy or CUDA in,
    same kind out."""
    require_cuda()
    on_dev = isinstance(q, torch.Tensor) and q.is_cuda
    dq, dt = to_device(q, np.uint8), to_device(t, np.uint8)
    if dq.dim() != 2 or dq.shape[1] != 32 or dt.dim() != 2 or dt.shape[1] != 32:
        raise ValueError("expected 256-bit descriptors [n, 32] uint8")
    idx = torch.empty((dq.shape[0], 2), dtype=torch.int32, device=dq.device)
    dist = torch.empty((dq.shape[0], 2), dtype=torch.int32, device=dq.device)
    _lib.check(_lib.lib().vel_match_knn2_hamming256(ptr(dq), dq.shape[0], ptr(dt), dt.shape[0], ptr(idx), ptr(dist),
                                                    stream_ptr()), "vel_match_knn2_hamming256")
    return (idx, dist) if on_dev else (idx.cpu().numpy(), dist.cpu().numpy())


def knn2_l2(q, t):
    require_cuda()
    on_dev = isinstance(q, torch.Tensor) and q.is_cuda
    dq, dt = to_device(q, np.float32), to_device(t, np.float32)
    if dq.dim() != 2 or dt.dim() != 2 or dq.shape[1] != dt.shape[1]:
        raise ValueError("descriptor shapes do not match")
    idx = torch.empty((dq.shape[0], 2), dtype=torch.int32, device=dq.device)
    dist = torch.empty((dq.shape[0], 2), dtype=torch.float32, device=dq.device)
    _lib.check(_lib.lib().vel_match_knn2_l2(ptr(dq), dq.shape[0], ptr(dt), dt.shape[0], dq.shape[1], ptr(idx), ptr(dist),
                                            stream_ptr()), "vel_match_knn2_l2")
    return (idx, dist) if on_dev else (idx.cpu().numpy(), dist.cpu().numpy())


def ratio_test(idx, dist, ratio=0.6):
    """Lowe ratio test of utils/KLT.py:26: keep query q when d0 < ratio * d1.  Returns (query idx, train idx)."""
    idx, dist = np.asarray(idx), np.asarray(dist)
    keep = (idx[:, 1] >= 0) & (dist[:, 0] < ratio * dist[:, 1])
    qi = np.nonzero(keep)[0]
    return qi, idx[qi, 0]
