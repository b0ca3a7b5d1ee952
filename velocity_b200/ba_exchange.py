"""Multi-GPU exchange step of the bundle adjustment.

Cameras 0..nc are split into `world` contiguous slices.  After K7 every rank holds
  * PARTIAL sums over its cameras in V [nt,6], g[:3nt] (point part) and cost  -> all-reduce (sum)
  * the rows of U [nc,21], W [6nc,3nt] and the camera part of g that belong to its cameras
                                                                               -> all-gather
so that every rank ends up with the identical full system and solves it redundantly.  Rank slices
can be uneven (camera 0 carries no parameters), so the gather is expressed as one broadcast per
owner into the owner's row block of the preallocated full buffers: no packing copy, no padding.
Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch.distributed as dist


def camera_slices(nc, world):
    """Contiguous split of camera indices 0..nc (inclusive) into `world` (first, count) slices."""
    bounds = [(nc + 1) * r // world for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1] - bounds[r]) for r in range(world)]


def param_rows(first, count):
    """Parameterised-camera row range [lo, hi) (camera c>=1 owns row c-1) of a camera slice."""
    return max(first, 1) - 1, max(first + count - 1, 0)


def exchange_blocks(V, U, W, g, cost, nt, nc, slices, group=None):
    dist.all_reduce(V, group=group)
    dist.all_reduce(g[:3 * nt], group=group)
    dist.all_reduce(cost, group=group)
    for owner, (first, count) in enumerate(slices):
        lo, hi = param_rows(first, count)
        if hi <= lo:
            continue
        dist.broadcast(U[lo:hi], src=owner, group=group)
        dist.broadcast(W[6 * lo:6 * hi], src=owner, group=group)
        dist.broadcast(g[3 * nt + 3 * lo:3 * nt + 3 * hi], src=owner, group=group)
        dist.broadcast(g[3 * nt + 3 * nc + 3 * lo:3 * nt + 3 * nc + 3 * hi], src=owner, group=group)
