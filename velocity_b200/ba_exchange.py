"""Multi-GPU exchange of the bundle adjustment: cameras (frames) shard across ranks, points are shared.

Per LM iteration there are exactly TWO data-path collectives before the solve:

  1. ONE all-reduce (sum) of the "small" buffer  [cost | V (nt x 6) | g (3nt + 6nc) | U (nc x 21)]:
     V, the point part of g and cost are PARTIAL sums over each rank's cameras; the camera entries (U rows, camera part of g)
     are non-zero on their owner only and zero elsewhere, so the same sum also gathers them (x + 0 is exact).
  2. ONE all_gather_into_tensor of the cross blocks: the buffer W_ext [6 * per * world][3nt] has one 6-row block per CAMERA
     (camera 0's block is unused padding, so that every rank's `per` consecutive cameras are one contiguous, equally sized
     row block); the matrix the solver sees is W = W_ext[6:] (camera c >= 1 owns rows 6(c-1)..6c-1).  Gathered in place.

and two small ones inside the solve (velocity_b200.NLS.BundleAdjuster.solve): the ranks' tile rows of the reduced camera system
S go to the owner (rank 0) point-to-point, the owner factors S once and broadcasts delta_c -- S is not factored on every rank.
Works on any torch.distributed backend (NCCL on the GPUs; gloo in the CPU test of the exchange logic).
"""
import torch
import torch.distributed as dist


def camera_slices(nc, world):
    """(per, slices): cameras 0..nc split into `world` blocks of `per` = ceil((nc+1)/world) consecutive cameras
    (frames shard naturally: a rank owns the cameras of its own frame block); slices[r] = (first, count) in the camera
    indexing of vel_ba_accumulate.  Camera 0 is the fixed one (no parameters, utils/NLS.py:207-208)."""
    per = max(1, -(-(nc + 1) // world))
    out = []
    for r in range(world):
        lo, hi = min(nc + 1, r * per), min(nc + 1, (r + 1) * per)
        out.append((lo, hi - lo))
    return per, out


def param_rows(first, count):
    """Parameterised-camera row range [lo, hi) (camera c >= 1 owns row c-1) of a camera slice."""
    return max(first, 1) - 1, max(first + count - 1, 0)


def small_buffer(nt, nc, device, dtype=torch.float64):
    """The all-reduce buffer and its views (cost [1], V [nt,6], g [3nt+6nc], U [nc,21])."""
    n = 2 + 6 * nt + 3 * nt + 6 * nc + 21 * max(nc, 1)
    buf = torch.zeros((n,), dtype=dtype, device=device)
    o = 2
    V = buf[o:o + 6 * nt].view(nt, 6)
    o += 6 * nt
    g = buf[o:o + 3 * nt + 6 * nc]
    o += 3 * nt + 6 * nc
    U = buf[o:o + 21 * max(nc, 1)].view(max(nc, 1), 21)
    return buf, buf[0:1], V, g, U


def exchange_blocks(small, W_ext, per, rank, world, group=None):
    """The two collectives.  `small` holds this rank's partial sums / own camera entries (zeros for foreign cameras);
    W_ext [6*per*world, 3nt] holds this rank's cameras in its own row block."""
    dist.all_reduce(small, group=group)
    rows = 6 * per
    dist.all_gather_into_tensor(W_ext, W_ext[rank * rows:(rank + 1) * rows], group=group)


def tile_row_ranges(nb, world):
    """Contiguous ranges of the nb tile rows of the lower-triangular product, balanced by tile count (row bi has bi+1 tiles)."""
    total = nb * (nb + 1) // 2
    bounds, acc, r = [0], 0, 1
    for bi in range(nb):
        acc += bi + 1
        while r < world and acc >= total * r / world:
            bounds.append(bi + 1)
            r += 1
    while len(bounds) < world + 1:
        bounds.append(nb)
    bounds[-1] = nb
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def gather_rows_to_owner(S, row_ranges, rank, owner=0, group=None):
    """Rank r holds rows row_ranges[r] of the row-major matrix S; the owner receives every other rank's rows in place."""
    ops = []
    if rank == owner:
        for r, (lo, hi) in enumerate(row_ranges):
            if r != owner and hi > lo:
                ops.append(dist.P2POp(dist.irecv, S[lo:hi], _global(r, group), group))
    else:
        lo, hi = row_ranges[rank]
        if hi > lo:
            ops.append(dist.P2POp(dist.isend, S[lo:hi], _global(owner, group), group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def _global(group_rank, group):
    return group_rank if group is None else dist.get_global_rank(group, group_rank)
