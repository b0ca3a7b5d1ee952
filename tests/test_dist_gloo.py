"""World-size-2 gloo tests (CPU) of the multi-GPU bundle-adjustment exchange (velocity_b200/ba_exchange.py): every rank
computes the partial blocks of its camera slice (here with the numpy oracle standing in for K7), runs the ONE all-reduce +
ONE all-gather of exchange_blocks, and must end up with the full single-process system; the owner collects the tile rows of
the reduced system point-to-point."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import sfm_oracle as S
        from util import golden
        from velocity_b200.ba_exchange import (camera_slices, exchange_blocks, gather_rows_to_owner, param_rows, small_buffer,
                                               tile_row_ranges)

        g = golden("ba_medium")
        z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
        per, slices = camera_slices(nc, world)
        first, count = slices[rank]
        V, U, W, gg, cost = S.ba_blocks(g["K"], x, z, nt, nc, first, count)
        iu3, iu6 = np.triu_indices(3), np.triu_indices(6)
        small, tc, tV, tg, tU = small_buffer(nt, nc, "cpu")
        tW_ext = torch.zeros((6 * per * world, 3 * nt), dtype=torch.float64)
        tW = tW_ext[6:]
        lo, hi = param_rows(first, count)
        tV.copy_(torch.from_numpy(np.ascontiguousarray(V[:, iu3[0], iu3[1]])))
        tc[0] = cost
        tg[:3 * nt] = torch.from_numpy(gg[:3 * nt].copy())
        tU[lo:hi] = torch.from_numpy(np.ascontiguousarray(U[lo:hi][:, iu6[0], iu6[1]]))
        tg[3 * nt + 3 * lo:3 * nt + 3 * hi] = torch.from_numpy(gg[3 * nt + 3 * lo:3 * nt + 3 * hi].copy())
        tg[3 * nt + 3 * nc + 3 * lo:3 * nt + 3 * nc + 3 * hi] = torch.from_numpy(gg[3 * nt + 3 * nc + 3 * lo:3 * nt + 3 * nc + 3 * hi].copy())
        tW[6 * lo:6 * hi] = torch.from_numpy(np.ascontiguousarray(W.transpose(0, 2, 1, 3).reshape(6 * nc, 3 * nt)[6 * lo:6 * hi]))
        exchange_blocks(small, tW_ext, per, rank, world)
        Vf, Uf, Wf, gf, cf = S.ba_blocks(g["K"], x, z, nt, nc)
        ok = (np.allclose(tV.numpy(), Vf[:, iu3[0], iu3[1]], rtol=1e-12, atol=1e-9)
              and np.array_equal(tU.numpy()[:nc], Uf[:, iu6[0], iu6[1]])                      # gathered through the sum: x + 0 is exact
              and np.array_equal(tW.numpy()[:6 * nc], Wf.transpose(0, 2, 1, 3).reshape(6 * nc, 3 * nt))
              and np.array_equal(tg.numpy()[3 * nt:], gf[3 * nt:])
              and np.allclose(tg.numpy()[:3 * nt], gf[:3 * nt], rtol=1e-12, atol=1e-9) and abs(tc.item() - cf) <= 1e-12 * cf)
        # tile rows of a row-major matrix travel to the owner
        n6 = 6 * nc
        ranges = [(min(n6, a * 16), min(n6, b * 16)) for a, b in tile_row_ranges((n6 + 15) // 16, world)]
        full = torch.arange(n6 * n6, dtype=torch.float64).view(n6, n6)
        mine = torch.zeros_like(full)
        mine[ranges[rank][0]:ranges[rank][1]] = full[ranges[rank][0]:ranges[rank][1]]
        gather_rows_to_owner(mine, ranges, rank, 0)
        if rank == 0:
            ok = ok and torch.equal(mine, full)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_camera_slices_cover_all_cameras():
    sys.path.insert(0, ROOT)
    from velocity_b200.ba_exchange import camera_slices, param_rows, tile_row_ranges

    for nc in (0, 1, 7, 299, 2399):
        for world in (1, 2, 3, 8):
            per, sl = camera_slices(nc, world)
            assert sum(c for _, c in sl) == nc + 1 and sl[0][0] == 0
            rows = [param_rows(f, c) for f, c in sl]
            covered = sorted(r for lo, hi in rows for r in range(lo, hi))
            assert covered == list(range(nc))
            for r, (f, c) in enumerate(sl):                     # every rank's cameras sit inside its equally sized gather block
                assert c == 0 or (f >= r * per and f + c <= (r + 1) * per)
    for nb in (1, 2, 15, 113):
        for world in (1, 2, 4, 8):
            rr = tile_row_ranges(nb, world)
            assert rr[0][0] == 0 and rr[-1][1] == nb and all(a[1] == b[0] for a, b in zip(rr[:-1], rr[1:]))
            if nb >= 4 * world:                                  # balanced by tile count, not by row count
                w = [sum(bi + 1 for bi in range(lo, hi)) for lo, hi in rr]
                assert max(w) <= 1.5 * (sum(w) / world)


@pytest.mark.timeout(180)
def test_ba_exchange_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
