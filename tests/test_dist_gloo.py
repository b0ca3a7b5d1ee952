"""World-size-2 gloo test (CPU) of the multi-GPU bundle-adjustment exchange: every rank computes
the partial blocks of its camera slice (here with the numpy oracle standing in for K7), runs
velocity_b200.ba_exchange.exchange_blocks, and must end up with the full single-process system."""
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
        from velocity_b200.ba_exchange import camera_slices, exchange_blocks

        g = golden("ba_medium")
        z, x, nt, nc = S._ba_pack(g["P"], g["pw0"], g["cw0"])
        slices = camera_slices(nc, world)
        first, count = slices[rank]
        V, U, W, gg, cost = S.ba_blocks(g["K"], x, z, nt, nc, first, count)
        iu3, iu6 = np.triu_indices(3), np.triu_indices(6)
        tV = torch.from_numpy(np.ascontiguousarray(V[:, iu3[0], iu3[1]]))
        tU = torch.from_numpy(np.ascontiguousarray(U[:, iu6[0], iu6[1]]))
        tW = torch.from_numpy(np.ascontiguousarray(W.transpose(0, 2, 1, 3).reshape(6 * nc, 3 * nt)))
        tg = torch.from_numpy(gg.copy())
        tc = torch.tensor([cost], dtype=torch.float64)
        exchange_blocks(tV, tU, tW, tg, tc, nt, nc, slices)
        Vf, Uf, Wf, gf, cf = S.ba_blocks(g["K"], x, z, nt, nc)
        ok = (np.allclose(tV.numpy(), Vf[:, iu3[0], iu3[1]], rtol=1e-12, atol=1e-9)
              and np.array_equal(tU.numpy(), Uf[:, iu6[0], iu6[1]])
              and np.array_equal(tW.numpy(), Wf.transpose(0, 2, 1, 3).reshape(6 * nc, 3 * nt))
              and np.allclose(tg.numpy(), gf, rtol=1e-12, atol=1e-9) and abs(tc.item() - cf) <= 1e-12 * cf)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_camera_slices_cover_all_cameras():
    sys.path.insert(0, ROOT)
    from velocity_b200.ba_exchange import camera_slices, param_rows

    for nc in (0, 1, 7, 299):
        for world in (1, 2, 3, 8):
            sl = camera_slices(nc, world)
            assert sum(c for _, c in sl) == nc + 1 and sl[0][0] == 0
            rows = [param_rows(f, c) for f, c in sl]
            covered = sorted(r for lo, hi in rows for r in range(lo, hi))
            assert covered == list(range(nc))


@pytest.mark.timeout(180)
def test_ba_exchange_world2_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
