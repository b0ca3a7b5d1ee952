"""torchrun --nproc-per-node N tools/ba_shard_check.py : the camera-sharded bundle adjustment over NCCL (real K7 per rank, one
all-reduce + one all-gather, tile rows of S to the owner, owner-only Cholesky, broadcast) must reproduce the unsharded run,
the reference golden and the C3-size oracle fixture, with bit-identical parameters on every rank.  Prints per-rank verdicts,
exits non-zero on any mismatch.  (Run on the GPU box; tests/test_dist_nccl.py wraps it.)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import sfm_oracle as S  # noqa: E402  (checker only)
from util import ba_c3_inputs, golden  # noqa: E402
from velocity_b200 import NLS  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok_all = True


def run_case(name, K, P, pw0, cw0, want_cw, want_pw, tol, iters):
    global ok_all
    z, x, nt, nc = S._ba_pack(P, pw0, cw0)
    ba = NLS.BundleAdjuster(K, z, x, nt, nc, shard=True)
    single = NLS.BundleAdjuster(K, z, x, nt, nc, shard=False)
    # one linearisation: the exchanged blocks equal the unsharded ones
    ba.accumulate()
    single.accumulate()
    n6 = 6 * nc
    blocks_ok = (torch.equal(ba.U[:nc], single.U[:nc]) and torch.equal(ba.W[:n6], single.W[:n6])
                 and torch.equal(ba.g[3 * nt:], single.g[3 * nt:])
                 and torch.allclose(ba.V, single.V, rtol=1e-12, atol=1e-9) and torch.allclose(ba.g[:3 * nt], single.g[:3 * nt], rtol=1e-11, atol=1e-7))
    ba.x.copy_(single.x)
    n_it = 0
    for it in range(iters):
        f, xr = ba.step()
        f1, xr1 = single.step()
        n_it += 1
        if xr < 1e-7:
            break
    xs, x1 = ba.x.cpu().numpy(), single.x.cpu().numpy()
    cw, pw = S._ba_unpack(xs, nt, nc)
    ok_ref = np.abs(cw - want_cw).max() <= tol * np.abs(want_cw).max() and np.abs(pw - want_pw).max() <= tol * np.abs(want_pw).max()
    ok_single = np.allclose(xs, x1, rtol=1e-7, atol=1e-8)        # the all-reduce sums the partial V in another order
    gathered = [torch.empty_like(ba.x) for _ in range(world)]
    dist.all_gather(gathered, ba.x)
    identical = all(torch.equal(gathered[0], t) for t in gathered)
    print("rank %d/%d %-10s nt=%d nc=%d iterations %d: blocks == unsharded %s, matches fixture %s, matches single-GPU %s, "
          "bit-identical across ranks %s" % (rank, world, name, nt, nc, n_it, blocks_ok, ok_ref, ok_single, identical), flush=True)
    ok_all = ok_all and blocks_ok and ok_ref and ok_single and identical


g = golden("ba_medium")
run_case("ba_medium", g["K"], g["P"], g["pw0"], g["cw0"], g["cw"], g["pw"], 1e-6, 10)
g = golden("ba_512x20")
run_case("ba_512x20", g["K"], g["P"], g["pw0"], g["cw0"], g["cw"], g["pw"], 1e-6, 10)
if "--c3" in sys.argv:
    g = golden("ba_c3_sparse")
    K, P, pw0, cw0 = ba_c3_inputs()
    run_case("c3", K, P, pw0, cw0, g["cw"], g["pw"], 1e-4, 10)
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
